"""TEST / MEASUREMENT INFRASTRUCTURE -- copies the UNMODIFIED reference files of the hot path (and its callers under harl/)
into baseline/_ref/ so that the reference itself can be timed next to the CUDA library on the GPU box, where
/root/reference does not exist.  baseline/_ref/ is git-ignored (never product source, never in
history) but travels with `gpurun`.  Nothing under dc_rl_b200/ reads it.

    python oracle/setup_baseline_ref.py          # no-op when /root/reference is absent (the GPU box uses the copied tree)
"""
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SDC_REFERENCE_SRC", "/root/reference")
DST = os.path.join(REPO, "baseline", "_ref")
FILES = ["sustaindc_env.py", "train_sustaindc.py", "requirements.txt", "LICENSE"]
TREES = ["envs", "utils", "harl"]
DATA = ["Weather/__init__.py", "CarbonIntensity/__init__.py", "Workload/__init__.py", "__init__.py",
        "Weather/USA_NY_New.York-LaGuardia.epw", "Weather/USA_AZ_Phoenix-Sky.Harbor.epw", "Weather/USA_WA_Seattle-Tacoma.epw",
        "CarbonIntensity/NY_NG_&_avgCI.csv", "CarbonIntensity/AZ_NG_&_avgCI.csv", "CarbonIntensity/WA_NG_&_avgCI.csv",
        "Workload/Alibaba_CPU_Data_Hourly_1.csv"]


def main():
    if not os.path.isfile(os.path.join(SRC, "sustaindc_env.py")):
        print("setup_baseline_ref: %s not present; keeping %s as is (%s)" % (SRC, DST, "present" if os.path.isdir(DST) else "absent"))
        return 0
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        if os.path.isfile(os.path.join(SRC, f)):
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    for t in TREES:
        shutil.copytree(os.path.join(SRC, t), os.path.join(DST, t), dirs_exist_ok=True,
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "dashboard.py"))
    for d in DATA:
        src = os.path.join(SRC, "data", d)
        if os.path.isfile(src):
            os.makedirs(os.path.dirname(os.path.join(DST, "data", d)), exist_ok=True)
            shutil.copy2(src, os.path.join(DST, "data", d))
    n = sum(len(fs) for _, _, fs in os.walk(DST))
    print("setup_baseline_ref: %d files under %s" % (n, DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
