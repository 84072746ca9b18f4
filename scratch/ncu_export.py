"""Turn one `ncu --set full --import-source on` capture of k_step into the tracked artefacts under profiles/:
    <tag>_kstep_ncu_raw.csv      every metric of the launch (ncu --page raw --csv)
    <tag>_kstep_ncu_details.txt  ncu --page details
    <tag>_kstep_hotspots.txt     per-source-line stall samples / instructions (scratch/ncu_hotspots.py)
    <tag>_kstep_ncu.json         the handful of numbers bench.py reads for `roofline.traffic` / `physical` / `limiters`
usage: python scratch/ncu_export.py gpurun_out/X.ncu-rep r02 [n_envs] [note]"""
import csv
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
n_envs = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
note = sys.argv[4] if len(sys.argv) > 4 else ""
out = os.path.join(REPO, "profiles")


def run(args):
    return subprocess.run(args, capture_output=True, text=True, check=True).stdout


raw = run(["ncu", "-i", rep, "--page", "raw", "--csv"])
open(os.path.join(out, tag + "_kstep_ncu_raw.csv"), "w").write(raw)
open(os.path.join(out, tag + "_kstep_ncu_details.txt"), "w").write(run(["ncu", "-i", rep, "--page", "details"]))
src = run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"])
tmp = "/tmp/_ncu_src.csv"
open(tmp, "w").write(src)
hot = run([sys.executable, os.path.join(REPO, "scratch", "ncu_hotspots.py"), tmp, "80"])
commit = run(["git", "-C", REPO, "rev-parse", "--short", "HEAD"]).strip()
open(os.path.join(out, tag + "_kstep_hotspots.txt"), "w").write("# capture of %s at commit %s %s\n" % (os.path.basename(rep), commit, note) + hot)

rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))


def f(name):
    return float(m[name].replace(",", ""))


unit = dict(zip(hdr, rows[1]))


def mbytes(name):
    v = f(name)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit[name].split("/")[0]]


stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: f(k) for k in hdr
          if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
tot = sum(stalls.values())
top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
cap = {
    "kernel": m["Kernel Name"], "commit": commit, "note": note, "n_envs": n_envs,
    "grid": int(f("launch__grid_size")), "block": int(f("launch__block_size")),
    "registers_per_thread": int(f("launch__registers_per_thread")),
    "smem_dynamic_bytes": int(mbytes("launch__shared_mem_per_block_dynamic")),
    "ncu_duration_us": f("gpu__time_duration.sum") * {"us": 1, "ms": 1e3, "ns": 1e-3}[unit["gpu__time_duration.sum"]],
    "dram_bytes_read": mbytes("dram__bytes_read.sum"), "dram_bytes_write": mbytes("dram__bytes_write.sum"),
    "dram_bytes_per_launch": mbytes("dram__bytes_read.sum") + mbytes("dram__bytes_write.sum"),
    "dram_throughput_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "l1_hit_pct": f("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": f("lts__t_sector_hit_rate.pct"),
    "issue_slot_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fp64_pipe_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "achieved_occupancy_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "ipc": f("sm__inst_executed.avg.per_cycle_active"),
    "warp_instructions": f("smsp__inst_executed.sum"),
    "top_stalls": {k: round(v / tot, 3) for k, v in top},
}
cap["dram_bytes_per_env_step"] = cap["dram_bytes_per_launch"] / n_envs
json.dump(cap, open(os.path.join(out, tag + "_kstep_ncu.json"), "w"), indent=1)
print(json.dumps(cap, indent=1))
