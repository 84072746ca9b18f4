#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched SustainDC step (BASELINE.json metric) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs 65536] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): 65 536 envs per GPU, synthetic
1-year NY traces (seed 1234), default dc_config (20 racks x 200 CPUs), 7-day episodes with de-synchronised
phases so that ~N/672 envs auto-reset every step, reward windows pre-filled to H = 10 000 with N(330, 40) kWh,
uniform random actions.  A "step" = one sdc_step over all envs of a rank.

  value      whole-job env-steps/s, inputs resident in HBM, CUDA events around the K timed steps (max over ranks)
  e2e        the same metric through the host-buffer C-ABI call (numpy in / numpy out): per step H2D of the
             actions and D2H of obs / share_obs / rewards / dones are inside the timed region
  roofline   algorithmic bytes (4*H + 1024 = 41 024 B per env-step, the fixed numerator) / k_step
             launch time vs the measured HBM copy bandwidth in MEASURED_PEAKS.json.  k_step maintains the reward
             normaliser incrementally and streams a window only when an env's incremental state needs a refresh
             (DESIGN.md section 4), so it touches ~8x fewer physical bytes than the algorithmic figure: `frac` > 1
             is expected ("report both"); `traffic` is the ncu DRAM figure of the same launch
  cpu_baseline  the oracle port of the reference's SustainDC.step (oracle/sdc_oracle.py) on the host cores,
             one env per process, reward window pre-filled the same way, bounded sample
`--impl reference` prints the CPU arm alone (the reference is pure Python/numpy: its own implementation of this
path IS the CPU path; /root/reference does not exist on the GPU box, so the pinned oracle port stands in).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

B_ALG_STEADY = 4 * 10000 + 1024      # bytes per env-step at H = 10 000
# dram__bytes_read.sum + dram__bytes_write.sum of one k_step launch at N = 65 536 (launch 151 of the bench workload, ~390
# window refreshes in flight) from the round-1 `ncu --set full` capture (profiles/r01_kstep_ncu_raw.csv); only
# meaningful for the default --envs
TRAFFIC_BYTES_PER_LAUNCH = 236.845312e6 + 76.829440e6
METRIC = "env-steps/sec at N=65536 parallel envs, 1/2/4/8xB200; HBM GB/s fraction"


def measured_peak():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if any(r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference step, one env per process
# ---------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    rank, n_steps, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import random
    import sdc_oracle                                     # CPU baseline leg: the only place bench.py runs oracle/
    from dc_rl_b200.traces import LocationTraces
    tr = LocationTraces.synthetic("ny", 1234)
    n = 35040
    o_tr = sdc_oracle.Traces.__new__(sdc_oracle.Traces)
    o_tr.workload, o_tr.ci, o_tr.temp_base, o_tr.wetb_base = tr.workload[:n + 32], tr.ci[:n + 32], tr.temp_base[:n], tr.wetb_base[:n]
    env = sdc_oracle.OracleEnv(o_tr, "ny", 6, 7)
    random.seed(seed + rank); np.random.seed(seed + rank)
    rng = np.random.RandomState(5678 + rank)
    env.history.extend((330 + 40 * rng.standard_normal(10000)).tolist())
    env.reset()
    for _ in range(20):
        env.step(*rng.randint(0, 3, 3))
    t0 = time.perf_counter()
    for _ in range(n_steps):
        _, _, term, _ = env.step(*rng.randint(0, 3, 3))
        if term:
            env.reset()
    return n_steps, time.perf_counter() - t0


def cpu_arm(steps_per_proc, cores=None):
    import multiprocessing as mp
    cores = cores or min(os.cpu_count() or 1, 64)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(r, steps_per_proc, 91011) for r in range(cores)])
        wall = time.perf_counter() - t0
    per_proc = [n / dt for n, dt in res]
    total = sum(per_proc)
    return dict(value=total, unit="env-steps/s", cores=cores, kind="port",
                sample="%d processes x %d warm steps of oracle/sdc_oracle.py OracleEnv.step (H=10000 pre-filled, NY synthetic "
                       "traces, random actions, auto-reset); per-core %.1f steps/s; pool wall %.1f s" % (
                           cores, steps_per_proc, total / cores, wall))


# ---------------------------------------------------------------------------------------------------
def build_engine(n_envs, device, seed_base=0):
    from dc_rl_b200.dc_config import size_datacenter
    from dc_rl_b200.engine import Engine
    from dc_rl_b200.traces import LocationTraces
    traces = LocationTraces.synthetic("ny", 1234)
    params, derived = size_datacenter("ny")
    ids = np.arange(n_envs)
    months = np.where(ids < 12, ids % 12, ids % 3 + 5)            # make_train_env rule, harl/utils/envs_tools.py:56-62
    eng = Engine(n_envs, [traces], [params], months=months, seeds=(ids + seed_base).astype(np.uint64) * 1000 + 91011,
                 days_per_episode=7, device=device)
    return eng, derived


def prepare(eng, n_envs, rank):
    """Reset all envs, pre-fill the reward windows to H = 10 000 and de-synchronise the episode phases."""
    rng = np.random.default_rng(5678 + rank)
    pool = (330.0 + 40.0 * rng.standard_normal((256, 10000), dtype=np.float32)).astype(np.float32)
    hist = pool[rng.integers(0, 256, n_envs)]
    hist += rng.standard_normal((n_envs, 1), dtype=np.float32)    # distinct windows per env
    eng.prefill_history(hist)
    del hist
    eng.reset_host()
    phase = rng.integers(0, eng.ep_len - 1, n_envs).astype(np.int32)
    eng.write_state("step_in_ep", phase)
    eng.write_state("t", eng.read_state("t0").astype(np.int32) + phase)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-steps", type=int, default=1500)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--tune", default="", help="k=v,... forwarded to sdc_set_tuning")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = {"workload": "configs[2]: N=65536 envs/GPU, synthetic 1-year NY traces, default dc_config 20 racks x 200 CPUs, "
                       "7-day episodes (desynchronised), H=10000 pre-filled, random actions",
           "envs_per_gpu": args.envs, "n_envs_total": args.envs * max(world, 1), "history_len": 10000,
           "l2": "inputs larger than L2 (reward windows: %.2f GB per GPU)" % (args.envs * 40000 / 1e9)}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(200, min(args.cpu_steps, 100 * max(args.steps, 1)))
        cb = cpu_arm(steps)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "env-steps/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = args.envs
    eng, _ = build_engine(n, local, seed_base=rank * n)
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        eng.set_tuning(**{k: int(v)})
    prepare(eng, n, rank)

    obs = torch.zeros(n, 3, 26, device=dev); share = torch.zeros(n, 29, device=dev); rew = torch.zeros(n, 3, device=dev)
    done = torch.zeros(n, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev); g.manual_seed(5678 + rank)
    n_act = 8
    acts = [torch.randint(0, 3, (n, 3), dtype=torch.int32, device=dev, generator=g) for _ in range(n_act)]
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_step(i):
        eng.step_device(acts[i % n_act], obs, share, rew, done, None, None, stream.cuda_stream)

    for i in range(max(args.warmup, 3)):
        one_step(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    eng.set_tuning(timing=1)
    eng.kernel_times()
    launches0 = eng.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record(stream)
    for i in range(args.steps):
        one_step(i)
        ev[i + 1].record(stream)
    if world > 1:
        m = torch.tensor(eng.metrics(), device=dev)        # the path's one collective: episode metrics
        gathered = [torch.zeros_like(m) for _ in range(world)]
        dist.all_gather(gathered, m)
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    launches = eng.launch_count - launches0
    ktimes = eng.kernel_times()             # CUDA events recorded by the library on the launch stream
    pass_stats = eng.read_state("pass_stats")   # window passes of the last timed step: plain, refresh, by brackets, by tails
    eng.set_tuning(timing=0)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = n * world * args.steps / (total_ms_max / 1e3)

    # ---- end to end through the host-buffer C-ABI call (numpy in, numpy out) ----
    rng = np.random.RandomState(5678 + rank)
    host_acts = [rng.randint(0, 3, size=(n, 3)).astype(np.int32) for _ in range(4)]
    for i in range(3):
        eng.step_host(host_acts[i % 4], want_info=False, want_term=False)
    e2e_steps = max(5, min(args.steps, 200))
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        eng.step_host(host_acts[i % 4], want_info=False, want_term=False)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n * world * e2e_steps / float(e2e_s.item())
    clocks = sampler.summary() if sampler else None
    err = int(np.bitwise_or.reduce(eng.read_state("err")))

    if rank == 0:
        peak, peak_src = measured_peak()
        k_ms = float(ktimes[1] / max(ktimes[0], 1))          # mean k_step launch duration over the timed region
        achieved = B_ALG_STEADY * n / (k_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 reward window / f64 scalar physics", "data": "synthetic",
            "config": cfg,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": n * 3 * 4,
                    "d2h_bytes_per_step": n * (78 + 29 + 3) * 4 + n, "steps": e2e_steps,
                    "call": "sdc_step_host (numpy actions in; obs, share_obs, rewards, dones out)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": TRAFFIC_BYTES_PER_LAUNCH if n == 65536 else None, "kernel": "k_step", "launch_ms_mean": k_ms,
                         "launch_ms_max": float(ktimes[3]),                          "step_ms_median": float(np.median(step_ms)),
                         "bytes_per_launch": B_ALG_STEADY * n, "peak_source": peak_src,
                         "physical_frac": (TRAFFIC_BYTES_PER_LAUNCH / (k_ms / 1e3) / 1e9 / peak) if n == 65536 else None,
                         "note": "algorithmic bytes are the fixed SURVEY 8d numerator; the kernel is incremental and latency / "
                                 "issue bound, not HBM bound (physical_frac = ncu DRAM bytes per launch / launch time / peak)",
                         "passes_last_step": [int(x) for x in pass_stats]},
            "clocks": clocks, "env_error_flags": err,
        }
        if not args.no_cpu and world >= 1:
            line["cpu_baseline"] = cpu_arm(args.cpu_steps)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
