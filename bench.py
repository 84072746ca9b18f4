#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched SustainDC step (BASELINE.json metric) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 3|4] [--envs E] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (--config 3, BASELINE.json configs[2], the configuration the metric is quoted on): 65 536 envs per GPU, synthetic
1-year NY traces (seed 1234), default dc_config (20 racks x 200 CPUs), 7-day episodes with de-synchronised phases so that
~N/672 envs auto-reset every step, uniform random actions, reward windows pre-filled to H = 10 000 with N(330, 40) kWh
 and then turned over ORGANICALLY: `--settle` (default 12 000) untimed steps run before anything is timed, so
that every window holds the env's own energies and the rate of window refresh passes is stationary whatever --steps is.
--config 4 (BASELINE.json configs[3]): 32 768 envs per GPU, env i -> location {az, ny, wa}[i mod 3], geometry
{dc1, dc2, dc3}[(i // 3) mod 3] through the tolerant dc_config loader, metrics all-gathered every 1 024 steps.

A "step" = one sdc_step over all envs of a rank.  One JSON line:
  value           whole-job env-steps/s, inputs resident in HBM, EVERY output of the real call produced (obs, share_obs,
                  rewards, dones, the 59-column info table, terminal observations); two CUDA events around the K timed steps
                  and nothing else on the stream (an event after every step plus the library's two per launch cost the
                  stream 9 %: those live in a second, instrumented pass of the same K steps that only feeds `roofline`),
                  max over ranks.  `value_core_outputs`: the same without info / terminal observations (round-1 definition).
  e2e             the same metric through the host-buffer C-ABI call a numpy caller makes (sdc_step_compact_host: actions
                  in; the 29 distinct observation values per env, rewards, dones and terminal rows out), H2D + D2H inside the timed
                  region.  `e2e_padded`: sdc_step_host (obs[N,3,26] + share[N,29]); `e2e_vec_env`: CudaShareVecEnv.step, the
                  object harl.runners drive (zero-copy views / reference-like copies).
  roofline        `achieved` / `frac`: algorithmic bytes (4*H + 1024 = 41 024 B per env-step, the fixed
                  numerator) / k_step launch time vs the measured HBM copy bandwidth.  k_step keeps the reward normaliser
                  incrementally exact and streams a window only when its incremental state needs a refresh, so it moves far
                  fewer PHYSICAL bytes: `traffic` (ncu dram bytes of one steady-state launch, read from
                  profiles/r02_kstep_ncu.json), `physical` (GB/s and fraction of peak) and `limiters` (issue slots, fp64
                  pipe, occupancy from the same capture) say what bounds the kernel: latency, not bandwidth.
  cpu_baseline    the reference's own SustainDC.step (unmodified files under baseline/_ref, `kind: reference`) on the host
                  cores, one env per process, reward window pre-filled the same way, bounded sample; the oracle port
                  (`kind: port`) only when baseline/_ref is absent.
`--impl reference` prints that CPU arm alone.
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
METRIC = "env-steps/sec at N=65536 parallel envs, 1/2/4/8xB200; HBM GB/s fraction"
REF_ROOT = os.path.join(REPO, "baseline", "_ref")
NCU_JSON = os.path.join(REPO, "profiles", "r02_kstep_ncu.json")


def measured_peak():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs: NVML in-process (no fork next to the launching
    thread; `sample_now` is also called by the main thread right after the timed launches are enqueued, i.e. while the GPU
    executes them, so that even a 20-step region has a sample from inside it), nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows, self.lock = index, threading.Event(), [], threading.Lock()
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(index))
        except Exception:
            self.nv = None

    def sample_now(self):
        try:
            if self.nv:
                nv, h = self.nv
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = int(get(h))
                bits = (0x8, 0x40, 0x20, 0x4)           # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
                row = [str(sm), str(mx), "0"] + ["Active" if r & b else "Not Active" for b in bits]
            else:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if not out:
                    return
                row = [x.strip() for x in out.split(",")]
            with self.lock:
                self.rows.append(row)
        except Exception:
            pass

    def run(self):
        while not self.stop_flag.is_set():
            self.sample_now()
            self.stop_flag.wait(0.02 if self.nv else 0.05)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if any(r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(self.rows),
                "source": "nvml" if self.nv else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's own step (baseline/_ref), or the oracle port when that tree is absent
# ---------------------------------------------------------------------------------------------------
def _ref_worker(args):
    """One live reference env (unmodified files under baseline/_ref) in this process."""
    rank, n_steps, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["SDC_REFERENCE_ROOT"] = REF_ROOT
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import random
    import live_ref                                       # shims only (gymnasium / matplotlib / psychrolib / dashboard stubs)
    env = live_ref.fresh_env({"location": "ny", "month": 6, "days_per_episode": 7})
    from utils import reward_creator                      # the reference's module: its window is a per-process global
    random.seed(seed + rank); np.random.seed(seed + rank)
    rng = np.random.RandomState(5678 + rank)
    env.reset()
    reward_creator.energy_history.extend((330 + 40 * rng.standard_normal(10000)).tolist())
    agents = ("agent_ls", "agent_dc", "agent_bat")
    for _ in range(10):
        env.step(dict(zip(agents, rng.randint(0, 3, 3))))
    t0 = time.perf_counter()
    for _ in range(n_steps):
        _, _, _, trunc, _ = env.step(dict(zip(agents, rng.randint(0, 3, 3))))
        if trunc["__all__"]:
            env.reset()
    return n_steps, time.perf_counter() - t0


def _port_worker(args):
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
    for _ in range(10):
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
    live = os.path.isfile(os.path.join(REF_ROOT, "sustaindc_env.py"))
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_ref_worker if live else _port_worker, [(r, steps_per_proc, 91011) for r in range(cores)])
        wall = time.perf_counter() - t0
    total = sum(n / dt for n, dt in res)
    what = ("the reference's own sustaindc_env.SustainDC.step (unmodified files under baseline/_ref, real NY data files)" if live
            else "oracle/sdc_oracle.py OracleEnv.step (port; baseline/_ref absent)")
    return dict(value=total, unit="env-steps/s", cores=cores, kind="reference" if live else "port",
                sample="%d processes x %d warm steps of %s, H=10000 pre-filled, 7-day episodes, random actions, auto-reset; "
                       "per-core %.1f steps/s; pool wall %.1f s" % (cores, steps_per_proc, what, total / cores, wall))


# ---------------------------------------------------------------------------------------------------
def build_engine(n_envs, device, seed_base=0, config=3):
    from dc_rl_b200.dc_config import size_datacenter
    from dc_rl_b200.engine import Engine
    from dc_rl_b200.traces import LocationTraces
    ids = np.arange(n_envs) + seed_base
    months = np.where(ids < 12, ids % 12, ids % 3 + 5)            # make_train_env rule, harl/utils/envs_tools.py:56-62
    seeds = ids.astype(np.uint64) * 1000 + 91011
    if config == 4:
        with open(os.path.join(REPO, "tests", "golden", "dc_configs.json")) as f:      # the reference's dc_config_dc{1,2,3}.json
            cfgs = json.load(f)
        locs, geos = ["az", "ny", "wa"], ["dc1", "dc2", "dc3"]
        traces = [LocationTraces.synthetic(l, 1234) for l in locs]
        params, derived = [], None
        for g in geos:
            for l in locs:
                p, derived = size_datacenter(l, cfgs[g])
                params.append(p)
        loc_id = (ids % 3).astype(np.uint8)
        cfg_id = (((ids // 3) % 3) * 3 + ids % 3).astype(np.uint8)
        eng = Engine(n_envs, traces, params, loc_id=loc_id, cfg_id=cfg_id, months=months, seeds=seeds, days_per_episode=7, device=device)
        return eng, derived
    traces = LocationTraces.synthetic("ny", 1234)
    params, derived = size_datacenter("ny")
    eng = Engine(n_envs, [traces], [params], months=months, seeds=seeds, days_per_episode=7, device=device)
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


def ncu_capture():
    try:
        with open(NCU_JSON) as f:
            return json.load(f)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--config", type=int, default=3, choices=[3, 4])
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: 65536 for config 3, 32768 for config 4)")
    ap.add_argument("--settle", type=int, default=12000, help="untimed steps that turn the pre-filled reward windows over")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-steps", type=int, default=1500)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--tune", default="", help="k=v,... forwarded to sdc_set_tuning")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.envs or (65536 if args.config == 3 else 32768)
    if args.config == 3:
        workload = ("configs[2]: N=%d envs/GPU, synthetic 1-year NY traces, default dc_config 20 racks x 200 CPUs, 7-day episodes "
                    "(desynchronised), H=10000 pre-filled then turned over by %d untimed steps, random actions" % (n, args.settle))
    else:
        workload = ("configs[3]: N=%d envs/GPU, env i -> {az,ny,wa}[i%%3] x {dc1,dc2,dc3}[(i//3)%%3] (tolerant dc_config loader), synthetic "
                    "1-year traces per location, 7-day episodes (desynchronised), H=10000 pre-filled then turned over by %d untimed "
                    "steps, random actions, metric all-gather every 1024 steps" % (n, args.settle))
    cfg = {"workload": workload, "envs_per_gpu": n, "n_envs_total": n * max(world, 1), "history_len": 10000,
           "l2": "inputs larger than L2 (reward windows: %.2f GB per GPU)" % (n * 40000 / 1e9)}

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
    from dc_rl_b200 import numa
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    placement = numa.bind_to_gpu(local)          # before the engine allocates its pinned host buffers
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    eng, _ = build_engine(n, local, seed_base=rank * n, config=args.config)
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        eng.set_tuning(**{k: int(v)})
    prepare(eng, n, rank)

    obs = torch.zeros(n, 3, 26, device=dev); share = torch.zeros(n, 29, device=dev); rew = torch.zeros(n, 3, device=dev)
    done = torch.zeros(n, dtype=torch.uint8, device=dev); info = torch.zeros(64, n, device=dev); term = torch.zeros(n, 3, 26, device=dev)
    g = torch.Generator(device=dev); g.manual_seed(5678 + rank)
    n_act = 8
    acts = [torch.randint(0, 3, (n, 3), dtype=torch.int32, device=dev, generator=g) for _ in range(n_act)]
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def core_step(i):
        eng.step_device(acts[i % n_act], obs, share, rew, done, None, None, stream.cuda_stream)

    def full_step(i):
        eng.step_device(acts[i % n_act], obs, share, rew, done, info, term, stream.cuda_stream)

    def gather_metrics():
        if world > 1:                                # the path's one collective: the logger's episode-metric vector
            m = torch.tensor(eng.metrics(), device=dev)
            dist.all_gather([torch.zeros_like(m) for _ in range(world)], m)

    # ---- untimed: turn the pre-filled windows over, then warm up the timed call ----
    t_settle = time.perf_counter()
    for i in range(args.settle):
        core_step(i)
    torch.cuda.synchronize(dev)
    t_settle = time.perf_counter() - t_settle
    for i in range(max(args.warmup, 3)):
        full_step(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()

    def timed(step_fn, steps, instrumented=False):
        """EXACTLY `steps` steps between two CUDA events on the launch stream, barrier + synchronize on both sides, max over
        ranks.  instrumented: additionally an event after every step and the library's two events around every launch
        (`sdc_kernel_times`) -- three more event records per step, which cost the stream ~2 % and therefore stay out of the
        pass that measures `value`."""
        eng.set_tuning(clear_pass_total=1)
        if instrumented:
            eng.set_tuning(timing=1)
            eng.kernel_times()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1 if instrumented else 2)]
        barrier()
        ev[0].record(stream)
        for i in range(steps):
            step_fn(i)
            if instrumented:
                ev[i + 1].record(stream)
            if args.config == 4 and (i + 1) % 1024 == 0:
                gather_metrics()
        if not instrumented:
            ev[1].record(stream)
        if sampler:
            sampler.sample_now()          # the launches above are enqueued, the GPU is executing them
        gather_metrics()
        barrier()
        total_ms = ev[0].elapsed_time(ev[-1])
        step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)] if instrumented else None
        kt = eng.kernel_times() if instrumented else None
        if instrumented:
            eng.set_tuning(timing=0)
        passes = eng.read_state("pass_total").astype(np.float64) / steps
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), step_ms, kt, passes

    launches0 = eng.launch_count
    total_ms_max, _, _, passes = timed(full_step, args.steps)
    launches = eng.launch_count - launches0
    value = n * world * args.steps / (total_ms_max / 1e3)
    instr_ms_max, step_ms, ktimes, _ = timed(full_step, args.steps, instrumented=True)      # per-launch kernel times for the roofline
    core_ms_max, _, _, _ = timed(core_step, args.steps)
    value_core = n * world * args.steps / (core_ms_max / 1e3)

    # ---- end to end through the host-buffer calls (numpy in, numpy out; H2D + D2H inside the timed region) ----
    rng = np.random.RandomState(5678 + rank)
    host_acts = [rng.randint(0, 3, size=(n, 3)).astype(np.int32) for _ in range(4)]
    e2e_steps = max(5, min(args.steps, 200))

    def host_timed(fn):
        for i in range(3):
            fn(host_acts[i % 4])
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            fn(host_acts[i % 4])
        barrier()
        s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(s, op=dist.ReduceOp.MAX)
        return n * world * e2e_steps / float(s.item())

    e2e_compact = host_timed(lambda a: eng.step_compact_host(a, want_info=False, want_term=True))
    e2e_padded = host_timed(lambda a: eng.step_host(a, want_info=False, want_term=True))
    from dc_rl_b200.vec_env import CudaShareVecEnv
    vec_args = {"nonoverlapping_shared_obs_space": True}
    vec_views = CudaShareVecEnv(dict(vec_args, output_views=True), n, engine=eng)
    e2e_vec_views = host_timed(lambda a: vec_views.step(a))
    vec_copy = CudaShareVecEnv(vec_args, n, engine=eng)
    e2e_vec_copy = host_timed(lambda a: vec_copy.step(a))
    # ---- what the host link gives all ranks at once: device -> pinned host copies issued concurrently by every rank ----
    link_bytes = 256 << 20
    d_src = torch.empty(link_bytes, dtype=torch.uint8, device=dev)
    h_dst = torch.empty(link_bytes, dtype=torch.uint8, pin_memory=True)
    h_dst.copy_(d_src, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        h_dst.copy_(d_src, non_blocking=True)
    barrier()
    link_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(link_s, op=dist.ReduceOp.MAX)
    link_gbs_rank = 4 * link_bytes / float(link_s.item()) / 1e9          # per rank, with all ranks copying
    del d_src, h_dst
    clocks = sampler.summary() if sampler else None
    err = int(np.bitwise_or.reduce(eng.read_state("err")))

    if rank == 0:
        peak, peak_src = measured_peak()
        k_ms = float(ktimes[1] / max(ktimes[0], 1))          # mean k_step launch duration over the timed region
        achieved = B_ALG_STEADY * n / (k_ms / 1e3) / 1e9
        cap = ncu_capture()
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "k_step", "launch_ms_mean": k_ms, "launch_ms_max": float(ktimes[3]),
                "launch_times_from": "a second pass of the same %d steps with an event after every step and the library's events around every "
                                     "launch (ms_per_step of that pass: %.5f)" % (args.steps, instr_ms_max / args.steps),
                "step_ms_median": float(np.median(step_ms)), "bytes_per_launch": B_ALG_STEADY * n, "peak_source": peak_src,
                "numerator": "SURVEY 8d algorithmic bytes, 4*H + 1024 per env-step (a window pass per step); the kernel maintains the "
                             "normaliser incrementally, so frac > 1 is expected -- `physical` and `limiters` bound it",
                "passes_per_step_mean": {"plain": passes[0], "refresh": passes[1], "by_brackets": passes[2], "by_bands": passes[3],
                                         "share_of_env_steps": float((passes[0] + passes[1]) / n)}}
        if cap and n == cap.get("n_envs"):
            roof["traffic"] = cap["dram_bytes_per_launch"]
            phys = cap["dram_bytes_per_launch"] / (k_ms / 1e3) / 1e9
            roof["physical"] = {"achieved": phys, "frac": phys / peak, "bytes_per_env_step": cap["dram_bytes_per_launch"] / n,
                                "source": "profiles/r02_kstep_ncu.json (%s)" % cap.get("commit", "?")}
            roof["limiters"] = {k: cap[k] for k in ("issue_slot_pct", "fp64_pipe_pct", "achieved_occupancy_pct", "ipc", "ncu_duration_us",
                                                    "warp_instructions", "top_stalls") if k in cap}
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 reward window / f64 scalar physics", "data": "synthetic",
            "config": dict(cfg, settle_steps=args.settle, settle_s=round(t_settle, 2), numa=placement),
            "value_core_outputs": value_core,
            "e2e": {"value": e2e_compact, "unit": "env-steps/s", "h2d_bytes_per_step": n * 3 * 4,
                    "d2h_bytes_per_step": n * (29 + 3) * 4 + n, "steps": e2e_steps,
                    "host_link": {"d2h_gbs_per_rank_all_ranks_copying": link_gbs_rank, "aggregate_gbs": link_gbs_rank * world,
                                  "link_bound_value": n * world / ((n * (29 + 3) * 4 + n + n * 12) / (link_gbs_rank * 1e9)),
                                  "note": "a 256 MB device -> pinned-host copy repeated by every rank at once; link_bound_value = env-steps/s if "
                                          "the step's H2D + D2H bytes moved at that rate and nothing else took time"},
                    "call": "sdc_step_compact_host (numpy actions in; the 29 distinct observation values per env (sdc_expand_obs rebuilds obs[N,3,26] "
                            "and share_obs bit for bit), rewards, dones, terminal rows of finished envs out; pinned host buffers of the handle)"},
            "e2e_padded": {"value": e2e_padded, "unit": "env-steps/s", "d2h_bytes_per_step": n * (78 + 29 + 3) * 4 + n,
                           "call": "sdc_step_host (obs[N,3,26] + share_obs[N,29] + rewards + dones)"},
            "e2e_vec_env": {"value": e2e_vec_views, "copies": e2e_vec_copy, "unit": "env-steps/s",
                            "call": "CudaShareVecEnv.step (harl ShareVecEnv surface: obs, share_obs[N,3,29], rewards, dones, lazy infos, "
                                    "avail); value: output_views=True, copies: fresh arrays like the reference"},
            "gpu_launches": launches, "roofline": roof, "clocks": clocks, "env_error_flags": err,
        }
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_arm(args.cpu_steps)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
