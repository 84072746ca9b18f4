"""Per-phase clock sums of k_step and refresh statistics on the bench workload (diagnostics)."""
import sys
sys.path[:0] = ['/root/repo']
import numpy as np, torch
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
eng, _ = bench.build_engine(n, 0)
for kv in filter(None, (sys.argv[3] if len(sys.argv) > 3 else "").split(",")):
    k, v = kv.split("="); eng.set_tuning(**{k: int(v)})
bench.prepare(eng, n, 0)
dev = torch.device("cuda:0")
obs = torch.zeros(n, 3, 26, device=dev); share = torch.zeros(n, 29, device=dev); rew = torch.zeros(n, 3, device=dev)
done = torch.zeros(n, dtype=torch.uint8, device=dev)
acts = [torch.randint(0, 3, (n, 3), dtype=torch.int32, device=dev) for _ in range(8)]
st = torch.cuda.current_stream().cuda_stream
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
scans = []
for i in range(steps):
    if i == 20:
        eng.set_tuning(phases=1)
    ev[i].record()
    eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
    if i < 12 or i % 25 == 0:
        scans.append((i,) + tuple(int(x) for x in eng.read_state("pass_stats")))
ev[steps].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
print("step ms: first", ["%.3f" % x for x in ms[:6]], "median %.4f" % float(np.median(ms[30:])), "min %.4f" % min(ms))
print("scans (step, plain, refresh, by lists, by tails):", scans)
pc = eng.read_state("phase_clocks").astype(np.float64)
units = max(pc[4], 1)
print("per-unit clocks: physics %.0f  normaliser %.0f  passes %.0f  finish+obs %.0f  (units %d)" % (pc[0] / units, pc[1] / units, pc[2] / units, pc[3] / units, units))
print("reward_finish %.0f emit_obs %.0f per unit" % (pc[8] / units, pc[9] / units))
# one isolated step with the timeline stamps
eng.set_tuning(phases=1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); eng.step_device(acts[0], obs, share, rew, done, None, None, st); e1.record(); torch.cuda.synchronize()
p1 = eng.read_state("phase_clocks").astype(np.float64)
print("single step: event %.1f us; first CTA start -> last unit done %.1f us -> last CTA done %.1f us; slowest unit %.0f clocks (scalar phase %.0f); passes %s" % (
    e0.elapsed_time(e1) * 1e3, (p1[14] - p1[13]) / 1e3, (p1[15] - p1[13]) / 1e3, p1[10], p1[11], eng.read_state("pass_stats")))
cnt = int(p1[12]); nres, npass, ngen = cnt & 0xffff, (cnt >> 16) & 0xffff, (cnt >> 32) & 0xffff
print("worker jobs in that step: %d resets x %.0f clk, %d passes x %.0f clk, %d generations x %.0f clk" % (
    nres, p1[5] / max(nres, 1), npass, p1[7] / max(npass, 1), ngen, p1[6] / max(ngen, 1)))
print("err", int(np.bitwise_or.reduce(eng.read_state("err"))), "valid tails", int((eng.read_state("tail_n").reshape(n, 2)[:, 0] >= 0).sum()))
fc = eng.read_state("fast_cfg")
print("alpha exp hist", np.bincount(fc & 0xff)[:14], "retry>0", int(((fc >> 8) & 0xff > 0).sum()))
tn = eng.read_state("tail_n").reshape(n, 2)
print("band sizes mean", tn.mean(0), "max", tn.max(0))
