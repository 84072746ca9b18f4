"""Step time and window-pass rate of the bench workload as the reward windows turn over (diagnostics):
prefill N(330, 40), then `total` device-resident steps; per block of `blk` steps the mean step time (CUDA events) and
the pass counters of the block's last step.  Usage: steady.py [total] [blk] [tune k=v,...]"""
import sys
sys.path[:0] = ['/root/repo']
import numpy as np, torch
import bench
n = 65536
total = int(sys.argv[1]) if len(sys.argv) > 1 else 14000
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
eng, _ = bench.build_engine(n, 0)
for kv in filter(None, (sys.argv[3] if len(sys.argv) > 3 else "").split(",")):
    k, v = kv.split("="); eng.set_tuning(**{k: int(v)})
bench.prepare(eng, n, 0)
dev = torch.device("cuda:0")
obs = torch.zeros(n, 3, 26, device=dev); share = torch.zeros(n, 29, device=dev); rew = torch.zeros(n, 3, device=dev)
done = torch.zeros(n, dtype=torch.uint8, device=dev)
acts = [torch.randint(0, 3, (n, 3), dtype=torch.int32, device=dev) for _ in range(8)]
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for b in range(total // blk):
    torch.cuda.synchronize(); e0.record()
    for i in range(blk):
        eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
    e1.record(); torch.cuda.synchronize()
    ps = eng.read_state("pass_stats")
    print("steps %6d..%6d  mean %.4f ms  (%.1f M env-steps/s)  passes last step: plain %d refresh %d (lists %d tails %d)" % (
        b * blk, (b + 1) * blk, e0.elapsed_time(e1) / blk, n * blk / e0.elapsed_time(e1) / 1e3, ps[0], ps[1], ps[2], ps[3]), flush=True)
# phase clocks over 100 steps at the end
eng.set_tuning(phases=1)
for i in range(100):
    eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
torch.cuda.synchronize()
pc = eng.read_state("phase_clocks").astype(np.float64)
units = max(pc[4], 1)
print("per-unit clocks: physics %.0f  normaliser %.0f  rewards(incr) %.0f  obs+sums %.0f  wait+passes %.0f  finish %.0f (units %d)" % (
    pc[0] / units, pc[1] / units, pc[8] / units, pc[9] / units, pc[2] / units, pc[3] / units, units))
cnt = int(pc[12]); nres, npass, ngen = cnt & 0xffff, (cnt >> 16) & 0xffff, (cnt >> 32) & 0xffff
print("worker jobs over 100 steps (counts mod 65536): %d resets x %.0f clk, %d passes x %.0f clk, %d generations x %.0f clk" % (
    nres, pc[5] / max(nres, 1), npass, pc[7] / max(npass, 1), ngen, pc[6] / max(ngen, 1)))
print("err", int(np.bitwise_or.reduce(eng.read_state("err"))))
# timeline of single steps in the steady state (phase clocks cleared before each)
for k in range(5):
    eng.set_tuning(phases=1)
    torch.cuda.synchronize(); e0.record(); eng.step_device(acts[k], obs, share, rew, done, None, None, st); e1.record(); torch.cuda.synchronize()
    p1 = eng.read_state("phase_clocks").astype(np.float64)
    cnt = int(p1[12])
    print("single step: event %.1f us; first CTA start -> last unit done %.1f us -> last CTA done %.1f us; slowest unit %.0f clk (before barrier %.0f); jobs: %d passes x %.0f clk, %d generations x %.0f clk" % (
        e0.elapsed_time(e1) * 1e3, (p1[14] - p1[13]) / 1e3, (p1[15] - p1[13]) / 1e3, p1[10], p1[11],
        (cnt >> 16) & 0xffff, p1[7] / max((cnt >> 16) & 0xffff, 1), (cnt >> 32) & 0xffff, p1[6] / max((cnt >> 32) & 0xffff, 1)))
