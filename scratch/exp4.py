import sys, time
sys.path[:0] = ['/root/repo']
import numpy as np, torch
import bench
n = 65536
dev = torch.device("cuda:0")
obs = torch.zeros(n, 3, 26, device=dev); share = torch.zeros(n, 29, device=dev); rew = torch.zeros(n, 3, device=dev)
done = torch.zeros(n, dtype=torch.uint8, device=dev)
acts = [torch.randint(0, 3, (n, 3), dtype=torch.int32, device=dev) for _ in range(8)]
st = torch.cuda.current_stream().cuda_stream
def run(eng, steps, label):
    for i in range(3): eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
    eng.kernel_times(); eng.set_tuning(phases=1)
    torch.cuda.synchronize()
    for i in range(steps): eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
    torch.cuda.synchronize()
    eng.set_tuning(phases=1); eng.step_device(acts[0], obs, share, rew, done, None, None, st); torch.cuda.synchronize()
    mk = eng.read_state("phase_clocks").astype(np.float64)
    print("   one step timeline (us from first CTA start): produce done %.0f | last scan done %.0f | last finish %.0f | last reset worker exit %.0f" % tuple((mk[k] - mk[8]) / 1e3 for k in (9, 10, 11, 12)))
    eng.set_tuning(phases=1)
    for i in range(steps): eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
    torch.cuda.synchronize()
    kt = eng.kernel_times(); ph = eng.read_state("phase_clocks").astype(np.float64)
    units = max(ph[4], 1); ghz = 1.965e3; warps = max(ph[7], 1); jobs = max(ph[6], 1)
    print("%-40s k_step %.3f ms | per unit us: core %.1f publish %.1f obs(deferred) %.1f finish %.1f | per warp us: scanning %.1f idle %.1f | per job us %.2f (jobs/warp %.1f)" % (
        label, kt[1] / kt[0], ph[0] / units / ghz, ph[1] / units / ghz, ph[13] / units / ghz, ph[3] / units / ghz, ph[2] / warps / ghz, ph[5] / warps / ghz,
        ph[2] / jobs / ghz, jobs / warps), flush=True)
eng, _ = bench.build_engine(n, 0)
eng.set_tuning(timing=1)
bench.prepare(eng, n, 0)
for lj in (0, 8):
    eng.set_tuning(unroll=8, prefetch=0, local_jobs=lj)
    run(eng, 30, "fresh unroll=8 local_jobs=%d" % lj)
eng.set_tuning(local_jobs=0)
blob = eng.get_state()
eng.write_state("step_in_ep", np.zeros(n, np.int32)); eng.write_state("t", eng.read_state("t0").astype(np.int32))
eng.set_tuning(unroll=8, prefetch=0)
run(eng, 30, "NO RESETS (synced episodes) unroll=8")
eng.set_state(blob); eng.set_tuning(timing=1)
for i in range(500): eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
for lj in (0, 8):
  for unroll in (4, 8):
    eng.set_tuning(unroll=unroll, prefetch=0, local_jobs=lj)
    run(eng, 30, "sustained unroll=%d local_jobs=%d" % (unroll, lj))
# box calibration: plain copy bandwidth (same recipe as MEASURED_PEAKS.json) and a read-only reduction
a_ = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev); b_ = torch.empty_like(a_)
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b_.copy_(a_); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("copy GB/s (r+w): %.0f" % (2 * a_.numel() * 2 / best / 1e6))
f_ = a_.view(torch.float32)
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f_.sum(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("read-only sum GB/s: %.0f" % (f_.numel() * 4 / best / 1e6))
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.mem,power.draw,temperature.gpu,pstate", "--format=csv,noheader"], capture_output=True, text=True).stdout)
