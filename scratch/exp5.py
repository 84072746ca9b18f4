import sys, time
sys.path[:0] = ['/root/repo']
import numpy as np, torch
import bench
from dc_rl_b200 import _lib
n = 65536
dev = torch.device("cuda:0")
obs = torch.zeros(n, 3, 26, device=dev); share = torch.zeros(n, 29, device=dev); rew = torch.zeros(n, 3, device=dev)
done = torch.zeros(n, dtype=torch.uint8, device=dev)
acts = [torch.randint(0, 3, (n, 3), dtype=torch.int32, device=dev) for _ in range(8)]
st = torch.cuda.current_stream().cuda_stream
def run(eng, steps, label):
    for i in range(3): eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
    eng.kernel_times()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
    e1.record(); torch.cuda.synchronize()
    kt = eng.kernel_times()
    print("%-40s step %.3f ms | k_step %.3f k_reset %.3f" % (label, e0.elapsed_time(e1) / steps, kt[1] / kt[0], kt[2] / kt[0]), flush=True)
import bench as B
from dc_rl_b200 import engine as E
def run2(eng, steps, label):
    for i in range(3): eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
    eng.kernel_times(); eng.set_tuning(phases=1)
    torch.cuda.synchronize()
    for i in range(steps): eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
    torch.cuda.synchronize()
    kt = eng.kernel_times(); ph = eng.read_state("phase_clocks").astype(np.float64)
    units = max(ph[4], 1); ghz = 1.965e3
    print("%-34s k_step %.3f ms | per-unit us: physics %.1f stage %.1f scans %.1f finish %.1f | flush %.1f lists %.1f prepare %.1f" % (
        label, kt[1] / kt[0], ph[0] / units / ghz, ph[1] / units / ghz, ph[2] / units / ghz, ph[3] / units / ghz,
        ph[5] / units / ghz, ph[6] / units / ghz, ph[7] / units / ghz), flush=True)
eng, _ = B.build_engine(n, 0)
eng.set_tuning(timing=1)
B.prepare(eng, n, 0)
for label, pf in (("baseline", 0), ("delay publish 50us", 8), ("smem +11KB", 11 << 8), ("smem +20KB", 20 << 8), ("both", 8 | (11 << 8))):
    eng.set_tuning(unroll=8, prefetch=pf)
    run(eng, 30, label)
