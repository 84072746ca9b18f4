"""e2e (sdc_step_host) per-step time vs the k_step launch time inside it, for the direct-store modes."""
import sys, time
sys.path[:0] = ['/root/repo']
import numpy as np, torch
import bench
n = 65536
eng, _ = bench.build_engine(n, 0)
bench.prepare(eng, n, 0)
rng = np.random.RandomState(0)
acts = [rng.randint(0, 3, size=(n, 3)).astype(np.int32) for _ in range(4)]
for mode in (0, 1, 2, 3):
    eng.set_tuning(direct_host=mode)
    for i in range(5):
        eng.step_host(acts[i % 4], want_info=False, want_term=False)
    eng.set_tuning(timing=1); eng.kernel_times()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(100):
        eng.step_host(acts[i % 4], want_info=False, want_term=False)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 100
    kt = eng.kernel_times(); eng.set_tuning(timing=0)
    print("direct_host=%d  step %.3f ms  k_step %.3f ms  rest %.3f ms" % (mode, dt * 1e3, kt[1] / kt[0], dt * 1e3 - kt[1] / kt[0]))
# raw D2H speed of the pinned buffers
b = eng._host_buffers()
d_obs = torch.zeros(n, 3, 26, device="cuda:0")
h = torch.from_numpy(b["obs"])
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    h.copy_(d_obs, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print("D2H of obs (%.1f MB) into the pinned buffer: %.3f ms = %.1f GB/s" % (h.numel() * 4 / 1e6, dt * 1e3, h.numel() * 4 / dt / 1e9))
