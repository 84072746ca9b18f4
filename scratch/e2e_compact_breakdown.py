"""e2e (sdc_step_compact_host) per-step wall time vs the k_step launch time inside it, per direct-store mode (diagnostics)."""
import sys, time
sys.path[:0] = ['/root/repo']
import numpy as np, torch
import bench
n = 65536
eng, _ = bench.build_engine(n, 0)
bench.prepare(eng, n, 0)
rng = np.random.RandomState(0)
acts = [rng.randint(0, 3, size=(n, 3)).astype(np.int32) for _ in range(4)]
modes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 7]
for mode in modes:
    eng.set_tuning(direct_host=mode)
    for i in range(20):
        eng.step_compact_host(acts[i % 4], want_info=False, want_term=True)
    eng.set_tuning(timing=1); eng.kernel_times()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(300):
        eng.step_compact_host(acts[i % 4], want_info=False, want_term=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 300
    kt = eng.kernel_times(); eng.set_tuning(timing=0)
    print("direct_host=%2d  step %.3f ms (%.1f M env-steps/s)  k_step %.3f ms  rest %.3f ms" % (mode, dt * 1e3, n / dt / 1e6, kt[1] / kt[0], dt * 1e3 - kt[1] / kt[0]))
