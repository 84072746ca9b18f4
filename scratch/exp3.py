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
    eng.kernel_times()
    torch.cuda.synchronize()
    for i in range(steps): eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
    torch.cuda.synchronize()
    kt = eng.kernel_times()
    print("%-60s k_step %.3f ms (max %.3f) resets/step %.1f" % (label, kt[1] / kt[0], kt[3], done.sum().item()), flush=True)
eng, _ = bench.build_engine(n, 0)
eng.set_tuning(timing=1)
bench.prepare(eng, n, 0)
blob = eng.get_state()
for sync in (0, 1):
    eng.set_state(blob)
    if sync:
        eng.write_state("step_in_ep", np.zeros(n, np.int32)); eng.write_state("t", eng.read_state("t0").astype(np.int32))
    for pf in (0, 2):
        for unroll in (4, 8):
            eng.set_tuning(prefetch=pf, unroll=unroll)
            run(eng, 30, "fresh sync=%d (no resets if 1) headprefetch=%d unroll=%d" % (sync, pf // 2, unroll))
