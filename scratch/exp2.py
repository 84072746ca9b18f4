import sys, time
sys.path[:0] = ['/root/repo']
import numpy as np, torch
import bench
n = 65536
eng, _ = bench.build_engine(n, 0)
dev = torch.device("cuda:0")
obs = torch.zeros(n, 3, 26, device=dev); share = torch.zeros(n, 29, device=dev); rew = torch.zeros(n, 3, device=dev)
done = torch.zeros(n, dtype=torch.uint8, device=dev)
acts = [torch.randint(0, 3, (n, 3), dtype=torch.int32, device=dev) for _ in range(8)]
st = torch.cuda.current_stream().cuda_stream
eng.set_tuning(timing=1)
def run(steps, label):
    for i in range(3): eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
    eng.kernel_times()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
    e1.record(); torch.cuda.synchronize()
    kt = eng.kernel_times()
    print("%-44s total %.3f ms/step | k_step %.3f  k_reset %.3f  (max k_step %.3f)" % (label, e0.elapsed_time(e1) / steps, kt[1] / kt[0], kt[2] / kt[0], kt[3]), flush=True)
bench.prepare(eng, n, 0)
run(20, "H=10000 fresh, default")
run(600, "H=10000 600 steps (warm into sustained)")
for unroll in (4, 8, 16):
    for pf in (0, 1):
        eng.set_tuning(unroll=unroll, prefetch=pf)
        run(40, "sustained unroll=%d bulkprefetch=%d" % (unroll, pf))
print("err", int(np.bitwise_or.reduce(eng.read_state("err"))))
