import sys, time, itertools, json
sys.path[:0] = ['/root/repo']
import numpy as np, torch
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
eng, _ = bench.build_engine(n, 0)
bench.prepare(eng, n, 0)
dev = torch.device("cuda:0")
obs = torch.zeros(n, 3, 26, device=dev); share = torch.zeros(n, 29, device=dev); rew = torch.zeros(n, 3, device=dev)
done = torch.zeros(n, dtype=torch.uint8, device=dev); info = torch.zeros(64, n, device=dev)
acts = [torch.randint(0, 3, (n, 3), dtype=torch.int32, device=dev) for _ in range(8)]
st = torch.cuda.current_stream().cuda_stream
def run(steps, with_info=False):
    for i in range(3): eng.step_device(acts[i % 8], obs, share, rew, done, info if with_info else None, None, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): eng.step_device(acts[i % 8], obs, share, rew, done, info if with_info else None, None, st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
res = []
for unit, unroll, pf, bps in itertools.product((32, 16, 8), (4, 8, 16), (0, 1), (1, 2)):
    eng.set_tuning(unit_envs=unit, unroll=unroll, prefetch=pf, blocks_per_sm=bps)
    ms = run(15)
    res.append((ms, unit, unroll, pf, bps))
    print("unit %2d unroll %2d prefetch %d bps %d : %.3f ms/step  %.1f M steps/s  %.0f GB/s alg" % (unit, unroll, pf, bps, ms, n / ms / 1e3, 41024 * n / ms / 1e6), flush=True)
res.sort()
print("best", res[:5])
ms, unit, unroll, pf, bps = res[0]
eng.set_tuning(unit_envs=unit, unroll=unroll, prefetch=pf, blocks_per_sm=bps)
print("with info table: %.3f ms" % run(15, True))
print("err flags", int(np.bitwise_or.reduce(eng.read_state("err"))))
