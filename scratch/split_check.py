"""Fused k_step vs the split-phase variant (k_phys -> k_obs -> k_step<false>): same inputs, outputs must be bit-identical;
then timing of both on the bench workload."""
import sys
sys.path[:0] = ['/root/repo']
import numpy as np, torch
import bench
dev = torch.device("cuda:0")
def run(n, steps, split, record):
    eng, _ = bench.build_engine(n, 0)
    eng.set_tuning(split=split)
    bench.prepare(eng, n, 0)
    obs = torch.zeros(n, 3, 26, device=dev); share = torch.zeros(n, 29, device=dev); rew = torch.zeros(n, 3, device=dev)
    done = torch.zeros(n, dtype=torch.uint8, device=dev); info = torch.zeros(64, n, device=dev); term = torch.zeros(n, 3, 26, device=dev)
    g = torch.Generator(device=dev); g.manual_seed(7)
    acts = [torch.randint(0, 3, (n, 3), dtype=torch.int32, device=dev, generator=g) for _ in range(8)]
    st = torch.cuda.current_stream().cuda_stream
    out = []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    for i in range(steps):
        ev[i].record()
        eng.step_device(acts[i % 8], obs, share, rew, done, info if record else None, term if record else None, st)
        if record and (i % 7 == 0 or i > steps - 4):
            out.append([x.clone() for x in (obs, share, rew, done, info)])
    ev[steps].record(); torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return out, float(np.median(ms[30:])), eng.metrics(), eng.hvac_histogram()[0], int(np.bitwise_or.reduce(eng.read_state("err")))
a, _, ma, ha, ea = run(8192, 300, 0, True)
b, _, mb, hb, eb = run(8192, 300, 1, True)
same = all(torch.equal(x, y) for sa, sb in zip(a, b) for x, y in zip(sa, sb))
print("outputs identical:", same, " metrics identical:", np.array_equal(ma, mb) or float(np.max(np.abs(ma - mb) / np.maximum(1, np.abs(ma)))),
      " histogram identical:", np.array_equal(ha, hb), " err", ea, eb)
for split in (0, 1, 0, 1):
    _, med, _, _, _ = run(65536, 300, split, False)
    print("split=%d  median step %.4f ms" % (split, med))
