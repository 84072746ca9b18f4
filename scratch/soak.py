"""Soak: many steps of generated-mode envs (device RNG resets), then the incremental normaliser state of a sample of envs
is checked against their windows (tests/helpers.check_incremental_state) and one step is priced against numpy."""
import sys
sys.path[:0] = ['/root/repo', '/root/repo/tests']
import numpy as np, torch
import bench
from helpers import check_incremental_state
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
eng, _ = bench.build_engine(n, 0)
bench.prepare(eng, n, 0)
dev = torch.device("cuda:0")
obs = torch.zeros(n, 3, 26, device=dev); share = torch.zeros(n, 29, device=dev); rew = torch.zeros(n, 3, device=dev)
done = torch.zeros(n, dtype=torch.uint8, device=dev); info = torch.zeros(64, n, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1)
acts = [torch.randint(0, 3, (n, 3), dtype=torch.int32, device=dev, generator=g) for _ in range(16)]
st = torch.cuda.current_stream().cuda_stream
passes = np.zeros(4, np.int64)
for i in range(steps):
    eng.step_device(acts[i % 16], obs, share, rew, done, info, None, st)
    if i % 1000 == 999:
        passes += eng.read_state("pass_stats")
torch.cuda.synchronize()
err = int(np.bitwise_or.reduce(eng.read_state("err")))
sample = list(range(0, n, max(1, n // 64)))
valid, checked = check_incremental_state(eng, envs=sample, tag="soak")
print("soak: %d envs x %d steps, err flags %d, incremental state exact for %d sampled envs (%d with valid bands), "
      "passes at the sampled steps %s, episodes %d" % (n, steps, err, checked, valid, passes.tolist(), int(eng.metrics()[11])))
assert err == 0 and torch.isfinite(obs).all() and torch.isfinite(rew).all()
