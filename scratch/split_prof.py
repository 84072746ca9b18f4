import sys
sys.path[:0] = ['/root/repo']
import numpy as np, torch
import bench
n = 65536
eng, _ = bench.build_engine(n, 0)
eng.set_tuning(split=int(sys.argv[1]))
bench.prepare(eng, n, 0)
dev = torch.device("cuda:0")
obs = torch.zeros(n, 3, 26, device=dev); share = torch.zeros(n, 29, device=dev); rew = torch.zeros(n, 3, device=dev)
done = torch.zeros(n, dtype=torch.uint8, device=dev)
acts = [torch.randint(0, 3, (n, 3), dtype=torch.int32, device=dev) for _ in range(8)]
st = torch.cuda.current_stream().cuda_stream
for i in range(int(sys.argv[2])):
    eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
torch.cuda.synchronize()
