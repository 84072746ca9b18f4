import sys, time, subprocess, threading
sys.path[:0] = ['/root/repo']
import numpy as np, torch
import bench
n = 65536
eng, _ = bench.build_engine(n, 0)
dev = torch.device("cuda:0")
obs = torch.zeros(n, 3, 26, device=dev); share = torch.zeros(n, 29, device=dev); rew = torch.zeros(n, 3, device=dev)
done = torch.zeros(n, dtype=torch.uint8, device=dev); info = torch.zeros(64, n, device=dev)
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
    print("%-40s total %.3f ms/step | k_step %.3f  k_reset %.3f  (max k_step %.3f) over %d steps" % (label, e0.elapsed_time(e1) / steps, kt[1] / kt[0], kt[2] / kt[0], kt[3], kt[0]), flush=True)
# physics only: empty history (reset all, H grows from 0)
eng.reset_host()
run(20, "H~20 (physics only), synced episodes")
bench.prepare(eng, n, 0)
run(20, "H=10000, desync, 20 steps")
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "100"], stdout=subprocess.PIPE, text=True)
run(1500, "H=10000, desync, 1500 steps")
smi.terminate()
out = smi.stdout.read().strip().split("\n")
print("clock samples:", len(out)); print("\n".join(out[::max(1, len(out)//12)]))
run(20, "H=10000 again 20 steps")
