import sys, os
sys.path[:0] = ['/root/repo', '/root/repo/tests', '/root/repo/oracle']
import numpy as np
import hostsim_build
from dc_rl_b200 import _lib, info_layout
from replay import make_engine, stage
from helpers import load_traj
np.set_printoptions(linewidth=200, precision=6)
for name, N in (("ny_m0_s0", 1), ("wa_m9_s3", 64)):
    g = load_traj(name)
    engs = [make_engine(g, hostsim_build.load(), N), make_engine(g, _lib.load(), N)]
    ids = np.arange(N, dtype=np.int32)
    for e in engs:
        stage(e, g, 0, ids); e.reset_host()
    for s in range(60):
        act = np.broadcast_to(g["actions"][s].astype(np.int32), (N, 3))
        out = [tuple(np.array(x, copy=True) for x in e.step_host(act)) for e in engs]
        bad = False
        for nm, i in (("obs", 0), ("share", 1), ("rew", 2), ("done", 3), ("info", 4)):
            a, b = out[0][i].astype(np.float64), out[1][i].astype(np.float64)
            err = np.abs(a - b) / np.maximum(1, np.abs(a))
            if err.max() > 1e-5:
                bad = True
                idx = np.argwhere(err > 1e-5)
                print(name, "step", s, nm, "max err", err.max(), "n bad", len(idx), "first", idx[:6].tolist())
                for j in idx[:6]:
                    print("   ", tuple(j), "host", a[tuple(j)], "cuda", b[tuple(j)], (info_layout.INFO_COLUMNS[j[0]] if nm == "info" and j[0] < 59 else ""))
        if bad:
            for st in ("t", "step_in_ep", "setpoint", "bat_load", "ls_len", "ls_head", "ls_sum", "dc_run", "dc_scale", "dc_last", "hist_len", "hist_head", "q_a", "q_m", "err"):
                x, y = engs[0].read_state(st), engs[1].read_state(st)
                if not np.array_equal(x, y):
                    print("   state", st, "host", np.ravel(x)[:8], "cuda", np.ravel(y)[:8])
            break
    else:
        print(name, "no divergence in 60 steps")
