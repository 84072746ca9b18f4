"""Per-unit phase clocks of one steady-state step (diagnostics): where do the slow units lose their time?"""
import sys
sys.path[:0] = ['/root/repo']
import numpy as np, torch
import bench
n = 65536
settle = int(sys.argv[1]) if len(sys.argv) > 1 else 12000
eng, _ = bench.build_engine(n, 0)
for kv in filter(None, (sys.argv[2] if len(sys.argv) > 2 else "").split(",")):
    k, v = kv.split("="); eng.set_tuning(**{k: int(v)})
bench.prepare(eng, n, 0)
dev = torch.device("cuda:0")
obs = torch.zeros(n, 3, 26, device=dev); share = torch.zeros(n, 29, device=dev); rew = torch.zeros(n, 3, device=dev)
done = torch.zeros(n, dtype=torch.uint8, device=dev)
acts = [torch.randint(0, 3, (n, 3), dtype=torch.int32, device=dev) for _ in range(8)]
st = torch.cuda.current_stream().cuda_stream
for i in range(settle):
    eng.step_device(acts[i % 8], obs, share, rew, done, None, None, st)
torch.cuda.synchronize()
names = ["physics", "normaliser", "rewards", "obs+sums", "finish"]
for rep in range(3):
    eng.set_tuning(phases=1)
    eng.step_device(acts[rep], obs, share, rew, done, None, None, st); torch.cuda.synchronize()
    L = eng.read_state("unit_log")[: n // 32].astype(np.int64)
    tot = L[:, :5].sum(1)
    start = (L[:, 6] - L[:, 6].min()) & 0xffffffff
    sm = L[:, 5] & 0xffff; cta = L[:, 5] >> 16
    ne = L[:, 7] & 0xff; nb = (L[:, 7] >> 8) & 0xff; npass = L[:, 7] >> 16
    print("step %d: units %d  total clk mean %.0f  p50 %.0f p90 %.0f p99 %.0f max %.0f; start offset ns mean %.0f max %.0f" % (
        rep, len(L), tot.mean(), *np.percentile(tot, [50, 90, 99]), tot.max(), start.mean(), start.max()))
    for k, nm in enumerate(names):
        v = L[:, k]
        print("   %-10s mean %7.0f p50 %7.0f p90 %7.0f p99 %7.0f max %7.0f   corr with total %.2f" % (nm, v.mean(), *np.percentile(v, [50, 90, 99]), v.max(), np.corrcoef(v, tot)[0, 1]))
    end_ns = start + tot / 1.965
    print("   end time ns: p50 %.0f p90 %.0f p99 %.0f max %.0f" % (*np.percentile(end_ns, [50, 90, 99]), end_ns.max()))
    for lab, x in (("bracket edits", ne), ("band edits", nb), ("passes asked", npass)):
        print("   %-14s mean %.2f max %d  corr with normaliser clk %.2f" % (lab, x.mean(), x.max(), np.corrcoef(x, L[:, 1])[0, 1]))
        for v in range(0, min(int(x.max()) + 1, 9)):
            m = x == v
            if m.sum(): print("        %d: %5d units, normaliser mean %.0f, total mean %.0f" % (v, m.sum(), L[m, 1].mean(), tot[m].mean()))
    # SMs that host a worker CTA next to a unit CTA vs two unit CTAs
    per_sm = {}
    for s_, c_ in zip(sm, cta): per_sm.setdefault(int(s_), set()).add(int(c_))
    one = np.array([len(per_sm[int(s_)]) == 1 for s_ in sm])
    print("   units on SMs with ONE unit CTA (shares the SM with a worker CTA): %d, total mean %.0f physics %.0f | with two unit CTAs: %d, total mean %.0f physics %.0f" % (
        one.sum(), tot[one].mean() if one.any() else 0, L[one, 0].mean() if one.any() else 0, (~one).sum(), tot[~one].mean(), L[~one, 0].mean()))
    worst = np.argsort(-tot)[:8]
    for u in worst: print("   slow unit %4d sm %3d cta %3d: %s edits %d/%d/%d start %d" % (u, sm[u], cta[u], L[u, :5].tolist(), ne[u], nb[u], npass[u], start[u]))
# phases of the window passes of the last step (upper half of the log: one row per pass ticket)
Lp = eng.read_state("unit_log")
half = len(Lp) // 2
P = Lp[half:half + 4096].astype(np.int64)
P = P[P[:, 4] > 0]
if len(P):
    print("passes logged %d: total clk mean %.0f | TMA wait %.0f  scan %.0f  classify+reduce %.0f  band sort + rank sorts %.0f  commit %.0f | collected c0 mean %.0f c1 mean %.0f" % (
        len(P), P[:, 4].mean(), P[:, 0].mean(), P[:, 1].mean(), P[:, 2].mean(), P[:, 3].mean(), (P[:, 4] - P[:, :4].sum(1)).mean(),
        (P[:, 5] & 0xffff)[(P[:, 5] & 0xffff) > 0].mean() if ((P[:, 5] & 0xffff) > 0).any() else 0, (P[:, 5] >> 16)[(P[:, 5] >> 16) > 0].mean() if ((P[:, 5] >> 16) > 0).any() else 0))
    kinds = P[:, 6]
    for lab, m in (("re-centre list 0 only", ((kinds >> 16) & 3) == 1), ("list 1 only", ((kinds >> 16) & 3) == 2), ("both lists", ((kinds >> 16) & 3) == 3), ("no list (bands only)", ((kinds >> 16) & 3) == 0)):
        if m.any(): print("   %-22s %4d passes, total %.0f, sorts %.0f, tails rebuilt in %d" % (lab, m.sum(), P[m, 4].mean(), P[m, 3].mean(), (((kinds[m] >> 8) & 0xff) > 0).sum()))
if len(P):
    lat = P[:, 7].astype(np.float64)                      # ns from publication to the end of the pass
    run_ns = P[:, 4] / 1.965
    wait = lat - run_ns
    print("pass queueing: publication -> pass start ns  mean %.0f  p50 %.0f  p90 %.0f  max %.0f   (pass itself mean %.0f ns)" % (
        wait.mean(), *np.percentile(wait, [50, 90]), wait.max(), run_ns.mean()))
