"""Is one iteration a pure function of the uploaded state?  Upload the same random state several times, run a few iterations
(no checks, or every k-th a fused check), download, compare the results bit for bit between repeats and between the aligned
and the haloed k_mult.    python tools/diag_determinism.py c4 [iters] [kkt_every]"""
import os, sys, hashlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import dotsocp_b200 as dp
wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kkt_every = int(sys.argv[3]) if len(sys.argv) > 3 else 0
nt, nx, ny = bench.WORKLOADS[wl]
N, L, Q = bench.sizes(nt, nx, ny)
var, model = bench.make_problem(nt, nx, ny)
rng = np.random.default_rng(7)
phi = rng.standard_normal(N); q = 0.1 * rng.standard_normal(Q); alpha = 0.1 * rng.standard_normal(Q)
beta = np.asfortranarray(0.1 * rng.standard_normal((L, 10)))
o = bench.level_opts(var, model, iters)
def run(al):
    os.environ["DOTSOCP_KM_AL"] = al
    with dp.Session("dot2d", nt, nx, ny) as s:
        s.upload(phi, q, None, alpha, beta, model.c)
        s.iter_begin(o)
        s.iterate(iters, kkt_every=kkt_every)
        s.iter_end()
        out = s.download()
    return out
names = ["phi", "q", "z", "alpha", "beta"]
ref = None
for rep, al in enumerate(["0", "1", "1", "1", "0"]):
    out = run(al)
    if ref is None:
        ref = out
        print(f"{wl} iters={iters} kkt_every={kkt_every}: reference AL=0 done", flush=True)
        continue
    msg = []
    for nme, a, b in zip(names, out, ref):
        same = np.array_equal(a, b, equal_nan=True)
        if not same:
            d = np.abs(a - b)
            idx = np.unravel_index(np.nanargmax(d), d.shape)
            msg.append(f"{nme}: DIFF max {np.nanmax(d):.3e} at {idx} count {(d > 0).sum()}")
    print(f"  run {rep} AL={al}: " + ("bit-identical" if not msg else "; ".join(msg)), flush=True)
