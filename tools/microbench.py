"""Iteration micro-benchmark on one GPU: ms per inPALM iteration and per kernel group, optionally with check iterations.
    python tools/microbench.py c4 [steps] [kkt_every]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import dotsocp_b200 as dp
wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
kkt_every = int(sys.argv[3]) if len(sys.argv) > 3 else 0
nt, nx, ny = bench.WORKLOADS[wl]
N, L, Q = bench.sizes(nt, nx, ny)
var, model = bench.make_problem(nt, nx, ny)
o = bench.level_opts(var, model, steps)
with dp.Session("dot2d", nt, nx, ny) as s:
    s.upload(var.phi, var.q, var.z, var.alpha, var.beta, model.c)
    s.iter_begin(o)
    s.iterate(3, kkt_every=kkt_every)      # warm-up (also loads the check-iteration kernels)
    ms, pk = s.iterate(steps, per_kernel=True, kkt_every=kkt_every)
    s.iter_end()
peak = bench.measured_peaks()[0]
mult = pk[2] / steps
print(f"{wl} PF={os.environ.get('DOTSOCP_KM_PF','default')} kkt_every={kkt_every}: {ms/steps:.3f} ms/it  poisson {pk[0]/steps:.3f}  qstep {pk[1]/steps:.3f}  "
      f"mult {mult:.3f} ({(N+4*Q+20*L)*8/mult/1e6/peak:.3f} of {peak:.0f} GB/s)  iter frac {(14*N+6*Q+30*L)*8/(ms/steps)/1e6/peak:.3f}", flush=True)
