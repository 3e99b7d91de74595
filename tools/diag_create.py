"""Host time of session creation / destruction (what a multilevel solve pays per level)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dotsocp_b200 as dp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 129
nt = (n - 1) // 2 + 1
for rep in range(4):
    t0 = time.perf_counter()
    s = dp.Session("dot2d", nt, n, n)
    t1 = time.perf_counter()
    s.close()
    t2 = time.perf_counter()
    print(f"n={n} rep {rep}: create {1e3*(t1-t0):.1f} ms  destroy {1e3*(t2-t1):.1f} ms  "
          f"[LAYOUT={os.environ.get('DOTSOCP_LAYOUT','pad')} PREP={os.environ.get('DOTSOCP_PREP','1')}]", flush=True)
