"""Diagnose BASELINE configs[2] (weighted 256x256x128) against the committed oracle golden: prints which assertion fails."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util
spec = importlib.util.spec_from_file_location("m", os.path.join(ROOT, "tests", "golden", "make_golden_baseline_configs.py"))
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "solver_baseline_configs.json")))["wdot2d_circle_256x256x128"]
import dotsocp_b200 as dp
from oracle import dotsocp_oracle as O
rho0, rho1, nt, levelN, opts = m.config_wdot2d()
for resident in (True, False):
    o = dict(opts); o["resident"] = resident
    t0 = time.perf_counter()
    try:
        out, _, ML, rh = dp.solver_wdotsocp2d(rho0, rho1, nt, levelN, o, "inPALM")
    except Exception as e:
        import traceback; traceback.print_exc(); continue
    print("resident", resident, "sec", time.perf_counter() - t0, flush=True)
    print(" level_iters", list(out.level_iters), "gold", gold["level_iters"])
    it = [int(v) for v in ML.iter]
    print(" hist_iter same:", it == gold["hist_iter"], len(it), len(gold["hist_iter"]))
    gk = np.array(gold["kkt"])
    n = min(len(it), len(gold["hist_iter"]))
    for i in range(n):
        d = np.abs(ML.kkt[i] - gk[i]).max()
        flag = "" if it[i] == gold["hist_iter"][i] else "  <-- iter differs"
        print(f"  row {i} it {it[i]} gold {gold['hist_iter'][i]} maxdiff {d:.3e}{flag}")
    print(" priVal", rh.priVal[-1], gold["priVal"], abs(rh.priVal[-1] - gold["priVal"]) / abs(gold["priVal"]))
    w2 = O.w2_cost(out, 2)
    print(" w2", w2, gold["w2"], abs(w2 - gold["w2"]) / abs(gold["w2"]), flush=True)
