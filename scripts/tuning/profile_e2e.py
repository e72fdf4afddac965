"""cProfile of the host side of the e2e step (Pyrat.calc_lbl_extinction) on configs[1]."""
import cProfile, pstats, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pyratbay_b200 import workloads, tli as ptli
from pyratbay_b200.pyrat import Pyrat
w = workloads.forward_model_workload(1_000_000, 81, 0.5, 5.0)
path = "/tmp/pb200_prof.tli"
ptli.write_tli(path, [w.db], [{"wn": w.wn, "elow": w.elow, "gf": w.gf, "iso_id": w.isoid,
               "n_lines_iso": np.bincount(w.isoid, minlength=w.db.niso)}], w.spec.wnlow, w.spec.wnhigh)
pyrat = Pyrat(dict(tlifile=[path], wl_low=w.spec.wl_low, wl_high=w.spec.wl_high, wnstep=1.0,
                   wnosamp=2160, verb=0), atm=w.atm)
for s in range(5):
    pyrat.calc_lbl_extinction(temp=workloads.layer_temperatures(81, realization=s))
n = 200
t0 = time.time()
pr = cProfile.Profile(); pr.enable()
for s in range(n):
    pyrat.calc_lbl_extinction(temp=workloads.layer_temperatures(81, realization=100 + s))
pr.disable()
print("ms per step (wall, under cProfile):", (time.time() - t0) / n * 1e3)
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(18)
print(pyrat.last_timing)
