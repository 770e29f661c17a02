"""Developer tool (GPU box): config 4 with the benchmarked solver — parity of the polarization solve against the oracle, per-class
kernel times, and the whole energy() between CUDA events.  `python tools/gs_check.py [noparity]`"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mpmcxx_b200 import engine, workloads as W

s = W.h2_framework(solver=W.SOLVER_GS_RANKED_PALMO, ensemble="nvt")
e = engine.Engine(s)
o = e.energy()
if "noparity" not in sys.argv:
    from oracle import port
    p = port.energy(s, want_sites=True)
    d = e.dipoles()
    print("polar gpu %.15e oracle %.15e rel %.2e ; mu err %.2e ; efic err %.2e" % (
        o["polarization_energy"], p["polar"], abs(o["polarization_energy"] - p["polar"]) / abs(p["polar"]),
        np.abs(d["mu"] - p["mu"]).max() / np.abs(p["mu"]).max(),
        np.abs(d["ef_induced_change"] - p["ef_induced_change"]).max() / np.abs(p["ef_induced"]).max()))
for _ in range(3):
    e.energy()
ext = torch.cuda.ExternalStream(e.stream())
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext); e.enqueue(); e1.record(ext); e.fetch(); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
print("energy() ms: min %.3f median %.3f" % (min(ts), float(np.median(ts))))
e.set_timing(True)
for _ in range(5):
    e.energy()
tm = e.timing()
print({k: (round(v[0] / 5, 4), v[1] // 5) for k, v in tm.items() if v[1]})
