"""Developer tool: a few pair sweeps for ncu: config 3 (N=4096 LJ), a large LJ system (N=17576), config 4's sweep, 8 five-site bead systems."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpmcxx_b200 import engine, workloads as W

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("lj", "all"):
    for s in (W.lj_argon(), W.lj_argon(n_side=26, L=97.5)):
        e = engine.Engine(s)
        for _ in range(2):
            e.energy()
        e.close()
if which in ("es", "all"):
    s4 = W.h2_framework(solver={"polar_max_iter": "1"})
    s4.opts["polarization"] = "off"
    e = engine.Engine(s4)
    for _ in range(2):
        e.energy()
    e.close()
    t, b = W.pi_h2_cluster(P=64, five_site=True)
    e = engine.Engine(t, beads=np.ascontiguousarray(b[:8]))
    for _ in range(2):
        e.energy_all()
    e.close()
