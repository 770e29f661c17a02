"""Developer tool: a few pair sweeps for ncu (second-generation kernel): config 3 and 8 five-site bead systems."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpmcxx_b200 import engine, workloads as W

which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("lj", "both"):
    e = engine.Engine(W.lj_argon())
    for _ in range(2):
        e.energy()
    e.close()
if which in ("es", "both"):
    t, b = W.pi_h2_cluster(P=64, five_site=True)
    e = engine.Engine(t, beads=np.ascontiguousarray(b[:8]))
    for _ in range(2):
        e.energy_all()
    e.close()
