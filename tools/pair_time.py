"""Developer tool: pair-sweep kernel time on the BASELINE configs (engine's CUDA-event timers).  Run on the GPU box."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpmcxx_b200 import engine, workloads as W

FLOP = {"lj": 54.0, "es": 82.0}


def timed(s, beads, reps=20):
    e = engine.Engine(s, beads=beads)
    for _ in range(3):
        outs = e.energy_all()
    e.set_timing(True)
    for _ in range(reps):
        outs = e.energy_all()
    tm = e.timing()
    e.close()
    return outs, tm["pair"][0] / tm["pair"][1]


peak, _ = engine.probe_fp64_peak(0)
print("FP64 peak %.2f TFLOP/s   (the percentages count the REFERENCE's flops per pair it visits: 54 LJ, 82 LJ + Coulomb; the per-kind\n loops of a five-site model issue about half of that, so the figure can exceed what the FP64 pipe itself shows in ncu)" % peak)
s4 = W.h2_framework(solver={"polar_max_iter": "1"})
s4.opts["polarization"] = "off"
t, b = W.pi_h2_cluster(P=64, five_site=True)
t1, b1 = W.pi_h2_cluster(P=64, five_site=False)
for name, s, beads, kind in (("config3 lj_argon N=4096", W.lj_argon(), None, "lj"), ("config4 h2_framework N=10000 (pair sweep only)", s4, None, "es"),
                             ("config5 pi five-site 64 beads x 2560", t, b, "es"), ("config5 pi five-site 8 beads x 2560", t, np.ascontiguousarray(b[:8]), "es"),
                             ("config5 pi single-site 64 beads x 512", t1, b1, "lj")):
    o, ms = timed(s, beads)
    npairs = sum(x["n_pair_evals"] for x in o)
    print("%-50s %8.1f us  %6.1f Gpair/s  %5.1f %% of FP64 peak (algorithmic %g flop/pair)" % (name, ms * 1e3, npairs / (ms * 1e-3) / 1e9,
                                                                                            100 * FLOP[kind] * npairs / (ms * 1e-3) / 1e12 / peak, FLOP[kind]))
