"""Developer tool: first-generation (tile) vs second-generation (warp-queue) pair sweep on the BASELINE configs — same inputs,
kernel time from the engine's CUDA-event timers, and the relative difference of every pair-sweep output.  Run on the GPU box."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpmcxx_b200 import engine, workloads as W

FLOP = {"lj": 54.0, "es": 82.0}


def timed(s, beads, v1, reps=20):
    os.environ["MPMC_PAIR_V1"] = "1" if v1 else "0"
    e = engine.Engine(s, beads=beads)
    for _ in range(3):
        outs = e.energy_all()
    e.set_timing(True)
    for _ in range(reps):
        outs = e.energy_all()
    tm = e.timing()
    e.close()
    return outs, tm["pair"][0] / tm["pair"][1]


def main():
    peak, _ = engine.probe_fp64_peak(0)
    print("FP64 peak %.2f TFLOP/s" % peak)
    cases = []
    cases.append(("config3 lj_argon N=4096", W.lj_argon(), None, "lj"))
    s4 = W.h2_framework(solver={"polar_max_iter": "1"})
    s4.opts["polarization"] = "off"
    cases.append(("config4 h2_framework N=10000 (pair sweep only)", s4, None, "es"))
    t, b = W.pi_h2_cluster(P=64, five_site=True)
    cases.append(("config5 pi five-site 64 beads x 2560", t, b, "es"))
    cases.append(("config5 pi five-site 8 beads x 2560", t, np.ascontiguousarray(b[:8]), "es"))
    t1, b1 = W.pi_h2_cluster(P=64, five_site=False)
    cases.append(("config5 pi single-site 64 beads x 512", t1, b1, "lj"))
    cases.append(("config5 pi single-site 8 beads x 512", t1, np.ascontiguousarray(b1[:8]), "lj"))
    for name, s, beads, kind in cases:
        o1, t1_ = timed(s, beads, True)
        o2, t2_ = timed(s, beads, False)
        npairs = sum(o["n_pair_evals"] for o in o2)
        worst = 0.0
        for a, b_ in zip(o1, o2):
            for k in ("rd_pair", "es_real", "es_self_intra"):
                if a[k] != 0.0:
                    worst = max(worst, abs(a[k] - b_[k]) / abs(a[k]))
            assert a["n_pairs_in_cutoff"] == b_["n_pairs_in_cutoff"], (a["n_pairs_in_cutoff"], b_["n_pairs_in_cutoff"])
        f = FLOP[kind]
        print("%-50s v1 %8.1f us (%5.1f%% peak)  v2 %8.1f us (%5.1f%% peak, %6.1f Gpair/s)  x%.2f  max rel diff %.1e" % (
            name, t1_ * 1e3, 100 * f * npairs / (t1_ * 1e-3) / 1e12 / peak, t2_ * 1e3, 100 * f * npairs / (t2_ * 1e-3) / 1e12 / peak,
            npairs / (t2_ * 1e-3) / 1e9, t1_ / t2_, worst))


main()
