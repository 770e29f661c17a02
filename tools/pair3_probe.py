"""Developer tool (GPU box): pair sweep of config 3 (N = 4096 LJ) under the item-size knob MPMC_PAIR_ROUNDS."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpmcxx_b200 import engine, workloads as W
s = W.lj_argon()
e = engine.Engine(s)
for _ in range(5): e.energy()
e.set_timing(True)
for _ in range(50): e.energy()
tm = e.timing()
us = 1e3 * tm["pair"][0] / tm["pair"][1]
peak, _ = engine.probe_fp64_peak(0)
print(os.environ.get("MPMC_PAIR_ROUNDS"), "pair us %.2f  -> %.1f %% of FP64 peak (54 flop/pair)" % (us, 100 * 54.0 * 8386560 / (us * 1e-6) / 1e12 / peak))
