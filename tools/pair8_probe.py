"""Developer tool (GPU box): pair sweep of the path-integral five-site config at 8 beads per GPU (the 8-GPU shard) under MPMC_PAIR_ROUNDS."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpmcxx_b200 import engine, workloads as W
t, b = W.pi_h2_cluster(P=8, five_site=True)
e = engine.Engine(t, beads=b)
for _ in range(3): e.energy_all()
e.set_timing(True)
for _ in range(30): e.energy_all()
tm = e.timing(); print(os.environ.get("MPMC_PAIR_ROUNDS"), "pair us", 1e3 * tm["pair"][0] / tm["pair"][1])
