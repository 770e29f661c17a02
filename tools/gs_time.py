import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from mpmcxx_b200 import engine, workloads as W
s = W.h2_framework(solver=W.SOLVER_GS_RANKED_PALMO, ensemble="nvt")
e = engine.Engine(s)
for _ in range(3): e.energy()
e.set_timing(True)
for _ in range(5): e.energy()
tm = e.timing()
print({k: (v[0]/v[1], v[1]) for k, v in tm.items() if v[1]})
