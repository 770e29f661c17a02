import sys; sys.path.insert(0,'/root/repo')
import numpy as np, tempfile
from mpmcxx_b200 import workloads as W, host_binding
from tests import cases
for name in ("traj_pi_argon_dimer", "traj_pi_h2_27x8"):
    s, r = cases.load_golden_traj(name)
    P, ref = int(r["P"]), r["traj"]
    d = tempfile.mkdtemp()
    inp = W.write_reference_job(s, d)
    log, summary = host_binding.run(inp, P=P, max_steps=12, capacity=12)
    np.set_printoptions(precision=10, linewidth=200)
    print(name)
    for i in range(12):
        print(i, "ours", log[i], "\n   ref ", ref[i])
