"""Worker processes of tests/test_gpu_multi.py: one process per GPU (spawned, not forked: each needs its own CUDA context)."""
import os
import traceback

import numpy as np


def pi_energy_worker(rank, nranks, uid, name, p2p, steps, out):
    try:
        os.environ["MPMC_PI_P2P"] = "1" if p2p else "0"
        from mpmcxx_b200 import engine, pi
        from tests import cases
        s, r = cases.load_golden(name)
        beads = r["beads"]
        P = beads.shape[0]
        lo, hi = pi.bead_range(P, rank, nranks)
        e = engine.Engine(s, beads=np.ascontiguousarray(beads[lo:hi]), device=rank)
        e.nccl_init(uid, rank, nranks)
        res = []
        chain = e.pi_chain_allreduce()               # of the fixture's configuration
        rs = np.random.RandomState(17)               # the same moves on every rank
        pos = beads.copy()
        starts = np.nonzero(np.diff(np.concatenate([[-1], s.mol])))[0]
        ends = np.concatenate([starts[1:], [s.n]])
        for it in range(steps):
            pot, means = e.pi_potential_allreduce(P)
            res.append([pot] + list(means))
            m = rs.randint(len(starts))
            a, b = int(starts[m]), int(ends[m])
            pos[:, a:b, :] += rs.normal(scale=0.05, size=(P, 1, 3))
            e.update_sites_all_beads(a, pos[lo:hi, a:b, :])
        coll = e.pi_collective()
        e.close()
        out.put((rank, "ok", np.array(res), chain, coll, pos))
    except Exception:
        out.put((rank, "error", traceback.format_exc(), None, None, None))


def pi_trajectory_worker(rank, nranks, uid, inp, P, steps, out):
    try:
        from mpmcxx_b200 import host_binding
        log, summary = host_binding.run_sharded(inp, P, rank, nranks, rank, uid, max_steps=steps, capacity=steps)
        out.put((rank, "ok", log, summary))
    except Exception:
        out.put((rank, "error", traceback.format_exc(), None))
