"""CPU tests of the bead-sharding host logic: two gloo ranks, each summing its own beads (the oracle stands in for the
per-bead energies), reproduce the single-process PI_calculate_potential aggregation."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpmcxx_b200 import pi
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port_no, name, q):
    sys.path.insert(0, ROOT)
    from oracle import port
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, r = cases.load_golden(name)
    beads = r["beads"]
    P = beads.shape[0]
    lo, hi = pi.bead_range(P, rank, world)
    o = port.pi_energy(s, beads[lo:hi])          # per-bead energies of the local slice
    local = o["per_bead"].sum(axis=0)

    def allred(a):
        t = torch.from_numpy(a)
        dist.all_reduce(t)

    pot, means = pi.combine_potential(local, P, allred)
    q.put((rank, pot, means.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("name", sorted(cases.PI))
def test_two_rank_sharding_matches_reference(name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    s, r = cases.load_golden(name)
    ref_pot = float(r["ref_pi_potential"])
    assert res[0][1] == res[1][1]                       # every rank sees the same potential
    scale = max(abs(ref_pot), 1.0)
    assert abs(res[0][1] - ref_pot) < 1e-9 * scale
    assert abs(res[0][2][0] - float(r["ref_pi_rd"])) < 1e-10 * abs(float(r["ref_pi_rd"]))


def test_bead_range_and_estimators():
    assert pi.bead_range(64, 3, 8) == (24, 32)
    with pytest.raises(ValueError):
        pi.bead_range(64, 0, 3)
    s, r = cases.load_golden("pi_h2_single_27x8")
    beads = r["beads"]
    P = beads.shape[0]
    total = sum(pi.chain_mass_len2(beads[:, m, :], float(s.mass[m])) for m in range(s.n))   # single-site molecules: COM = site
    assert abs(total - float(r["ref_pi_chain_mass_len2"])) < 1e-13 * total
    kin = pi.kinetic_estimator(total, s.n, P, 20.0)
    assert abs(kin - float(r["ref_pi_kinetic"])) < 1e-10 * abs(kin)
