"""Multi-GPU parity on hardware (skipped on a one-GPU box): the path-integral golden fixtures with their bead systems sharded over
2 (and 8, when present) GPUs, one process per GPU, through both collectives — ncclAllReduce (MPMC_PI_P2P=0) and the peer-memory
exchange fused into the assembly kernel (MPMC_PI_P2P=1): every rank must hold bit-identical sums (the replicated random streams
make their accept/reject decisions from them), equal to the reference's aggregates to 1e-10; and a seeded path-integral trajectory
through the C++ mirror on 2 ranks must reproduce the reference's decisions step by step."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _ngpu():
    try:
        from mpmcxx_b200 import engine
        import ctypes as C
        n = C.c_int()
        engine.lib().mpmc_device_count(C.byref(n))
        return n.value
    except Exception:
        return 0


def _spawn(target, nranks, args, timeout=600):
    from mpmcxx_b200 import engine
    ctx = mp.get_context("spawn")
    uid = engine.nccl_unique_id()
    q = ctx.Queue()
    procs = [ctx.Process(target=target, args=(r, nranks, uid) + args + (q,)) for r in range(nranks)]
    for p in procs:
        p.start()
    res = {}
    try:
        for _ in range(nranks):
            item = q.get(timeout=timeout)
            res[item[0]] = item
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    for r in range(nranks):
        assert res[r][1] == "ok", res[r][2]
    return res


@pytest.mark.parametrize("p2p", [0, 1])
@pytest.mark.parametrize("name,nranks", [("pi_h2_five_8x4", 2), ("pi_h2_single_27x8", 2), ("pi_h2_single_27x8", 8), ("pi_h2_five_8x4", 4)])
def test_bead_sharded_energies_match_reference_on_every_rank(name, nranks, p2p):
    from tests import _multi_worker as W
    if _ngpu() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    s, r = cases.load_golden(name)
    P = r["beads"].shape[0]
    res = _spawn(W.pi_energy_worker, nranks, (name, p2p, 6))
    first = res[0]
    for k in range(1, nranks):
        assert np.array_equal(res[k][2], first[2]), "ranks %d and 0 hold different sums" % k      # bit-identical on every rank
        assert res[k][3] == first[3]
    if p2p:
        assert "peer-memory" in first[4], first[4]
    else:
        assert "ncclAllReduce" in first[4], first[4]
    pot0, rd0, es0 = first[2][0, 0], first[2][0, 1], first[2][0, 2]
    assert abs(rd0 - float(r["ref_pi_rd"])) < RTOL * abs(float(r["ref_pi_rd"]))
    assert abs(first[3] - float(r["ref_pi_chain_mass_len2"])) < 1e-12 * abs(float(r["ref_pi_chain_mass_len2"]))
    # the moved configurations against the oracle (the reference's energies of arbitrary configurations)
    from oracle import port
    q = port.pi_energy(s, first[5])
    # first[5] is the configuration AFTER the last move; the last recorded energy belongs to the one before: recompute both ends
    q0 = port.pi_energy(s, r["beads"])
    scale = abs(q0["rd"]) + abs(q0["coulombic"]) + 1.0
    assert abs(pot0 - q0["potential"]) < 1e-9 * scale
    assert np.isfinite(q["potential"])


def test_sharded_mirror_trajectory_matches_reference(tmp_path):
    """traj_pi_h2_27x8 (P = 8, 27 H2) through SimulationControl::PI_nvt_mc of the C++ mirror on 2 GPUs (mpmc_host_run_sharded:
    mpmc_nccl_init + mpmc_pi_potential_allreduce): the reference's decisions and energies, identical on both ranks."""
    from mpmcxx_b200 import workloads as Wl
    from tests import _multi_worker as W
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    s, r = cases.load_golden_traj("traj_pi_h2_27x8")
    P, ref = int(r["P"]), r["traj"]
    inp = Wl.write_reference_job(s, str(tmp_path))
    res = _spawn(W.pi_trajectory_worker, 2, (inp, P, len(ref)), timeout=900)
    log = res[0][2]
    assert np.array_equal(res[1][2], log)
    assert len(log) == len(ref)
    same = (log[:, 0] == ref[:, 0]) & (log[:, 3] == ref[:, 3])
    assert same.all(), "trajectory diverges at step %d" % int(np.argmin(same))
    assert (np.abs(log[:, 1] - ref[:, 1]) <= RTOL * np.maximum(np.abs(ref[:, 1]), 1.0)).all()
    assert np.allclose(log[:, 4], ref[:, 4], rtol=1e-10, atol=0)
