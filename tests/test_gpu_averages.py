"""Ensemble averages (north_star's third correctness criterion): seeded Markov chains of the host mirror over the GPU engine against
chains of the unmodified reference run with DIFFERENT seeds (tests/golden/make_golden.py averages): the two sets are statistically
independent samples of the same ensemble, so their means must agree within the combined standard error of block means —
energy and N for the grand-canonical pore (src/System.Averages.cpp:8-208: the quantities update_root_averages accumulates), potential
and kinetic energy for a path-integral cluster (the kinetic estimator is a constant minus 1/2 omega^2 sum m <r^2> of the bead chains,
src/SimulationControl.PathIntegral.cpp:810-828, so this is the <r^2> comparison)."""
import numpy as np
import pytest

from mpmcxx_b200 import averages, workloads as W
from tests import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(cases.AVERAGES))
def test_ensemble_averages_agree_with_reference_within_statistical_error(name, tmp_path):
    import os
    from mpmcxx_b200 import host_binding
    build, P, steps, _, our_seeds = cases.AVERAGES[name]
    z = np.load(os.path.join(cases.GOLDEN_DIR, name + ".npz"))
    e_blocks, a_blocks, acc = [], [], []
    for seed in our_seeds:
        s = build()
        s.opts["seed"] = str(seed)
        d = tmp_path / ("seed%d" % seed)
        inp = W.write_reference_job(s, str(d))
        log, summary = host_binding.run(inp, P=P, max_steps=steps, capacity=steps)
        assert len(log) == steps
        ser = averages.chain_series(log, log[0, 1])
        e_blocks.append(averages.block_means(ser["energy"], cases.AVG_BLOCKS))
        a_blocks.append(averages.block_means(ser["aux"], cases.AVG_BLOCKS))
        acc.append(ser["accepted"].mean())
    ce = averages.compare(np.stack(e_blocks), z["energy"])
    ca = averages.compare(np.stack(a_blocks), z["aux"])
    what = ("potential", "kinetic energy") if P else ("energy", "N")
    print("\n%s: <%s> ours %.6g +- %.3g, reference %.6g +- %.3g (z = %.2f); <%s> ours %.6g +- %.3g, reference %.6g +- %.3g (z = %.2f); acceptance ours %.3f reference %.3f"
          % (name, what[0], ce["mean_a"], ce["sem_a"], ce["mean_b"], ce["sem_b"], ce["z"], what[1], ca["mean_a"], ca["sem_a"], ca["mean_b"], ca["sem_b"], ca["z"],
             float(np.mean(acc)), float(z["acceptance"].mean())))
    # 3 sigma of the combined standard error (fixed seeds on both sides: the outcome is deterministic, the threshold is the
    # statistical statement); the acceptance ratios within 0.03
    assert ce["z"] < 3.0, ce
    assert ca["z"] < 3.0, ca
    assert abs(float(np.mean(acc)) - float(z["acceptance"].mean())) < 0.03
