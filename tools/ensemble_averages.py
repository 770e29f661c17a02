"""Developer tool: ensemble averages of N seeded replicas of a job through the C++ host mirror (one replica per GPU when several are
visible, else one after the other), accumulated the way the reference does (src/System.Averages.cpp:8-208; chains merged as
src/System.MonteCarlo.cpp:1973-2022 merges MPI ranks), optionally next to chains of the unmodified reference (oracle/_ref, build
container only).   python tools/ensemble_averages.py <case from tests.cases.AVERAGES> [--reference] [--seeds 1 2 3 4]"""
import argparse
import multiprocessing as mp
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from mpmcxx_b200 import averages, workloads as W
from tests import cases


def _ours(args):
    name, seed, device = args
    os.environ["CUDA_VISIBLE_DEVICES"] = str(device)
    from mpmcxx_b200 import host_binding
    build, P, steps, _, _ = cases.AVERAGES[name]
    s = build()
    s.opts["seed"] = str(seed)
    inp = W.write_reference_job(s, tempfile.mkdtemp(prefix="avg_"))
    log, _ = host_binding.run(inp, P=P, max_steps=steps, capacity=steps)
    return averages.chain_series(log, log[0, 1])


def _ref(args):
    name, seed = args
    from oracle import ref
    build, P, steps, _, _ = cases.AVERAGES[name]
    s = build()
    s.opts["seed"] = str(seed)
    r = ref.RefSystem(s, P=P)
    traj = r.pi_trajectory(steps) if P else r.mc_trajectory(steps)
    return averages.chain_series(traj, traj[0, 1])


def report(tag, series_list, corrtime=20):
    for key in ("energy", "aux"):
        blocks = np.stack([averages.block_means(s[key], cases.AVG_BLOCKS) for s in series_list])
        roots = [averages.root_average(s[key][::corrtime]) for s in series_list]      # what the reference prints per chain
        print("%-10s %-7s mean %.8g  blocked s.e.m. %.3g   per chain (update_root_averages, every %d steps): %s" % (
            tag, key, blocks.mean(), blocks.std(ddof=1) / np.sqrt(blocks.size), corrtime, ", ".join("%.6g+-%.2g" % r for r in roots)))
    return series_list


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("case", choices=sorted(cases.AVERAGES))
    ap.add_argument("--seeds", type=int, nargs="+", default=[201, 202, 203, 204])
    ap.add_argument("--reference", action="store_true")
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    ctx = mp.get_context("spawn")
    with ctx.Pool(min(len(a.seeds), max(1, a.gpus))) as pool:
        ours = pool.map(_ours, [(a.case, sd, i % max(1, a.gpus)) for i, sd in enumerate(a.seeds)])
    report("engine", ours)
    if a.reference:
        with ctx.Pool(min(len(a.seeds), os.cpu_count() or 1)) as pool:
            refs = pool.map(_ref, [(a.case, sd + 1000) for sd in a.seeds])
        report("reference", refs)
        for key in ("energy", "aux"):
            c = averages.compare(np.stack([averages.block_means(s[key], cases.AVG_BLOCKS) for s in ours]),
                                 np.stack([averages.block_means(s[key], cases.AVG_BLOCKS) for s in refs]))
            print("difference of the means of %-6s: %.3g = %.2f combined standard errors" % (key, c["mean_a"] - c["mean_b"], c["z"]))
