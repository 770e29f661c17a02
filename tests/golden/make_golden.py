"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built by `make -C oracle ref`).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Each fixture stores the full input (flat, list-ordered sites + cell + input keywords) and the reference's
outputs read as doubles through oracle/ref_harness.cpp: every energy sub-term, and for polarizable cases the
per-site dipoles / fields / rank metric.  tests/cases.py names the cases; this script only evaluates them.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref  # noqa: E402
from tests import cases  # noqa: E402


def pack_system(s):
    return dict(basis=s.basis, pos=s.pos, charge_e=s.charge_e, alpha=s.alpha, eps=s.eps, sigma=s.sigma, mass=s.mass,
                mol=s.mol, frozen=s.frozen, opts=np.array(json.dumps(s.opts)), atomtype=np.array(s.atomtype),
                moltype=np.array(s.moltype))


def main(only=None):
    for name, build in cases.CLASSIC.items():
        if only and name not in only:
            continue
        s = build()
        r = ref.RefSystem(s, ensemble="nvt")
        t = r.terms()
        d = r.dipoles()
        c = r.cell()
        out = pack_system(s)
        out.update({"ref_" + k: np.float64(v) for k, v in t.items()})
        out.update({"ref_" + k: v for k, v in d.items()})
        out.update({"cell_" + k: v for k, v in c.items()})
        # the MC-loop view: energy() as mc() calls it, cold then after moving one molecule (warm pair cache)
        r2 = ref.RefSystem(s, ensemble="nvt")
        e0 = r2.energy()
        moved = cases.displaced(s)
        r2.set_pos(moved.pos)
        e1 = r2.energy()
        out["ref_energy_cold"] = np.array([e0[k] for k in ("energy", "rd", "coulombic", "polar")])
        out["ref_energy_moved"] = np.array([e1[k] for k in ("energy", "rd", "coulombic", "polar")])
        out["moved_pos"] = moved.pos
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("%-22s n=%4d  E=%.15g  pol=%.15g  it=%d" % (name, s.n, e0["energy"], t["polar"], t["iterations"]))
    for name, build in cases.PI.items():
        if only and name not in only:
            continue
        tmpl, beads = build()
        P = beads.shape[0]
        r = ref.RefSystem(tmpl, P=P)
        for b in range(P):
            r.set_pos(beads[b], b)
        e = r.pi_energy()
        out = pack_system(tmpl)
        out["beads"] = beads
        out.update({"ref_pi_" + k: np.float64(v) for k, v in e.items()})
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("%-22s P=%d n=%4d  U=%.15g  K=%.15g" % (name, P, tmpl.n, e["potential"], e["kinetic"]))


def parsed():
    """What the reference's own readers (input file + PQR) make of a job: the flat site table and the cell, with the job's text —
    the CPU-side yardstick for the host mirror's readers (tests/test_host_cpu.py)."""
    import tempfile
    from mpmcxx_b200 import workloads
    for name, (build, P) in cases.PARSED.items():
        s = build()
        with tempfile.TemporaryDirectory(prefix="mref_parsed_") as d:
            r = ref.RefSystem(s, P=P, workdir=d)
            out = {"input_in": np.array(open(os.path.join(d, "input.in")).read()), "input_pqr": np.array(open(os.path.join(d, "input.pqr")).read()),
                   "P": np.int32(P)}
            out.update({"ref_" + k: v for k, v in r.sites(0 if P else -1).items()})
            out.update({"cell_" + k: v for k, v in r.cell(0 if P else -1).items()})
        np.savez_compressed(os.path.join(HERE, "parsed_" + name + ".npz"), **out)
        print("parsed_%-20s n=%d" % (name, len(out["ref_charge"])))


def parsed_box_from_pqr():
    """`read_pqr_box on`: the cell comes from the REMARK BOX BASIS lines of the geometry file (here: a file the reference itself wrote for
    a triclinic system, CRYST1 / CONECT records included), not from the basis lines of the input, which this job sets to a wrong cube."""
    import tempfile
    z = np.load(os.path.join(HERE, "pqr_written.npz"))
    s = cases._with(cases.W.triclinic_mix(solver=cases.W.SOLVER_GS_RANKED_PALMO), read_pqr_box="on")
    s.basis = np.eye(3) * 11.0
    with tempfile.TemporaryDirectory(prefix="mref_parsed_") as d:
        cases.W.write_reference_job(s, d)
        with open(os.path.join(d, "input.pqr"), "w") as fp:
            fp.write(str(z["text_tri_gs_ranked_palmo"]))
        r = ref.RefSystem.from_directory(d, "input.in", 0)
        out = {"input_in": np.array(open(os.path.join(d, "input.in")).read()), "input_pqr": np.array(open(os.path.join(d, "input.pqr")).read()),
               "P": np.int32(0)}
        out.update({"ref_" + k: v for k, v in r.sites(-1).items()})
        out.update({"cell_" + k: v for k, v in r.cell(-1).items()})
    np.savez_compressed(os.path.join(HERE, "parsed_tri_box_from_pqr.npz"), **out)
    print("parsed_tri_box_from_pqr n=%d basis %s" % (len(out["ref_charge"]), out["cell_basis"].tolist()))


def one_trajectory(name):
    build, P, steps = cases.TRAJ[name]
    s = build()
    r = ref.RefSystem(s, P=P)
    traj = r.pi_trajectory(steps) if P else r.mc_trajectory(steps)
    out = pack_system(s)
    out["P"] = np.int32(P)
    out["traj"] = traj
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-24s steps=%d acceptance=%.3f moves=%s" % (name, steps, traj[:, 3].mean(), np.bincount(traj[:, 0].astype(int)).tolist()), flush=True)


def trajectories():
    """Seeded runs of the reference's own Markov chains (System::mc loop body / PI_nvt_mc loop body replayed by the harness).
    One fresh process per case: the reference's global Rando keeps a cached Box-Muller value inside its normal_distribution
    (src/Rando.h:12-14) that Rando::seed does not clear, so a run's draws depend on what the process did before."""
    import subprocess
    for name in cases.TRAJ:
        subprocess.run([sys.executable, os.path.abspath(__file__), "traj1", name], check=True)


def one_shipped(name):
    """A sample directory shipped with the reference, copied as it is and run by the reference itself (fresh process, see trajectories())."""
    import shutil
    import tempfile
    sub, names, inp, P, steps = cases.SHIPPED[name]
    src = os.path.join("/root/reference/sample-input", sub)
    with tempfile.TemporaryDirectory(prefix="mref_shipped_") as d:
        for f in names:
            shutil.copy(os.path.join(src, f), os.path.join(d, f))
        r = ref.RefSystem.from_directory(d, inp, P)
        traj = r.pi_trajectory(steps) if P else r.mc_trajectory(steps)
    texts = [open(os.path.join(src, f)).read() for f in names]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), file_names=np.array(names), file_texts=np.array(texts), input_name=np.array(inp),
                        P=np.int32(P), traj=traj)
    print("%-28s steps=%d acceptance=%.3f kinetic[0..2]=%s" % (name, steps, traj[:, 3].mean(), traj[:3, 4].tolist()), flush=True)


def one_average(name, seed):
    """One chain of the reference for the ensemble-average comparison: block means of the current-state series."""
    from mpmcxx_b200 import averages
    build, P, steps, _, _ = cases.AVERAGES[name]
    s = build()
    s.opts["seed"] = str(seed)
    r = ref.RefSystem(s, P=P)
    first = r.pi_potential() if P else r.energy()["energy"]
    r2 = ref.RefSystem(s, P=P)          # the trajectory replay starts from its own fresh state
    traj = r2.pi_trajectory(steps) if P else r2.mc_trajectory(steps)
    ser = averages.chain_series(traj, first)
    out = {"energy": averages.block_means(ser["energy"], cases.AVG_BLOCKS), "aux": averages.block_means(ser["aux"], cases.AVG_BLOCKS),
           "acceptance": np.float64(ser["accepted"].mean())}
    np.savez(os.path.join(HERE, "_avg_%s_%d.npz" % (name, seed)), **out)


def averages_golden():
    """Per chain a fresh process (Rando's cached normal, see trajectories()); the per-seed pieces are merged into one small fixture."""
    import subprocess
    for name, (build, P, steps, ref_seeds, _) in cases.AVERAGES.items():
        for seed in ref_seeds:
            subprocess.run([sys.executable, os.path.abspath(__file__), "avg1", name, str(seed)], check=True)
        parts = [np.load(os.path.join(HERE, "_avg_%s_%d.npz" % (name, seed))) for seed in ref_seeds]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), seeds=np.array(ref_seeds), energy=np.stack([p["energy"] for p in parts]),
                            aux=np.stack([p["aux"] for p in parts]), acceptance=np.array([p["acceptance"] for p in parts]), steps=np.int32(steps))
        for seed in ref_seeds:
            os.remove(os.path.join(HERE, "_avg_%s_%d.npz" % (name, seed)))
        z = np.load(os.path.join(HERE, name + ".npz"))
        print("%-18s <E> %.6g +- %.3g   <aux> %.6g +- %.3g   acceptance %s" % (name, z["energy"].mean(), z["energy"].std(ddof=1) / np.sqrt(z["energy"].size),
              z["aux"].mean(), z["aux"].std(ddof=1) / np.sqrt(z["aux"].size), z["acceptance"].round(3).tolist()), flush=True)


def written():
    """PQR files written by the reference's own writer, and what its restart selection picks for the shipped pi000 `input.in`
    (parallel_restarts on, `-P 8`: four restart files in the directory, the other four bead systems fall back to system 0's ".last" copy)."""
    import shutil
    import tempfile
    from mpmcxx_b200 import workloads
    out = {}
    for name, (build, P, sidx) in cases.WRITTEN.items():
        s = build()
        with tempfile.TemporaryDirectory(prefix="mref_written_") as d:
            r = ref.RefSystem(s, P=P, workdir=d)
            path = os.path.join(d, "written.pqr")
            r.write_pqr(path, sidx)
            out["text_" + name] = np.array(open(path).read())
            out["names_" + name] = np.array("\n".join(r.io_filenames(sidx)))
        print("written %-22s %d bytes  %s" % (name, len(str(out["text_" + name])), str(out["names_" + name]).replace("\n", " | ")))
    src = "/root/reference/sample-input/pi000-free-argon-2K"
    with tempfile.TemporaryDirectory(prefix="mref_restart_") as d:
        files = ["input.in"] + ["Ar2K.restart-%04d.pqr" % i for i in range(4)] + ["Ar2K.restart-0000.pqr.last"]
        for f in files:
            shutil.copy(os.path.join(src, f), os.path.join(d, f))
        r = ref.RefSystem.from_directory(d, "input.in", 8)
        out["restart_files"] = np.array(files)
        out["restart_texts"] = np.array([open(os.path.join(src, f)).read() for f in files])
        out["restart_names"] = np.array(["\n".join(os.path.basename(x) for x in r.io_filenames(i)) for i in range(8)])
        out["restart_pos"] = np.stack([r.sites(i)["pos"] for i in range(8)])
        e = r.pi_energy()
        out["restart_kinetic"] = np.float64(e["kinetic"])
        out["restart_chain"] = np.float64(e["chain_mass_len2"])
        print("restart selection:", [str(x).replace("\n", " | ") for x in out["restart_names"]], "kinetic", e["kinetic"])
    np.savez_compressed(os.path.join(HERE, "pqr_written.npz"), **out)


def shipped():
    import subprocess
    for name in cases.SHIPPED:
        subprocess.run([sys.executable, os.path.abspath(__file__), "shipped1", name], check=True)


ROOT_AVG_KEYS = ["energy", "energy_error", "N", "N_error", "coulombic_energy", "coulombic_energy_error", "rd_energy", "rd_energy_error",
                 "polarization_energy", "polarization_energy_error", "density", "density_error", "heat_capacity", "heat_capacity_error",
                 "compressibility", "compressibility_error", "percent_wt", "percent_wt_me", "excess_ratio", "qst", "pore_density", "NU"]


def root_averages_golden():
    """System::update_root_averages of the unmodified reference over a synthetic series of 300 samples on the uVT pore job (frozen
    framework, free volume): the means, errors and derived adsorption observables mpmcxx_b200/averages.py restates."""
    s = cases.uvt_pore_for_averages()
    r = ref.RefSystem(s)
    x = cases.root_average_samples()
    o = r.root_averages(x)
    np.savez_compressed(os.path.join(HERE, "root_averages.npz"), samples=x, keys=np.array(ROOT_AVG_KEYS), values=o[:22], frozen_mass=o[22], volume=o[23], fugacity=o[24])
    print("root averages:", dict(zip(ROOT_AVG_KEYS, o[:22].tolist())))


def one_mc_average(name):
    build, steps, corrtime = cases.MC_AVERAGES[name]
    s = build()
    s.opts.update({"numsteps": str(steps), "corrtime": str(corrtime)})
    r = ref.RefSystem(s)
    o = r.mc_averages(steps, corrtime)
    np.save(os.path.join(HERE, "_mcavg_%s.npy" % name), o)
    print("%-12s" % name, dict(zip(ROOT_AVG_KEYS, o[:22].tolist())), flush=True)


def mc_averages_golden():
    """Fresh process per chain (function-static sample counter, Rando's cached normal)."""
    import subprocess
    out = {}
    for name in cases.MC_AVERAGES:
        subprocess.run([sys.executable, os.path.abspath(__file__), "mcavg1", name], check=True)
        out[name] = np.load(os.path.join(HERE, "_mcavg_%s.npy" % name))
        os.remove(os.path.join(HERE, "_mcavg_%s.npy" % name))
    np.savez_compressed(os.path.join(HERE, "mc_averages.npz"), keys=np.array(ROOT_AVG_KEYS + ["frozen_mass", "volume", "fugacity"]), **out)


PI_AVG_KEYS = ["energy", "energy_error", "kinetic_energy", "kinetic_energy_error", "rd_energy", "rd_energy_error", "coulombic_energy",
               "coulombic_energy_error", "polarization_energy", "polarization_energy_error", "N", "N_error", "density", "density_error",
               "heat_capacity", "heat_capacity_error", "compressibility", "compressibility_error", "frozen_mass", "volume"]


def one_pi_average(name):
    build, P, steps, corrtime = cases.PI_AVERAGES[name]
    s = build()
    s.opts.update({"numsteps": str(steps), "corrtime": str(corrtime)})
    r = ref.RefSystem(s, P=P)
    o = r.pi_averages(steps, corrtime)
    np.save(os.path.join(HERE, "_piavg_%s.npy" % name), o)
    print("%-16s" % name, dict(zip(PI_AVG_KEYS, o.tolist())), flush=True)


def pi_averages_golden():
    import subprocess
    out = {}
    for name in cases.PI_AVERAGES:
        subprocess.run([sys.executable, os.path.abspath(__file__), "piavg1", name], check=True)
        out[name] = np.load(os.path.join(HERE, "_piavg_%s.npy" % name))
        os.remove(os.path.join(HERE, "_piavg_%s.npy" % name))
    np.savez_compressed(os.path.join(HERE, "pi_averages.npz"), keys=np.array(PI_AVG_KEYS), **out)


def one_final_pqr(name):
    import tempfile
    build, P, steps, sidx = cases.FINAL_PQR[name]
    s = build()
    s.opts.update({"numsteps": str(steps)})
    with tempfile.TemporaryDirectory(prefix="mref_final_") as d:
        r = ref.RefSystem(s, P=P, workdir=d)
        traj = r.pi_trajectory(steps) if P else r.mc_trajectory(steps)
        path = os.path.join(d, "final_state.pqr")
        r.write_pqr(path, sidx)
        text = open(path).read()
    np.save(os.path.join(HERE, "_final_%s.npy" % name), np.array(text))
    print("%-12s %d steps, acceptance %.3f, %d bytes" % (name, steps, traj[:, 3].mean(), len(text)), flush=True)


def final_pqr_golden():
    import subprocess
    out = {}
    for name in cases.FINAL_PQR:
        subprocess.run([sys.executable, os.path.abspath(__file__), "final1", name], check=True)
        out[name] = np.load(os.path.join(HERE, "_final_%s.npy" % name))
        os.remove(os.path.join(HERE, "_final_%s.npy" % name))
    np.savez_compressed(os.path.join(HERE, "final_pqr.npz"), **out)


def one_input_error(name):
    """The error code the reference throws while it reads and validates a (malformed) job; 0 = accepted.  Fresh process per job."""
    build, P, mutate = cases.INPUT_ERRORS[name]
    s = build()
    mutate(s)
    try:
        ref.RefSystem(s, P=P)
        code = 0
    except RuntimeError as e:
        code = int(str(e).split()[-1])
    print("%s %d" % (name, code), flush=True)


def input_errors():
    import subprocess
    out = {}
    for name in cases.INPUT_ERRORS:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "error1", name], capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith(name + " ")][-1]
        out[name] = int(line.split()[-1])
        print(line)
    with open(os.path.join(HERE, "input_errors.json"), "w") as fp:
        json.dump(out, fp, indent=1, sort_keys=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "error1":
        one_input_error(sys.argv[2])
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[1] == "final1":
        one_final_pqr(sys.argv[2])
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "finalpqr":
        final_pqr_golden()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[1] == "piavg1":
        one_pi_average(sys.argv[2])
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "piavg":
        pi_averages_golden()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[1] == "mcavg1":
        one_mc_average(sys.argv[2])
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "mcavg":
        mc_averages_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "rootavg":
        root_averages_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "errors":
        input_errors()
        sys.exit(0)
    if len(sys.argv) > 3 and sys.argv[1] == "avg1":
        one_average(sys.argv[2], int(sys.argv[3]))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "written":
        written()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "averages":
        averages_golden()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[1] == "shipped1":
        one_shipped(sys.argv[2])
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "shipped":
        shipped()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[1] == "traj1":
        one_trajectory(sys.argv[2])
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "parsedbox":
        parsed_box_from_pqr()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "parsed":
        parsed()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "traj":
        trajectories()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[1] == "only":          # regenerate just the named energy cases
        main(set(sys.argv[2:]))
        sys.exit(0)
    main()
    parsed()
    trajectories()
    shipped()
    averages_golden()
    written()
