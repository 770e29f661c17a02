#!/usr/bin/env python
"""bench.py — MC moves/s of the mpmc++ energy hot path on B200 (one JSON line; contract in the task brief).

A "step" is one Monte Carlo trial move on the headline workload: BASELINE.json config 4, the polarizable
H2-in-framework system (8000 frozen framework sites + 400 five-site H2 = 10 000 sites, Ewald kmax 7, Thole
exponential damping, solver = GS-ranked x4 + Palmo unless --solver says otherwise).  One move = displace one H2
rigidly on the host -> mpmc_update_sites -> mpmc_energy (a FULL System::energy(): LJ + Ewald real/reciprocal/self +
Thole solve) -> Metropolis accept/reject on the host (restore through mpmc_update_sites on reject).

  value : moves/s with the configuration resident in HBM: K full energy evaluations timed with CUDA events on the
          engine's stream (per-step events, L2 flushed between steps), max over ranks.
  e2e   : moves/s through the C-ABI with HOST buffers: the move's coordinates go host->device and every energy
          component comes back device->host inside the timed region (wall clock around each step, max over ranks).
  N > 1 : the classic ensembles do not shard inside a move: N independent Markov chains, one per GPU,
          no data-path collective -> weak scaling.  (The bead-sharded path-integral config is `--workload pi_h2`.)
  --impl reference : the reference's own CPU energy() (oracle/_ref, the unmodified reference compiled from its
          sources) on the host cores, on a bounded sample of the same workload (see `cpu_baseline.sample`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mpmcxx_b200 import workloads as W  # noqa: E402

FLOP_PER_DIPOLE_PAIR = 100.0     # K6, geometry + damping recomputed on the fly, per ORDERED pair per sweep
FLOP_PER_LJ_PAIR = 54.0          # K1
FLOP_PER_ES_PAIR = 82.0          # K1 + K2
BYTES_PER_SITE_SWEEP = 56.0 + 4.0 + 24.0 + 24.0   # 56 B/site + id/flags + mu read + mu written


def _np_default(o):
    if isinstance(o, np.generic):
        return o.item()
    raise TypeError(type(o))


def env_int(k, d):
    return int(os.environ.get(k, d))


# ----------------------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------------------
SOLVERS = {"gs_ranked_palmo": W.SOLVER_GS_RANKED_PALMO, "jacobi10": W.SOLVER_JACOBI10}


def build_workload(name, solver, scale=1.0):
    if name == "h2_framework":
        ncell = max(2, int(round(20 * scale)))
        n_h2 = max(2, int(round(400 * scale ** 3)))
        s = W.h2_framework(ncell=ncell, n_h2=n_h2, solver=SOLVERS[solver], ensemble="nvt")
        desc = "config4: H2 in frozen framework, %d sites (%d frozen + %d x 5-site H2), L=%g A, Ewald kmax 7, Thole exp damping, solver %s" % (
            s.n, ncell ** 3, n_h2, ncell * 4.0, solver)
        return s, desc
    if name == "lj_argon":
        side = max(2, int(round(16 * scale)))
        s = W.lj_argon(n_side=side, L=60.0 * side / 16)
        return s, "config3: bulk LJ argon NVT, %d atoms, L=%g A, rd_lrc on, rd_only" % (s.n, 60.0 * side / 16)
    raise ValueError(name)


class MoveGen:
    """Host-side trial moves: rigid translation (uniform in +-step per axis) and rotation about a random axis of one
    uniformly chosen mobile molecule (System.MonteCarlo.cpp:875 displace); Metropolis at temperature T."""

    def __init__(self, s, seed, step=0.5, max_angle_deg=18.0):
        self.s = s
        self.rs = np.random.RandomState(seed)
        starts = np.nonzero(np.diff(np.concatenate([[-1], s.mol])))[0]
        ends = np.concatenate([starts[1:], [s.n]])
        self.mols = [(int(a), int(b)) for a, b in zip(starts, ends) if not s.frozen[a]]
        self.step, self.ang = step, np.deg2rad(max_angle_deg)
        self.T = float(s.opts.get("temperature", 77.0))

    def propose(self, pos):
        a, b = self.mols[self.rs.randint(len(self.mols))]
        old = pos[a:b].copy()
        m = self.s.mass[a:b]
        com = (old * m[:, None]).sum(0) / m.sum() if m.sum() > 0 else old.mean(0)
        ax = self.rs.normal(size=3)
        ax /= np.linalg.norm(ax)
        th = (self.rs.random_sample() * 2 - 1) * self.ang
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
        new = (old - com) @ R.T + com + (self.rs.random_sample(3) * 2 - 1) * self.step
        return a, old, new

    def accept(self, e_old, e_new):
        if not np.isfinite(e_new):
            return False                         # System.MonteCarlo.cpp:56-59
        return self.rs.random_sample() < np.exp(min(0.0, -(e_new - e_old) / self.T))


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU through NVML, in-process: one synchronous sample when the timed region starts, one
    when it ends, and a background thread sampling every 10 ms in between (nvidia-smi -lms needs ~0.3 s to produce its first row,
    longer than a short timed region, and one subprocess per rank)."""
    BITS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.idx, self.rows, self.h, self.nv, self.run = gpu_index, [], None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        try:
            sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            try:
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            try:
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
            except Exception:
                util = -1
            self.rows.append((sm, mx, rs, util))
        except Exception:
            pass

    def _loop(self):
        while self.run:
            self._sample()
            time.sleep(0.01)

    def start(self):
        if not self.nv:
            return
        self._sample()
        self.run = True
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()

    def stop(self):
        if self.nv:
            self.run = False
            try:
                self.th.join(timeout=1)
            except Exception:
                pass
            self._sample()
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        sm = [r[0] for r in self.rows]
        reasons = sorted(k for k, bit in self.BITS.items() if any(r[2] & bit for r in self.rows))
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": float(max(r[1] for r in self.rows)), "reasons": reasons,
                "samples": len(sm), "how": "NVML in-process, every 10 ms over the timed regions + one sample at each end"}


# ----------------------------------------------------------------------------------------------------------------
# CPU baseline (the one place bench.py may execute oracle/)
# ----------------------------------------------------------------------------------------------------------------
def _ref_worker(args):
    """One independent chain on one host core: reference energy() after one-molecule moves (warm pair cache)."""
    workload, solver, scale, nsteps, seed, kind = args
    from oracle import port, ref
    s, _ = build_workload(workload, solver, scale)
    gen = MoveGen(s, seed)
    pos = s.pos.copy()
    if kind == "reference":
        r = ref.RefSystem(s, ensemble="nvt")
        r.energy()                                  # cold first evaluation (mc_initial_energy), untimed
        t0 = time.perf_counter()
        for _ in range(nsteps):
            a, old, new = gen.propose(pos)
            pos[a:a + len(new)] = new
            r.set_pos(pos)
            r.energy()
        dt = time.perf_counter() - t0
    else:
        port.energy(s, pos=pos, want_sites=False)
        t0 = time.perf_counter()
        for _ in range(nsteps):
            a, old, new = gen.propose(pos)
            pos[a:a + len(new)] = new
            port.energy(s, pos=pos, want_sites=False)
        dt = time.perf_counter() - t0
    return dt, s.n


def _ref_full_worker(args):
    """One independent chain of the reference at the FULL size of the workload: open (pair-list allocation), one cold energy()
    (mc_initial_energy), then `nwarm` timed energy() calls each after a one-molecule move (the warm-cache case every MC step is)."""
    workload, solver, nwarm, seed = args
    import resource
    from oracle import ref
    s, _ = build_workload(workload, solver, 1.0)
    gen = MoveGen(s, seed)
    pos = s.pos.copy()
    t0 = time.perf_counter()
    r = ref.RefSystem(s, ensemble="nvt")
    t_open = time.perf_counter() - t0
    t0 = time.perf_counter()
    e0 = r.energy()
    t_cold = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(nwarm):
        a, old, new = gen.propose(pos)
        pos[a:a + len(new)] = new
        r.set_pos(pos)
        r.energy()
    dt = time.perf_counter() - t0
    return dt, t_open, t_cold, resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6, e0["energy"], s.n


REF_CHAIN_GB = {"h2_framework": 18.0, "lj_argon": 2.2}     # resident set of one reference chain (10 GB of Pair nodes + 7.2 GB A_matrix at N = 10^4), measured


def cpu_baseline(workload, solver, steps, budget_s=20.0, scale=None, chains=None, nwarm=3):
    """moves/s of the reference's own CPU energy() on the host cores AT THE FULL SIZE of the workload.  The reference is single-threaded
    inside energy(); its all-cores mode is independent Markov chains, so this runs min(cores, RAM / footprint) chains of
    oracle/_ref side by side, each: open + 1 cold energy() untimed, then `nwarm` timed moves (one-molecule displace -> energy(), warm
    pair cache).  value = chains x nwarm / slowest chain's timed seconds.  A secondary figure on a 1/8-size system is kept for
    continuity with round 1 (`scaled_sample`) when budget_s > 0.  Where oracle/_ref is absent the C restatement (oracle.c, OpenMP over
    all cores) is timed at full size instead (kind "port")."""
    import multiprocessing as mp
    from oracle import ref
    kind = "reference" if ref.available() else "port"
    cores = os.cpu_count() or 1
    full, _ = build_workload(workload, solver, 1.0)
    if kind == "port":
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        t_probe, _ = _ref_worker((workload, solver, 1.0, 1, 1, kind))
        nsteps = int(max(1, min(steps, budget_s / max(t_probe, 1e-3))))
        dt, _ = _ref_worker((workload, solver, 1.0, nsteps, 100, kind))
        return {"value": nsteps / dt, "unit": "moves/s", "cores": cores, "kind": kind, "host_cpus": cores,
                "sample": "%d moves of the full-size system (N=%d) through oracle.c with %d OpenMP threads (oracle/_ref not built on this box)" % (nsteps, full.n, cores)}
    try:
        import psutil
        avail_gb = psutil.virtual_memory().available / 1e9
    except Exception:
        avail_gb = 64.0
    per_chain = REF_CHAIN_GB.get(workload, 18.0)
    chains = chains or max(1, min(cores, int(0.85 * avail_gb / per_chain)))
    nwarm = max(1, min(nwarm, steps))
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(chains) as pool:
        res = pool.map(_ref_full_worker, [(workload, solver, nwarm, 100 + c) for c in range(chains)])
    wall_total = time.perf_counter() - t0
    wall = max(r[0] for r in res)
    value = chains * nwarm / wall
    out = {"value": value, "unit": "moves/s", "cores": chains, "kind": kind, "host_cpus": cores,
           "sample": "%d independent chains of the unmodified reference (oracle/_ref) at the full size N=%d, each 1 cold energy() untimed + %d timed moves "
                     "(one-molecule displace -> energy(), warm pair cache); chains = min(%d cores, %.0f GB available / %.0f GB per chain)"
                     % (chains, full.n, nwarm, cores, avail_gb, per_chain),
           "same_config": True, "chains": chains, "moves_per_chain": nwarm,
           "s_per_move_per_chain": wall / nwarm, "s_open": max(r[1] for r in res), "s_cold_energy": max(r[2] for r in res),
           "rss_gb_per_chain": max(r[3] for r in res), "host_ram_available_gb": avail_gb, "wall_s": wall_total,
           "step0_energy_K": res[0][4]}
    if budget_s > 0:
        # round 1's figure: the same system at half the linear size, all cores, scaled by (N_s/N)^2
        sc = 0.5 if scale is None else scale
        small, _ = build_workload(workload, solver, sc)
        t_probe, _ = _ref_worker((workload, solver, sc, 1, 1, kind))
        nsteps = int(max(1, min(steps, budget_s / max(t_probe, 1e-3))))
        with mp.get_context("spawn").Pool(cores) as pool:
            rs = pool.map(_ref_worker, [(workload, solver, sc, nsteps, 100 + c, kind) for c in range(cores)])
        ms = cores * nsteps / max(r[0] for r in rs)
        out["scaled_sample"] = {"sites": small.n, "chains": cores, "moves_per_chain": nsteps, "measured_moves_per_s": ms,
                                "scaled_to_full_by_N2": ms * (small.n / full.n) ** 2}
    return out


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from mpmcxx_b200 import engine

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    s, desc = build_workload(args.workload, args.solver, args.scale)
    eng = engine.Engine(s, device=local)
    gen = MoveGen(s, seed=1000 + rank)
    pos = s.pos.copy()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def flush_l2():
        flush_buf.zero_()
        torch.cuda.synchronize()

    ext = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local))
    peak_tflops, _ = engine.probe_fp64_peak(local)

    step0 = eng.energy()                               # the start configuration, every component: any run can be checked after the fact
    e_cur = step0["energy"]
    naccept = 0

    def mc_step():
        nonlocal e_cur, naccept
        a, old, new = gen.propose(pos)
        eng.update_sites(a, new)                       # host -> device
        out = eng.energy()                             # kernels + device -> host of every component
        if gen.accept(e_cur, out["energy"]) and not out["iterator_failed"]:
            pos[a:a + len(new)] = new
            e_cur = out["energy"]
            naccept += 1
        else:
            eng.update_sites(a, old)                   # restore()
        return out

    for _ in range(args.warmup):
        mc_step()
    # ---- e2e: wall clock around each step through the C-ABI with host buffers --------------------------------------
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    launches0 = eng.launches()
    t_e2e = 0.0
    for _ in range(args.steps):
        flush_l2()
        t0 = time.perf_counter()
        last = mc_step()
        t_e2e += time.perf_counter() - t0
    barrier()
    launches_e2e = eng.launches() - launches0
    # ---- value: device-resident, CUDA events on the engine's stream -------------------------------------------------
    t_dev_ms = 0.0
    launches1 = eng.launches()
    def mark_one_moved():
        """What the engine tracks between moves (the structure-factor chunks of moved sites) must be redone in a device-resident step as
        it is after a real move: mark one H2 as moved, without a copy (mpmc_debug_mark_moved)."""
        a, b = gen.mols[gen.rs.randint(len(gen.mols))]
        eng.mark_moved(a, b - a)

    for _ in range(args.steps):
        flush_l2()
        mark_one_moved()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        eng.enqueue()
        e1.record(ext)
        eng.fetch()
        e1.synchronize()
        t_dev_ms += e0.elapsed_time(e1)
    barrier()
    launches_dev = eng.launches() - launches1
    # ---- the same K evaluations once more with a CUDA-event pair around every kernel class (the roofline's launch durations).  Its own
    #      pass: the event pairs serialise the two streams the untimed schedule overlaps, so neither `value` nor `e2e` is measured with them on.
    eng.set_timing(True)
    t_prof_ms = 0.0
    for _ in range(args.steps):
        flush_l2()
        mark_one_moved()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        eng.enqueue()
        e1.record(ext)
        eng.fetch()
        e1.synchronize()
        t_prof_ms += e0.elapsed_time(e1)
    timing = eng.timing()
    eng.set_timing(False)
    barrier()
    clk = clocks.stop()

    t_dev = max_over_ranks(t_dev_ms * 1e-3)
    t_wall = max_over_ranks(t_e2e)
    value = world * args.steps / t_dev
    e2e_value = world * args.steps / t_wall
    nmol_sites = len(gen.mols) and (gen.mols[0][1] - gen.mols[0][0])
    polar_on = s.opts.get("polarization") == "on"

    # ---- roofline of the dominant kernel (live CUDA-event timing of that kernel class inside the e2e region) ---------
    np_pol = int(np.count_nonzero(s.alpha))
    cands = {k: v for k, v in timing.items() if k != "energy_total" and v[1] > 0}
    dom = max(cands, key=lambda k: cands[k][0]) if cands else None
    roof = None
    if dom:
        ms_per_launch = cands[dom][0] / cands[dom][1]
        if dom in ("gs_sweep", "dipole_sweep", "palmo"):
            flops = FLOP_PER_DIPOLE_PAIR * np_pol * (np_pol - 1)
            work = "%d x %d ordered dipole pairs x %g flop (SURVEY 8d K6)" % (np_pol, np_pol - 1, FLOP_PER_DIPOLE_PAIR)
        elif dom == "pair":
            flops = (FLOP_PER_LJ_PAIR if s.opts.get("rd_only") == "on" else FLOP_PER_ES_PAIR) * last["n_pair_evals"]
            work = "%d pairs x %g flop (SURVEY 8d K1/K2)" % (last["n_pair_evals"], flops / last["n_pair_evals"])
        else:
            flops, work = float("nan"), "n/a"
        ach = flops / (ms_per_launch * 1e-3) / 1e12
        bytes_alg = BYTES_PER_SITE_SWEEP * s.n
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the round's `ncu --set full` capture of this command (profiles/):
        # kept in a small JSON next to the summaries so that the number in the line is the one the committed capture shows
        ncu_traffic, traffic_src = {}, None
        try:
            with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as fp:
                tj = json.load(fp)
            ncu_traffic, traffic_src = tj.get("bytes_per_launch", {}), tj.get("source")
        except Exception:
            pass
        roof = {"kernel": dom, "bound": "fp64", "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach / peak_tflops,
                "traffic": ncu_traffic.get(dom), "traffic_source": traffic_src,
                "traffic_note": "a Gauss-Seidel sweep streams what was precomputed for its order once: the tensors between each 64-site block and the 4 blocks after it (113 MB) and the block inverses (25 MB); the roofline is the FP64 pipe, not HBM",
                "ms_per_launch": ms_per_launch, "launches_timed": cands[dom][1],
                "launch": "one Gauss-Seidel sweep = k_gs_pipeline (8-CTA cluster) + k_gs_updaters (140 CTAs) side by side" if dom == "gs_sweep" else dom,
                "work_per_launch": work,
                "peak_source": "measured in this run: register-resident DFMA loop on all SMs (MEASURED_PEAKS.json has no FP64 entry)",
                "share_of_step": cands[dom][0] / max(timing["energy_total"][0], 1e-9),
                "hbm_view": {"algorithmic_bytes": bytes_alg, "achieved_gbs": bytes_alg / (ms_per_launch * 1e-3) / 1e9,
                             "note": "the contraction reads O(N) bytes per sweep; HBM is not the binding limit (FP64 pipe is)"},
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in timing.items()},
                "timed_pass_ms_per_step": t_prof_ms / args.steps}

    out = {"metric": "mc_moves_per_sec", "value": value, "unit": "moves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": desc, "chains": world, "parallelism": "replicas only (one independent Markov chain per GPU, no collective)",
                      "l2": "flushed between timed iterations (256 MiB write); working set itself is < 1 MB",
                      "moves": "rigid displace+rotate of one H2, Metropolis T=%g K" % gen.T,
                      "pair_evals_per_move": last["n_pair_evals"], "polarization_iterations": last["polarization_iterations"],
                      "step0_energy_K": {k: step0[k] for k in ("energy", "rd_energy", "coulombic_energy", "polarization_energy", "rd_pair", "rd_lrc_pair",
                                                               "rd_lrc_self", "es_real", "es_self_intra", "es_reciprocal", "es_self", "polarization_iterations",
                                                               "n_pairs_in_cutoff")}},
           "clocks": clk,
           "e2e": {"value": e2e_value, "unit": "moves/s", "h2d_bytes_per_step": 32 * nmol_sites * (2 - naccept / max(1, args.steps + args.warmup)),
                   "d2h_bytes_per_step": 64, "ms_per_step": 1e3 * t_wall / args.steps,
                   "acceptance": naccept / max(1, args.steps + args.warmup)},
           "gpu_launches": int(sum_over_ranks(launches_e2e + launches_dev)),
           "pair_evals_per_sec": value * last["n_pair_evals"],
           "roofline": roof}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args.workload, args.solver, args.steps, budget_s=args.cpu_budget)
    eng.close()
    if rank == 0 and world == 1 and not args.no_extra:
        out["extra"] = extra_workloads(engine, local, peak_tflops)
        try:
            out["extra"]["h2_framework_uvt_mirror"] = uvt_through_mirror(local)
        except Exception as ex:                                         # secondary figure: never lose the line over it
            out["extra"]["h2_framework_uvt_mirror"] = {"error": str(ex)}
    if not args.no_extra:
        # BASELINE config 5, the bead-sharded path-integral workloads (STRONG scaling over the same N GPUs; one exchange of 4 doubles per
        # sweep): the five-site + Ewald variant (enough work per GPU for the scaling target) and the primary single-site variant
        import copy
        for key, wl in (("pi_h2_five_bead_sharded", "pi_h2_five"), ("pi_h2_single_bead_sharded", "pi_h2")):
            a2 = copy.copy(args)
            a2.workload, a2.steps, a2.warmup = wl, 200, 20
            pi_res = run_pi(a2, embedded=True)
            if rank == 0:
                out[key] = {k: pi_res[k] for k in ("value", "unit", "n_gpus", "ms_per_step", "scaling", "config", "e2e", "pair_evals_per_sec", "roofline")}
    if rank == 0:
        print(json.dumps(out, default=_np_default))
    if world > 1:
        dist.destroy_process_group()


def run_pi(args, embedded=False):
    """BASELINE config 5: path-integral H2 cluster, 512 molecules x P=64 beads, beads sharded over the GPUs (strong scaling).
      value : potential sweeps per second with the beads resident in HBM: K calls of mpmc_pi_potential_allreduce (structure factor of the
              moved chunk, pair sweep over the local beads, reductions + per-bead assembly + cross-GPU exchange in one kernel, result
              copy — one CUDA graph) timed with CUDA events on the engine's stream, max over ranks.
      e2e   : MC moves per second through the C++ host mirror (SimulationControl::PI_nvt_mc over mpmc_host_run_sharded, one process per
              GPU, every rank replaying the same Rando stream as the reference's MPI ranks do): move generation on the host, coordinates
              host->device, sweep, sums device->host, PI_NVT_boltzmann_factor accept/reject, restore — wall clock of the step loop in
              C++.  An accepted move costs a second sweep (PathIntegral.cpp:148), so e2e also reports sweeps/s."""
    import tempfile
    import torch
    import torch.distributed as dist
    from mpmcxx_b200 import engine, pi, host_binding

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    if world > 1 and not embedded:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def shared_uid():
        uid = [engine.nccl_unique_id() if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(uid, src=0)
        return uid[0]

    P = args.beads
    five = args.workload == "pi_h2_five"
    tmpl, beads = W.pi_h2_cluster(n_side=8, P=P, five_site=five)
    lo, hi = pi.bead_range(P, rank, world)
    eng = engine.Engine(tmpl, beads=np.ascontiguousarray(beads[lo:hi]), device=local)
    if world > 1:
        eng.nccl_init(shared_uid(), rank, world)
    ext = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local))
    rs = np.random.RandomState(4242)            # the same stream on every rank
    starts = np.nonzero(np.diff(np.concatenate([[-1], tmpl.mol])))[0]
    ends = np.concatenate([starts[1:], [tmpl.n]])
    pos = beads.copy()
    u_first, _ = eng.pi_potential_allreduce(P)
    collective = eng.pi_collective()

    def move_and_sweep():
        """one molecule moves in every bead system (what a step changes on the device), then the sweep"""
        m = rs.randint(len(starts))
        a, b = int(starts[m]), int(ends[m])
        pos[:, a:b, :] += rs.normal(scale=0.02, size=(P, 1, 3))
        eng.update_sites_all_beads(a, pos[lo:hi, a:b, :])
        return eng.pi_potential_allreduce(P)[0]

    peak_tflops, _ = engine.probe_fp64_peak(local)
    for _ in range(args.warmup):
        move_and_sweep()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    l0 = eng.launches()
    # device-resident: K sweeps back to back over the beads as they sit in HBM, one event pair around all of them on the engine's
    # stream.  Before every sweep one molecule is marked as moved (mpmc_debug_mark_moved: no copy), so that the sweep does what it does
    # after a real move — the structure factor of the moved chunk is recomputed — and nothing is skipped.  (Timing each sweep on
    # its own after a host-side move measures the ranks' launch skew instead: the exchange inside a sweep waits for the slowest.)
    marks = [(int(starts[m]), int(ends[m] - starts[m])) for m in rs.randint(len(starts), size=args.steps)]
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for a, cnt in marks:
        eng.mark_moved(a, cnt)
        u_last = eng.pi_potential_allreduce(P)[0]
    e1.record(ext)
    e1.synchronize()
    t_dev_ms = e0.elapsed_time(e1)
    barrier()
    t_dev = max_over_ranks(t_dev_ms * 1e-3)
    launches = eng.launches() - l0
    # per-kernel-class durations (their own pass: event pairs disable the graph)
    eng.set_timing(True)
    for _ in range(args.steps):
        move_and_sweep()
    timing = eng.timing()
    eng.set_timing(False)
    out1 = eng.energy_all()[0]
    eng.close()
    # ---- e2e through the C++ mirror -----------------------------------------------------------------------------------
    e2e_steps = max(args.steps, 200)
    t = tmpl.copy()
    t.opts.update({"seed": "1", "numsteps": str(e2e_steps + 50), "corrtime": "1000000", "pqr_restart": "off", "pqr_output": "off"})   # no restart / final files in the timed loop
    d = tempfile.mkdtemp(prefix="mpmc_pi_bench_")
    inp = W.write_reference_job(t, d)
    uid2 = shared_uid()
    barrier()
    host_binding.run_sharded(inp, P, rank, world, local, uid2, max_steps=50, capacity=50)          # warm-up run (graph capture, allocations)
    barrier()
    log, summary = host_binding.run_sharded(inp, P, rank, world, local, shared_uid(), max_steps=e2e_steps, capacity=e2e_steps)
    loop_s, loop_sweeps = host_binding.last_stats()
    barrier()
    clk = clocks.stop()
    t_wall = max_over_ranks(loop_s)
    acc = float(log[:, 3].mean())
    pairs_per_sweep = out1["n_pair_evals"] * P
    pk = timing["pair"]
    kms = pk[0] / max(pk[1], 1)
    flop = (FLOP_PER_ES_PAIR if five else FLOP_PER_LJ_PAIR) * out1["n_pair_evals"] * (hi - lo)
    ach = flop / (kms * 1e-3) / 1e12 if kms > 0 else float("nan")
    res = {"metric": "mc_moves_per_sec", "value": args.steps / t_dev, "unit": "sweeps/s (one potential sweep over all beads = one trial move; the reference evaluates an accepted move a second time with unchanged coordinates — the host mirror reuses the answer it already has)",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "config5: path-integral H2 cluster, 512 molecules x %d beads, %s, beads sharded %d per GPU, one exchange of 4 doubles per sweep (see `collective`)"
                                  % (P, "five-site H2 + Ewald (N=2560 per bead)" if five else "single-site H2, rd_only (N=512 per bead)", hi - lo),
                      "l2": "inputs are < 1 MB per rank and L2-resident by nature; no flush (latency-bound path)", "pair_evals_per_sweep": pairs_per_sweep,
                      "collective": collective,
                      "potential_of_start_configuration_K": u_first, "potential_after_run_K": u_last},
           "clocks": clk,
           "e2e": {"value": e2e_steps / t_wall, "unit": "moves/s", "sweeps_per_s": loop_sweeps / t_wall, "sweeps_per_move": loop_sweeps / e2e_steps,
                   "h2d_bytes_per_step": 32 * int(ends[0] - starts[0]) * (hi - lo) * (2 - acc), "d2h_bytes_per_step": 48 * loop_sweeps / e2e_steps,
                   "ms_per_step": 1e3 * t_wall / e2e_steps, "acceptance": acc, "steps": e2e_steps,
                   "through": "C++ host mirror: SimulationControl::PI_nvt_mc via mpmc_host_run_sharded, one process per GPU"},
           "gpu_launches": int(launches), "pair_evals_per_sec": pairs_per_sweep * args.steps / t_dev,
           "roofline": {"kernel": "pair", "bound": "fp64", "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach / peak_tflops, "traffic": None,
                        "ms_per_launch": kms, "peak_source": "measured in this run (DFMA probe)",
                        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in timing.items() if v[1]}}}
    if embedded:
        return res
    if rank == 0:
        print(json.dumps(res, default=_np_default))
    if world > 1:
        dist.destroy_process_group()


def extra_workloads(engine, device, peak_tflops, steps=20):
    """Secondary single-GPU numbers: config 3 (bulk LJ argon, the pair kernel alone) and the Jacobi solver variant of config 4."""
    import torch
    res = {}
    for name, wl, solver in (("lj_argon_4096", "lj_argon", "jacobi10"), ("h2_framework_jacobi10", "h2_framework", "jacobi10")):
        s, desc = build_workload(wl, solver, 1.0)
        eng = engine.Engine(s, device=device)
        ext = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", device))
        for _ in range(3):
            eng.energy()
        eng.set_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(steps):
            eng.enqueue()
            out = eng.fetch()[0]
        e1.record(ext)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / steps
        tm = eng.timing()
        eng.set_timing(False)
        r = {"workload": desc, "moves_per_sec": 1e3 / ms, "ms_per_move": ms, "pair_evals_per_sec": out["n_pair_evals"] * 1e3 / ms,
             "kernel_ms_per_move": {k: v[0] / steps for k, v in tm.items() if v[1]}}
        if wl == "lj_argon":
            kms = tm["pair"][0] / tm["pair"][1]
            ach = FLOP_PER_LJ_PAIR * out["n_pair_evals"] / (kms * 1e-3) / 1e12
            r["pair_kernel"] = {"ms_per_launch": kms, "achieved_tflops": ach, "frac_of_fp64_peak": ach / peak_tflops,
                                "work": "%d pairs x %g flop" % (out["n_pair_evals"], FLOP_PER_LJ_PAIR)}
        else:
            npol = int(np.count_nonzero(s.alpha))
            kms = tm["dipole_sweep"][0] / tm["dipole_sweep"][1]
            ach = FLOP_PER_DIPOLE_PAIR * npol * (npol - 1) / (kms * 1e-3) / 1e12
            r["dipole_kernel"] = {"ms_per_launch": kms, "achieved_tflops": ach, "frac_of_fp64_peak": ach / peak_tflops}
        res[name] = r
        eng.close()
    return res


def uvt_through_mirror(device, steps=150):
    """BASELINE config 4 as the grand-canonical run it is named as (insert_probability 0.3: insertions, removals and displacements) through
    the C++ host mirror's System::mc over the engine — every insertion / removal re-sends the site table (mpmc_insert_sites /
    mpmc_remove_sites semantics), every rejection restores it: MC moves per second, wall clock of the step loop in C++."""
    import tempfile
    from mpmcxx_b200 import host_binding
    s = W.h2_framework(solver=W.SOLVER_GS_RANKED_PALMO, ensemble="uvt")
    s.opts.update({"h2_fugacity": "off", "pressure": "1.0", "seed": "7", "numsteps": str(steps), "corrtime": "1000000", "pqr_output": "off", "pqr_restart": "off"})
    inp = W.write_reference_job(s, tempfile.mkdtemp(prefix="mpmc_uvt_bench_"))
    host_binding.run(inp, max_steps=20, capacity=20)                    # warm-up (allocations, tables)
    log, summary = host_binding.run(inp, max_steps=steps, capacity=steps)
    loop_s, _ = host_binding.last_stats()
    kinds = np.bincount(log[:, 0].astype(int), minlength=3)
    return {"workload": "config4 as uVT (insert_probability 0.3, fugacity = pressure) through the C++ mirror's System::mc", "moves_per_sec": steps / loop_s,
            "ms_per_move": 1e3 * loop_s / steps, "steps": steps, "moves": {"insert": int(kinds[0]), "remove": int(kinds[1]), "displace": int(kinds[2])},
            "acceptance": float(log[:, 3].mean()), "final_N": float(summary[5])}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    t0 = time.perf_counter()
    cb = cpu_baseline(args.workload, args.solver, args.steps + args.warmup, budget_s=args.cpu_budget)
    _, desc = build_workload(args.workload, args.solver, args.scale)
    out = {"impl": "reference", "metric": "mc_moves_per_sec", "value": cb["value"], "unit": "moves/s", "n_gpus": env_int("WORLD_SIZE", 1),
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": desc},
           "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(out, default=_np_default))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="h2_framework", choices=["h2_framework", "lj_argon", "pi_h2", "pi_h2_five"])
    ap.add_argument("--beads", type=int, default=64, help="Trotter number of the path-integral workloads")
    ap.add_argument("--solver", default="gs_ranked_palmo", choices=sorted(SOLVERS))
    ap.add_argument("--scale", type=float, default=1.0, help="linear size factor of the synthetic system (1.0 = the named config)")
    ap.add_argument("--cpu-budget", type=float, default=10.0, help="seconds of CPU work for the secondary, reduced-size CPU sample (0 = skip); the primary CPU figure is always the full-size run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.workload.startswith("pi_"):
        run_pi(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
