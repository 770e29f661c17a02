"""Developer tool: where does a Gauss-Seidel sweep spend its time?  Prints per-block SM-clock stamps of the solver CTA (start,
helpers delivered, fold done, right-hand side done, matrix-vector product of thread 191 start / end, barrier after it, panel
published) and the phases of one updater warp for the first sweep of one energy() on config 4 (run on the GPU box)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpmcxx_b200 import engine, workloads as W

s = W.h2_framework(solver={"polar_gs": "on", "polar_max_iter": "2"}, ensemble="nvt")
e = engine.Engine(s)
L = engine.lib()
e.energy()
DBG = int(sys.argv[1]) if len(sys.argv) > 1 else 0
L.mpmc_debug_gs_profile(e.h, 1 | DBG, None, 0, None)
e.energy()
MAXB = 256
buf = np.zeros(4 * 8 * MAXB, dtype=np.int64)
nb = C.c_int()
L.mpmc_debug_gs_profile(e.h, 0x100 | DBG, buf.ctypes.data_as(C.c_void_p), MAXB, C.byref(nb))
nb = nb.value
sol = buf[:8 * MAXB].reshape(MAXB, 8)[:nb]
upd = buf[8 * MAXB:16 * MAXB].reshape(MAXB, 8)[:nb]
handed = buf[16 * MAXB:32 * MAXB].reshape(MAXB, 16)[:nb]
cols = [0, 4, 1, 1, 1, 6, 6, 2]
labels = ["start", "helpers' sums in", "rhs done, inverse in", "(same)", "matvec start (thread 191)", "matvec end (thread 191)", "(same)", "panel sent + written back"]
rel = (sol[1:-1][:, cols] - sol[1:-1, :1]).astype(np.float64)
print("solver: SM-clock cycles after the start of the block, median over %d blocks (~1.9 GHz)" % (nb - 2))
for i, nme in enumerate(labels):
    print("  %-28s %8.0f" % (nme, np.median(rel[:, i])))
per_blk = np.diff(sol[:, 0])
print("  block period             mean %8.0f cycles" % per_blk.mean())
for b in (1, 2, 50, 100):
    print(" blk", b, "solver clk", (sol[b, [0, 4, 1, 1, 1, 6, 6, 2]] - sol[b, 0]).tolist(), "(start, helpers in, fold done, rhs done, thread 191 matvec start, end, matvec barrier, walk end) next start", int(sol[b + 1, 0] - sol[b, 0]))
per = np.diff(sol[:, 0])
worst = np.argsort(per)[-5:][::-1]
print("longest block periods:", [(int(b), int(per[b])) for b in worst], " total sweep cycles (first..last block start):", int(sol[nb - 1, 0] - sol[0, 0]))
u = upd[:, :4].astype(np.float64)
ok = (u[:, 2] > 0) & (u[:, 0] > 0)
polled = ok & (u[:, 3] > 0)
print("updater warp (cta 0 warp 0), NANOSECONDS per panel (mean over %d panels): top -> next panel's changes in shared memory %.0f, contraction (+ hand-over when due) %.0f; its period %.0f" % (
    ok.sum(), (u[ok, 1] - u[ok, 0]).mean(), (u[ok, 2] - u[ok, 1]).mean(), np.diff(u[ok][:, 0]).mean()))
print("   panels it had to poll for: %d of %d; there: waiting for the flag %.0f, fence + load of the changes %.0f" % (
    polled.sum(), ok.sum(), (u[polled, 3] - u[polled, 0]).mean() if polled.any() else 0, (u[polled, 1] - u[polled, 3]).mean() if polled.any() else 0))
wait = sol[:, 5] >> 16
late = sol[:, 5] & 0xffff
idx = np.argsort(wait)[-12:][::-1]
print("loader warp: cycles spent waiting for the updaters' flags of the NEXT block's rows, worst blocks (block, cycles, chunk it waited for last):")
print("  ", [(int(b), int(wait[b]), int(late[b])) for b in idx])
print("   blocks with any wait: %d of %d; total wait %d cycles of %d" % ((wait > 0).sum(), nb, wait.sum(), int(sol[nb - 1, 0] - sol[0, 0])))

# the cluster's round trip in nanoseconds (globaltimer): panel b sent by the solver -> arrived at helper 0 -> helper's sums computed ->
# helper's iteration done (delivered) -> all helpers' sums seen by the solver (block b + 1)
sent = upd[:-1, 6]; arrived = upd[:-1, 4]; computed = upd[:-1, 5]; done = upd[:-1, 7]; seen = sol[1:, 7]; woke = sol[1:, 3]
ok2 = (sent > 0) & (arrived > 0) & (seen > 0) & (wait[1:] == 0) & (wait[:-1] == 0)
print("cluster round trip, ns after the panel was sent (median over %d undisturbed blocks): arrived at helper 0 %+.0f, its sums computed %+.0f, LAST helper's iteration done %+.0f, solver thread 0 woke %+.0f, all solver threads past the barrier %+.0f" % (
    ok2.sum(), np.median((arrived - sent)[ok2]), np.median((computed - sent)[ok2]), np.median((done - sent)[ok2]), np.median((woke - sent)[ok2]), np.median((seen - sent)[ok2])))

# per chunk: how long after the panel it had to wait for (panel c-5, published when the solver sent it) were its rows handed over
pub = upd[:, 6]                      # solver's globaltimer when panel b was sent
lat = np.full(handed.shape, np.nan)
for cb in range(5, nb):
    ok3 = handed[cb] > 0
    lat[cb, ok3] = handed[cb, ok3] - pub[cb - 5]
fl = lat[5:].ravel(); fl = fl[np.isfinite(fl)]
print("hand-over of a chunk's rows, ns after the last panel it needed was sent: median %.0f, 90%% %.0f, 99%% %.0f, max %.0f  (the solver needs them after ~3 block periods = %.0f ns)" % (
    np.median(fl), np.percentile(fl, 90), np.percentile(fl, 99), fl.max(), 3 * np.median(np.diff(pub[pub > 0]))))
bywarp = {}
for cb in range(5, nb):
    for k in range(16):
        ch = cb * 16 + k
        if np.isfinite(lat[cb, k]): bywarp.setdefault(ch // 140, []).append(lat[cb, k])
print("  by warp index of the updater CTA (warp = chunk // 140; scheduler = warp % 4): " + ", ".join("%d:%.0f" % (w, np.median(v)) for w, v in sorted(bywarp.items())))
print("  hand-over latency by block (median over its chunks, ns):", [(cb, int(np.nanmedian(lat[cb]))) for cb in list(range(5, 30)) + list(range(60, 70)) + list(range(130, nb))])
print("  solver: ns between consecutive panel sends, blocks 0..40:", np.diff(pub[:41]).tolist())

lag = u[:, 1] - pub
print("  updater warp (cta 0 warp 0): ns between the solver sending panel b and this warp starting its contraction, panels 0..143 step 8:", [int(x) for x in lag[::8]])
