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
buf = np.zeros(2 * 8 * MAXB, dtype=np.int64)
nb = C.c_int()
L.mpmc_debug_gs_profile(e.h, 0 | DBG, buf.ctypes.data_as(C.c_void_p), MAXB, C.byref(nb))
nb = nb.value
sol = buf[:8 * MAXB].reshape(MAXB, 8)[:nb]
upd = buf[8 * MAXB:].reshape(MAXB, 8)[:nb]
cols = [0, 4, 1, 5, 3, 6, 7, 2]
labels = ["start", "helpers in", "fold done", "rhs done", "matvec start (thread 191)", "matvec end (thread 191)", "matvec barrier", "panel published"]
rel = (sol[1:-1][:, cols] - sol[1:-1, :1]).astype(np.float64)
print("solver: SM-clock cycles after the start of the block, median over %d blocks (~1.9 GHz)" % (nb - 2))
for i, nme in enumerate(labels):
    print("  %-28s %8.0f" % (nme, np.median(rel[:, i])))
per_blk = np.diff(sol[:, 0])
print("  block period             mean %8.0f cycles" % per_blk.mean())
for b in (1, 2, 50, 100):
    print(" blk", b, "solver clk", (sol[b, [0, 4, 1, 5, 3, 6, 7, 2]] - sol[b, 0]).tolist(), "(start, helpers in, fold done, rhs done, thread 191 matvec start, end, matvec barrier, walk end) next start", int(sol[b + 1, 0] - sol[b, 0]))
per = np.diff(sol[:, 0])
worst = np.argsort(per)[-5:][::-1]
print("longest block periods:", [(int(b), int(per[b])) for b in worst], " total sweep cycles (first..last block start):", int(sol[nb - 1, 0] - sol[0, 0]))
u = upd[:, :4].astype(np.float64)
ok = (u[:, 3] > 0) & (u[:, 0] > 0)
du = np.diff(u[ok], axis=1)
print("updater warp (cta %d warp 0): wait for panel %.0f, dmu load + compute %.0f, ordering wait + atomics + fence + flag %.0f cycles (mean over %d panels); its period %.0f" % (
    4, du[:, 0].mean(), du[:, 1].mean(), du[:, 2].mean(), ok.sum(), np.diff(u[ok][:, 0]).mean()))

