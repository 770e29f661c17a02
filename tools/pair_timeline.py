"""Developer tool (GPU box): per-warp timeline of one pair sweep (mpmc_debug_pair_profile).
usage: python tools/pair_timeline.py [lj|pi8|h2fw]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mpmcxx_b200 import engine, workloads as W

which = sys.argv[1] if len(sys.argv) > 1 else "lj"
if which == "lj":
    e = engine.Engine(W.lj_argon()); run = e.energy
elif which == "pi8":
    t, b = W.pi_h2_cluster(P=8, five_site=True); e = engine.Engine(t, beads=b); run = e.energy_all
else:
    e = engine.Engine(W.h2_framework()); run = e.energy
L = engine.lib()
for _ in range(5):
    run()
engine._ck(L.mpmc_debug_pair_profile(e.h, 1, None, 0, None))
for _ in range(3):
    run()
cap = 148 * 32
out = np.zeros((cap, 4), dtype=np.int64)
n = C.c_int()
engine._ck(L.mpmc_debug_pair_profile(e.h, 0, out.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
p = out[: n.value]
t0 = p[:, 0].min()
ent, first, last = p[:, 0] - t0, p[:, 1] - t0, p[:, 2] - t0
items, sm = p[:, 3] & 0xffffffff, p[:, 3] >> 32
q = lambda a: "min %6d  10%% %6d  50%% %6d  90%% %6d  max %6d" % (a.min(), *np.percentile(a, [10, 50, 90]).astype(int), a.max())
print("warps %d, items per warp: %s" % (len(p), np.bincount(items.astype(int)).tolist()))
print("ns after the first warp's entry")
print("  entry               ", q(ent))
print("  first sites loaded  ", q(first[p[:, 1] > 0]), " (entry -> loaded: median %d)" % np.median((first - ent)[p[:, 1] > 0]))
print("  last item summed    ", q(last[p[:, 2] > 0]))
busy = (last - ent)[p[:, 2] > 0]
print("  entry -> last item  ", q(busy))
per_sm = {}
for s_, a, b in zip(sm, ent, last):
    lo, hi = per_sm.get(int(s_), (1 << 60, 0)); per_sm[int(s_)] = (min(lo, a), max(hi, b))
spans = np.array([hi - lo for lo, hi in per_sm.values()]); ends = np.array([hi for lo, hi in per_sm.values()]); starts = np.array([lo for lo, hi in per_sm.values()])
print("  per SM: first entry ", q(starts), "| last end", q(ends), "| span", q(spans))
# second CTA of an SM against the first
print("  warps by 8 (CTA) entry spread inside an SM: median %d ns" % np.median([np.ptp(ent[sm == s_]) for s_ in per_sm]))
# who ends early?  by warp index inside the CTA, and by the CTA's order of arrival on its SM
gw = np.arange(len(p)); wic = gw % 8; cta = gw // 8
print("  end by warp-in-CTA (median ns):", [int(np.median(last[wic == w])) for w in range(8)])
order = np.zeros(len(p), dtype=int)
for s_ in per_sm:
    ctas = sorted(set(cta[sm == s_]), key=lambda c_: ent[cta == c_].min())
    for o, c_ in enumerate(ctas):
        order[cta == c_] = o
print("  end by CTA arrival order on the SM (median ns):", [int(np.median(last[order == o])) for o in range(order.max() + 1)], " CTAs per SM:", np.bincount(np.bincount(sm.astype(int))[np.bincount(sm.astype(int)) > 0] // 8).tolist())
for s_ in list(per_sm)[:3]:
    m = sm == s_
    print("  SM %d: (warp-in-CTA, order, end us)" % s_, sorted((int(order[i]), int(wic[i]), round(last[i] / 1000.0, 1)) for i in np.where(m)[0]))
print("  end vs item index (median ns per 8th of the list):", [int(np.median(last[(gw * 8) // len(p) == o])) for o in range(8)])
