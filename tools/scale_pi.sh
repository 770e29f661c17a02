#!/bin/bash
# Developer tool (8-GPU box): strong scaling of the bead-sharded path-integral workloads over 1/2/4/8 GPUs, one JSON line per run.
# usage: tools/scale_pi.sh <out.jsonl> [workload ...]
out=$1; shift
wls=${@:-pi_h2_five pi_h2}
: > $out
for w in $wls; do
  for n in 1 2 4 8; do
    if [ $n -eq 1 ]; then
      python bench.py --workload $w --steps 200 --warmup 20 2>/dev/null | tail -1 >> $out
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --workload $w --steps 200 --warmup 20 2>/dev/null | tail -1 >> $out
    fi
  done
done
python - "$out" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip().startswith("{")]
base = {}
for d in rows:
    w = d["config"]["workload"][:60]
    if d["n_gpus"] == 1: base[w] = (d["value"], d["e2e"]["value"])
    b = base.get(w, (d["value"], d["e2e"]["value"]))
    print("%-62s N=%d  %.4f ms/sweep  %8.0f sweeps/s (eff %.2f)   e2e %8.0f moves/s (eff %.2f)  %s" % (w, d["n_gpus"], d["ms_per_step"], d["value"], d["value"] / (b[0] * d["n_gpus"]), d["e2e"]["value"], d["e2e"]["value"] / (b[1] * d["n_gpus"]), d["config"]["collective"][:24]))
PY
