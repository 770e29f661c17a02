#!/bin/bash
# usage: [PI_ONLY=1] [FULL=1] tools/scale_run.sh N  — run the three bench modes on N GPUs of this box, one JSON line each into gpurun_out/scale_N.jsonl
N=$1
OUT=gpurun_out/scale_${N}.jsonl
: > $OUT
run() {
  if [ "$N" = "1" ]; then timeout 400 python bench.py --gpus 1 "$@" 2>gpurun_out/scale_err_${N}.log | grep -E '^\{' >> $OUT
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" 2>gpurun_out/scale_err_${N}.log | grep -E '^\{' >> $OUT; fi
}
[ "${PI_ONLY:-0}" = "1" ] || run --steps 10 --warmup 3 --no-cpu-baseline --no-extra
run --workload pi_h2_five --steps 100 --warmup 10
[ "${FULL:-0}" = "1" ] && run --workload pi_h2 --steps 200 --warmup 20
python - <<PY
import json
for l in open("$OUT"):
    d = json.loads(l)
    print(d["n_gpus"], d["config"]["workload"][:40], "value %.1f" % d["value"], "e2e %.1f" % d["e2e"]["value"], "ms/step %.3f" % d["ms_per_step"])
PY
