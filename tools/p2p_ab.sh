#!/bin/bash
# usage: tools/p2p_ab.sh N — the bead-sharded PI workload on N GPUs with the fused peer-memory all-reduce and with ncclAllReduce
N=$1
for mode in 1 0; do
  MPMC_PI_P2P=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload pi_h2_five --steps 100 --warmup 10 2>gpurun_out/p2p_err_${N}_${mode}.log | grep -E '^\{' > gpurun_out/p2p_${N}_${mode}.json
  python - <<PY
import json
d = json.load(open("gpurun_out/p2p_${N}_${mode}.json"))
c = d["config"]
print("P2P=$mode N=$N value %.1f ms/step %.4f e2e %.1f | %s | U0 %.12f U_end %.12f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], c["collective"][:40], c["potential_of_start_configuration_K"], c["potential_after_run_K"]))
PY
done
tail -3 gpurun_out/p2p_err_${N}_1.log
