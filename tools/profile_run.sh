#!/bin/bash
# Round profile capture (run on the GPU box through gpurun): a plain bench run first, then the ncu launch list of the same command,
# then one `--set full` capture of each dominant kernel.  Outputs under gpurun_out/ with the tag $1.
TAG=${1:-r01b}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra"
# ncu serialises kernel launches; the Gauss-Seidel pipeline is two kernels that must run side by side, so under ncu the engine is
# switched to its single-launch fallback (MPMC_GS_FUSED=1: updaters inside the solver's launch, one CTA per SM, ~25 % slower)
$CMD > gpurun_out/plain_${TAG}.log 2> gpurun_out/plain_${TAG}.err || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.err; exit 1; }
MPMC_GS_FUSED=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
MPMC_GS_FUSED=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_gs_pipeline|k_contract_parts|k_field_parts|k_pair_sweep|k_rank_min_parts|k_field_recip|k_gs_inverse|k_gs_near' -c 16 \
    -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out/prof_${TAG}.ncu-rep
