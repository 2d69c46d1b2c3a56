#!/bin/bash
# tools/profile_round.sh <tag> -- the ncu evidence kept under profiles/ for one build (run on the GPU box via gpurun):
#   * launch list of the timed window of bench.py (gpu__time_duration per launch, clocks not locked)
#   * one --set full capture per kernel of interest, inside the same window
# Outputs go to gpurun_out/<tag>_*; summarise with profiles/ncu_summarize.py.
tag=${1:-r02_final}
NCU="ncu --profile-from-start off --clock-control none"
B="python bench.py --steps ${STEPS:-6} --warmup 3 --cpu-steps 0 --batch-scenes 0 --free-steps 0 --kernel-steps 1 --profile-range timed"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_launches.log 2>&1
for k in k_ilu_factor_strips k_trisolve k_query k_refit_inc k_canon_small k_imp_flow k_xdot_walk; do
  $NCU --set full --import-source on -k regex:^$k --launch-skip 2 -c 2 -f -o gpurun_out/${tag}_$k $B > gpurun_out/${tag}_$k.log 2>&1
done
ls -la gpurun_out | grep ${tag}_ | awk '{print $5, $9}'
