#!/bin/bash
# tools/launch_list.sh <tag> [steps] -- ncu launch list (gpu__time_duration per launch) of the timed window of bench.py
tag=${1:-r02_final}; steps=${2:-20}
timeout ${NCU_LIMIT:-1400} ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps $steps --warmup 3 --cpu-steps 0 --batch-scenes 0 --free-steps 0 --kernel-steps 1 --profile-range timed > gpurun_out/${tag}_launches.log 2>&1
tail -1 gpurun_out/${tag}_launches.log | cut -c1-300
