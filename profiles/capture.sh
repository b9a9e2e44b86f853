#!/bin/bash
# Run on the GPU box (gpurun): launch list of one bench command + a full ncu capture of one kernel.
#   profiles/capture.sh <tag> [kernel-regex]
# Outputs gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_<kernel>.ncu-rep; summarise here with
# profiles/summarise.py and commit the CSV summaries under profiles/.
set -u
TAG=${1:-rX}
KERN=${2:-imi_scan_kernel}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-scan-probe"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${KERN} -s 4 -c 2 \
    -f -o gpurun_out/${TAG}_${KERN} $CMD > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out/
