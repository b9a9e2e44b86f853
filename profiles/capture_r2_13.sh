set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_13_bench.json 2> gpurun_out/r2_13_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-scan-probe"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_13_launches.csv $CMD > gpurun_out/r2_13_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:imi_scan_kernel -s 4 -c 1 -f -o gpurun_out/r2_13_imi_scan $CMD > gpurun_out/r2_13_ncu1.log 2>&1
ls -la gpurun_out/ | grep r2_13
