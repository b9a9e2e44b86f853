set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-scan-probe"
ncu --set full --clock-control none --import-source on -k regex:projection_tmem_kernel -c 1 -f -o gpurun_out/r2_10_projection python profiles/microbench/projection_only.py 16000000 > gpurun_out/r2_10_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:covis_kernel -s 4 -c 1 -f -o gpurun_out/r2_10_covis $CMD > gpurun_out/r2_10_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kd_search_kernel -s 6 -c 1 -f -o gpurun_out/r2_10_kd $CMD > gpurun_out/r2_10_ncu4.log 2>&1
ls -la gpurun_out/ | grep r2_10
