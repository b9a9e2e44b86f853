#!/bin/bash
# experiment driver (scratch): bench each library variant
cp maplab_b200/libmaplab_lc_b200.so /tmp/base.so
for v in base "$@"; do
  if [ "$v" = base ]; then cp /tmp/base.so maplab_b200/libmaplab_lc_b200.so; else cp gpurun_variants/lib_$v.so maplab_b200/libmaplab_lc_b200.so; fi
  timeout 120 python -m pytest tests/test_gpu_knn.py -m gpu -x -q 2>&1 | tail -1
  timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-scan-probe > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/var_$v.json').read().strip().splitlines()[-1]); print('$v', round(d['value']), round(d['e2e']['value']), d['stage_ms'], round(d['roofline']['frac'],4))"
done
cp /tmp/base.so maplab_b200/libmaplab_lc_b200.so
