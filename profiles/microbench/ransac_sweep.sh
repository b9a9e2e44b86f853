#!/bin/bash
# RANSAC schedule sweep on the headline step (env overrides of DESIGN.md section 5); prints stage_ms per setting.
run() { env "$@" python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --no-scan-probe 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['value']), d['stage_ms'])"; }
run X=1
run MLC_RANSAC_FIRST_HYP=16
run MLC_RANSAC_SLOTS=32 MLC_RANSAC_FIRST_HYP=32
run MLC_RANSAC_SLOTS=32 MLC_RANSAC_FIRST_HYP=24
run MLC_RANSAC_SLOTS=32 MLC_RANSAC_FIRST_HYP=16
run MLC_RANSAC_GROUPS=1
run MLC_RANSAC_GROUPS=2
run MLC_RANSAC_GROUPS=4
run MLC_RANSAC_GROUPS=5
run MLC_RANSAC_GROUPS=4 MLC_RANSAC_FIRST_HYP=16
