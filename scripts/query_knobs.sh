#!/bin/bash
# A/B of the query kernel's launch shape (warps per CTA x tiles per warp) on config 5, one B200:
#   gpurun -- bash scripts/query_knobs.sh > gpurun_out/query_knobs.txt
for w in 4 2 1; do for t in 16 64; do
  B200_NN_WARPS_PER_CTA=$w B200_NN_TILES_PER_WARP=$t python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-stages 2>/dev/null |
    python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('warps/CTA $w tiles/warp $t: step', round(d['ms_per_step'],3), 'ms, query', round(d['stages']['query_gather']['ms'],3), 'ms')"
done; done
