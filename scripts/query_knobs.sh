#!/bin/bash
# A/B of the query kernel's launch shape on config 5, one B200 (warps per CTA, tiles per warp):
#   gpurun -- bash scripts/query_knobs.sh > gpurun_out/query_knobs.txt
run() {
  B200_NN_WARPS_PER_CTA=$1 B200_NN_TILES_PER_WARP=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-stages 2>/dev/null |
    python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('warps/CTA $1 tiles/warp $2: step', round(d['ms_per_step'],3), 'ms, query', round(d['stages']['query_gather']['ms'],3), 'ms')"
}
if [ $# -gt 0 ]; then run "$@"; else run 1 64; run 2 64; run 4 64; run 1 16; run 1 128; fi
