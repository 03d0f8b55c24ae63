#!/bin/bash
# Everything profiles/r02_* comes from, in ONE gpurun call on one B200:
#   gpurun --timeout 2400 -- bash scripts/r02_profile.sh
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest.log 2>&1; tail -3 gpurun_out/r02_final_pytest.log
python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err
python scripts/variant_bench.py --out gpurun_out/r02_variant_bench.json > gpurun_out/r02_variant_bench.log 2>&1
python scripts/sunz_bench.py > gpurun_out/r02_sunz_bench.log 2>&1
bash scripts/query_knobs.sh > gpurun_out/r02_query_knobs.txt 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'nn_fixed_query_rect_kernel<\(bool\)1>|sunz_fast|abi_calibrate_lane|nn_source_points' -s 12 -c 5 -o gpurun_out/r02_step_kernels $CMD > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'nn_gather_bulk|bil_sample|fornav_kernel' -c 4 -o gpurun_out/r02_stage_kernels $CMD > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
python -c "
import json
d=json.load(open('gpurun_out/r02_final_bench_n1.json'))
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['parity'], d['clocks'])
for k,v in d['stages'].items(): print(k, round(v['ms'],3), round(v['gbs'],1))
"
cat gpurun_out/r02_query_knobs.txt
