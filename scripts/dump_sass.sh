#!/bin/bash
# SASS of the hot kernels of the built library -> profiles/<round>_sass_<kernel>.txt (instruction text only, the
# encoding lines are dropped) plus a mnemonic census that shows which kernels use the bulk-copy engine (UBLKCP) and
# mbarriers (SYNCS), 128-bit global accesses (LDG.E.128 / STG.E.128) and the FP64 pipe.
#   bash scripts/dump_sass.sh r02
set -e
ROUND=${1:-r02}
LIB=satpy_b200/libsatpy_b200.so
declare -A K=(
  [nn_fixed_query_rect_fused]=_Z26nn_fixed_query_rect_kernelILb1EEv12GeosQueryArgPyPKfjji
  [nn_source_points_fast]=_Z23nn_source_points_kernelILb1EEv6SrcArgPKhPhP7double4P6float4PiPKiddi8GeosFast
  [bil_sample_bulk]=_Z22bil_sample_bulk_kernelPKfPfPKiPKdS5_lf
  [sunz_fast_f32]=_Z16sunz_fast_kernelILi0ELb1EEv11SunzFastArg7SunzDev
  [sunz_fast_fused_abi]=_Z16sunz_fast_kernelILi1ELb1EEv11SunzFastArg7SunzDev
  [abi_calibrate_lane]=_Z25abi_calibrate_lane_kernelILb0EEvPKsPfl9AbiCalDev
  [abi_calibrate_bulk]=_Z25abi_calibrate_bulk_kernelILb0EEvPKsPfl9AbiCalDev
  [nn_gather_bulk]=_Z21nn_gather_bulk_kernelILb0EEvPKjPjPKiljill
  [nn_gather_direct]=_Z16nn_gather_kernelIjEvPKT_PS0_PKilillS0_
)
CENSUS=profiles/${ROUND}_sass_census.txt
echo "cuobjdump -sass $LIB (sm_100a); per kernel: instructions, then the count of selected mnemonics" > $CENSUS
for name in "${!K[@]}"; do
  out=profiles/${ROUND}_sass_${name}.txt
  cuobjdump -sass -fun "${K[$name]}" $LIB 2>/dev/null | grep -v '^\s*/\* 0x' | sed -E 's@\s*/\* 0x[0-9a-f]+ \*/\s*$@@' > $out
  n=$(grep -c '^\s*/\*[0-9a-f]\{4\}\*/' $out || true)
  echo "" >> $CENSUS
  echo "$name (${K[$name]}): $n instructions" >> $CENSUS
  for m in UBLKCP SYNCS UTMALDG UTMASTG CCTL VIMNMX LOP3 "LDG.E.128" "LDG.E.64" "STG.E.128" "STG.E.NA.128" "LDG.E.NA.64" LDS STS MUFU DFMA DMUL F2F I2F "FFMA" BSSY; do
    c=$(grep -c "$m" $out || true)
    [ "$c" != "0" ] && echo "    $m: $c" >> $CENSUS
  done
done
cat $CENSUS
