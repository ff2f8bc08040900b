#!/bin/bash
# ncu --set full captures (one launch each) of the kernels named on the command line, e.g.
#   gpurun --timeout 900 -- 'bash tools/ncu_round.sh r02s c2-elem:k_q1_gather c3:k_dg_gather c5:k_q2_gather'
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for spec in "$@"; do
  cfg=${spec%%:*}; kern=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s 1 -c 1 -f -o $OUT/${TAG}_${cfg} \
    python tools/bench_configs.py $cfg > $OUT/${TAG}_ncu_${cfg}.log 2>&1
  tail -2 $OUT/${TAG}_ncu_${cfg}.log
  # digest on the box (counters, stall reasons, source hot spots); the .ncu-rep itself stays behind unless KEEP_REP=1
  # (gpurun copies at most 64 MiB back)
  python tools/ncu_digest.py $OUT/${TAG}_${cfg}.ncu-rep 18 > $OUT/${TAG}_digest_${cfg}.txt 2>&1
  [ -n "$KEEP_REP" ] || rm -f $OUT/${TAG}_${cfg}.ncu-rep
done
ls -la $OUT/${TAG}_*
