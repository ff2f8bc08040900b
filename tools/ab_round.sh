#!/bin/bash
# A/B round on one GPU: every lib variant under dune-gdt_b200/lib/ab plus the default, FV / C2-elem / C3 / C5 timings
TAG=${1:-ab}
OUT=gpurun_out
mkdir -p $OUT
for L in default $(ls dune-gdt_b200/lib/ab/*.so 2>/dev/null); do
  if [ "$L" = default ]; then unset GDTB_LIB; else export GDTB_LIB=$PWD/$L; fi
  echo "== $L" >> $OUT/${TAG}_ab.jsonl
  timeout 300 python tools/fv_ab.py >> $OUT/${TAG}_ab.jsonl 2>> $OUT/${TAG}_ab.err
  timeout 600 python tools/bench_configs.py ${AB_CONFIGS:-c2-elem c3 c5} >> $OUT/${TAG}_ab.jsonl 2>> $OUT/${TAG}_ab.err
done
cat $OUT/${TAG}_ab.jsonl
