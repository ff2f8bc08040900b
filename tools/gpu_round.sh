#!/bin/bash
# One gpurun call: parity tests, bench line, other configs, ncu launch list + full captures of the dominant kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench.json
timeout 300 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench_reference.json
timeout 600 python tools/bench_configs.py > $OUT/${TAG}_configs.jsonl 2> $OUT/${TAG}_configs.err
cat $OUT/${TAG}_configs.jsonl
# launch list of the bench command (per-launch times are cold-cache + serialised: compare shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
# full captures (one launch each)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_q1_gather -s 3 -c 1 -f -o $OUT/${TAG}_q1_gather \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_q1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fv_march -s 3 -c 1 -f -o $OUT/${TAG}_fv_march \
  python tools/fv_once.py > $OUT/${TAG}_ncu_fv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_q2_gather -s 1 -c 1 -f -o $OUT/${TAG}_q2_gather \
  python tools/bench_configs.py c5 > $OUT/${TAG}_ncu_q2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dg_gather -s 1 -c 1 -f -o $OUT/${TAG}_dg_gather \
  python tools/bench_configs.py c3 > $OUT/${TAG}_ncu_dg.log 2>&1
ls -la $OUT
