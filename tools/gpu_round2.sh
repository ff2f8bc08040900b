#!/bin/bash
# One gpurun call (1 GPU): parity tests, bench line (+ reference arm), ncu launch list of the bench command and full
# captures of the dominant kernels.   gpurun --timeout 1800 -- 'bash tools/gpu_round2.sh TAG'
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
timeout 400 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench_reference.json
# launch list of the bench command (per-launch times are cold-cache + serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
bash tools/ncu_round.sh $TAG c2:k_q1_gather c5:k_q2_gather c3:k_dg_gather c2-elem:k_q1_gather c5-qp:k_q2_qp c2-qp:k_q1_gather_qp > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fv_march -s 3 -c 1 -f -o $OUT/${TAG}_fv_march \
  python tools/fv_once.py > $OUT/${TAG}_ncu_fv.log 2>&1
python tools/ncu_digest.py $OUT/${TAG}_fv_march.ncu-rep 18 > $OUT/${TAG}_digest_fv_march.txt 2>&1
rm -f $OUT/${TAG}_fv_march.ncu-rep
ls -la $OUT/${TAG}_*
