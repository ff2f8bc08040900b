#!/bin/bash
# One `gpurun --gpus 8` call: bench.py at N = 8 (the communicating rows ride along under extra.multi_gpu), the strong C5 / C3
# rows at N = 1 and 8, and the 2-GPU pytest.   gpurun --gpus 8 --timeout 1500 -- 'bash tools/scale8.sh TAG'
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
tr() { local n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"; }
tr 8 bench.py --gpus 8 > $OUT/${TAG}_bench_8gpu.json 2> $OUT/${TAG}_bench_8gpu.err
tail -c 400 $OUT/${TAG}_bench_8gpu.json
: > $OUT/${TAG}_rows.jsonl
for w in c5 c3; do
  timeout 300 python bench.py --gpus 1 --workload $w --steps 50 2>>$OUT/${TAG}_rows.err | tail -1 >> $OUT/${TAG}_rows.jsonl
  tr 8 bench.py --gpus 8 --workload $w --steps 50 2>>$OUT/${TAG}_rows.err | tail -1 >> $OUT/${TAG}_rows.jsonl
done
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_parallel_gpu.py -m gpu -q > $OUT/${TAG}_pytest_2gpu.log 2>&1
tail -3 $OUT/${TAG}_pytest_2gpu.log
python - <<PY
import json
for line in open("$OUT/${TAG}_rows.jsonl"):
    try:
        d = json.loads(line)
    except ValueError:
        print("??", line[:200]); continue
    print(d["n_gpus"], d["metric"][:70], "%.4g %s" % (d["value"], d["unit"]), "ms/step %.4f" % d["ms_per_step"])
PY
