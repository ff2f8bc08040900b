#!/bin/bash
# One `gpurun --gpus 8` call: the driver's 1 -> 8 scaling line of bench.py plus the other multi-GPU rows.
#   gpurun --gpus 8 --timeout 1200 -- 'bash tools/scale_round.sh TAG'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/${TAG}_scale.jsonl
run() { # N, extra args
  local n=$1; shift
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 "$@" 2>>$OUT/${TAG}_scale.err | tail -1 >> $OUT/${TAG}_scale.jsonl
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n "$@" 2>>$OUT/${TAG}_scale.err | tail -1 >> $OUT/${TAG}_scale.jsonl
  fi
}
for n in 1 2 4 8; do run $n --steps 100 --warmup 5 --no-cpu-baseline; done
for n in 1 2 4 8; do run $n --workload c5 --steps 50; done
for n in 1 8; do run $n --workload c4-weak --steps 100; done
for n in 2 8; do run $n --workload c4 --steps 100; done
run 8 --workload c2-halo --steps 30
python - <<PY
import json
for line in open("$OUT/${TAG}_scale.jsonl"):
    try:
        d = json.loads(line)
    except ValueError:
        print("??", line[:200]); continue
    print(d["n_gpus"], d["metric"][:60], "%.4g %s" % (d["value"], d["unit"]), "ms/step %.4f" % d["ms_per_step"], d["config"].get("partition", ""))
PY
