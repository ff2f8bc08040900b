#!/bin/bash
# A/B of the CG Q2 gather kernel (C5): parity of the long-line cases first, then timed runs "name|ENV=.. ENV=..|lib"
# (lib: default or a file under dune-gdt_b200/lib/ab).   gpurun --timeout 900 -- 'bash tools/q2_ab.sh TAG [runs...]'
TAG=${1:-q2ab}; shift
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_full_size_oracle_gpu.py tests/test_full_size_gpu.py tests/test_coefficients_gpu.py -m gpu -x -q -k "q2 or Q2" > $OUT/${TAG}_pytest.log 2>&1
tail -5 $OUT/${TAG}_pytest.log
export GDTB_REPS=${GDTB_REPS:-30}
J=$OUT/${TAG}_ab.jsonl
: > $J
RUNS=("$@")
if [ ${#RUNS[@]} -eq 0 ]; then
  RUNS=("no-ln|GDTB_Q2_NO_LN=1|default" "default||default")
  for L in $(ls dune-gdt_b200/lib/ab/*.so 2>/dev/null); do RUNS+=("$(basename $L .so)||$L"); done
fi
for R in "${RUNS[@]}"; do
  IFS='|' read -r NAME ENVS LIB <<< "$R"
  echo "== $NAME [$ENVS] $LIB" >> $J
  if [ "$LIB" = default ] || [ -z "$LIB" ]; then LIBENV=""; else LIBENV="GDTB_LIB=$PWD/$LIB"; fi
  env $ENVS $LIBENV timeout 300 python tools/bench_configs.py ${Q2AB_CONFIGS:-c5} >> $J 2>> $OUT/${TAG}_ab.err
done
python - <<'P' $J
import json, sys
name = None
for line in open(sys.argv[1]):
    if line.startswith("=="):
        name = line.strip()
    elif line.startswith("{"):
        d = json.loads(line)
        print(f"{name:70s} {d['config'][:30]:30s} {d['ms_per_assembly']:.4f} ms  {d['plan']}")
P
