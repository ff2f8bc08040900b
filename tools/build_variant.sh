#!/bin/bash
# A/B build: recompiles the named source files with extra flags (in parallel) and links them with the default objects
# into dune-gdt_b200/lib/ab/libgdtb_<name>.so   (select at run time with GDTB_LIB=<path>)
# usage: tools/build_variant.sh <name> <file.cu> "<flags>" [<file2.cu> "<flags2>" ...]
set -e
cd "$(dirname "$0")/../dune-gdt_b200/csrc"
NAME=$1; shift
mkdir -p ../build/ab ../lib/ab
declare -A REPL
PIDS=""
while [ $# -gt 0 ]; do
  F=$1; FL=$2; shift 2
  O=../build/ab/${NAME}_${F%.cu}.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
    -Xcompiler -fPIC,-Wall,-Wno-unknown-pragmas --expt-relaxed-constexpr $FL -c $F -o $O &
  PIDS="$PIDS $!"
  REPL[${F%.cu}]=$O
done
make -s -j4 &
PIDS="$PIDS $!"
for p in $PIDS; do wait $p; done
OBJS=""
for f in capi solve assemble_generic assemble_q1_gather assemble_q2_gather assemble_q2_qp assemble_dg_gather assemble_dg_fast pattern fv fv_system fv_tma; do
  if [ -n "${REPL[$f]}" ]; then OBJS="$OBJS ${REPL[$f]}"; else OBJS="$OBJS ../build/$f.o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/ab/libgdtb_${NAME}.so $OBJS
echo built dune-gdt_b200/lib/ab/libgdtb_${NAME}.so
