#!/bin/bash
# tools/mprun.sh N prog [args...]: starts N processes of a C++ facade driver, one per GPU (RANK / LOCAL_RANK / WORLD_SIZE
# as torchrun sets them, plus a fresh rendezvous directory for Parallel::FileRendezvous); exit code 0 iff all ranks succeed.
N=$1; shift
DIR=$(mktemp -d /dev/shm/gdtb_rv.XXXXXX 2>/dev/null || mktemp -d)
pids=()
for r in $(seq 0 $((N - 1))); do
  RANK=$r LOCAL_RANK=$r WORLD_SIZE=$N GDTB_RENDEZVOUS_DIR=$DIR "$@" &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
rm -rf "$DIR"
exit $rc
