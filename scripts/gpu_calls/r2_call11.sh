#!/usr/bin/env bash
# Round 2, call 11 (2 GPUs): ncu captures of the kernels that spin on a peer (fused ring, pull ring, two-shot, NVLS).
# A replayed pass cannot wait for a peer, so rank 0 repeats its LAST launch with the same epochs (every word already
# satisfied, peers idle) between cudaProfilerStart/Stop: bin/allreduce --profile-relaunch.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c11; mkdir -p $OUT
timeout 60 bin/allreduce -n 2 -p 25 --iters 2 --profile-relaunch 2>&1 | tail -3
cap() {  # name, allreduce args...
  local name=$1; shift
  timeout 170 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $OUT/prof_$name \
    bin/allreduce -n 2 -p 25 --iters 2 --profile-relaunch "$@" > $OUT/ncu_$name.log 2>&1
  echo "[$name] rc=$? $(grep -c '==PROF==' $OUT/ncu_$name.log) prof lines"; tail -4 $OUT/ncu_$name.log | cut -c1-200
}
cap ring
cap twoshot -a --coll twoshot
cap nvls -a --coll nvls
cap ring_pull --pull
cap ring_slots2 --slots 2
ls -la $OUT
echo "== r2 call11 done"
