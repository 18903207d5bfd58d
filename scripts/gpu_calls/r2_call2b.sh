#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c2b; mkdir -p $OUT
export CUDA_VISIBLE_DEVICES=0
for t in get sendrecv memcpy; do
  echo "=== peer2pear $t" | tee -a $OUT/virt.txt
  timeout 60 bin/peer2pear v -n 2 --transport $t --bytes 4194304 --iters 2 2>&1 | tail -6 | tee -a $OUT/virt.txt
done
echo "=== allreduce -n 4" | tee -a $OUT/virt.txt
timeout 100 bin/allreduce -n 4 -p 18 --iters 2 2>&1 | tail -8 | tee -a $OUT/virt.txt
echo "=== allreduce -n 4 --ctas 16" | tee -a $OUT/virt.txt
timeout 100 bin/allreduce -n 4 -p 18 --iters 2 --ctas 16 2>&1 | tail -8 | tee -a $OUT/virt.txt
echo "=== allreduce -n 3" | tee -a $OUT/virt.txt
timeout 100 bin/allreduce -n 3 -p 18 --iters 2 2>&1 | tail -8 | tee -a $OUT/virt.txt
echo "=== allreduce -n 4 -a twoshot" | tee -a $OUT/virt.txt
timeout 100 bin/allreduce -n 4 -p 18 --iters 2 -a --coll twoshot 2>&1 | tail -8 | tee -a $OUT/virt.txt
echo "=== halo -n 4 stock" | tee -a $OUT/virt.txt
timeout 100 bin/halo -n 4 --bytes 4198400 --rows 3 --steps 5 --iters 2 --stock memcpy 2>&1 | tail -8 | tee -a $OUT/virt.txt
echo "=== halo -n 3 stock" | tee -a $OUT/virt.txt
timeout 100 bin/halo -n 3 --bytes 4198400 --rows 3 --steps 5 --iters 2 --stock memcpy 2>&1 | tail -8 | tee -a $OUT/virt.txt
