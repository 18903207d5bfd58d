#!/usr/bin/env bash
# Round 2, last 2-GPU call: the tree as committed — driver-protocol bench at N=2 and the multi-GPU tests a 1-GPU box skips.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c13; mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29677 \
  bench.py --gpus 2 --steps 20 --warmup 5 2> $OUT/bench_n2.err | grep '^{' > $OUT/bench_n2.json; cut -c1-420 $OUT/bench_n2.json
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_gemm_collective.py -m gpu -q --timeout 200 -k "not virtual_ranks" \
  > $OUT/pytest_multi.txt 2>&1; tail -5 $OUT/pytest_multi.txt
echo "== r2 call13 done"
