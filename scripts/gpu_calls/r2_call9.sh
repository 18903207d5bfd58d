#!/usr/bin/env bash
# 4-GPU: the driver's N=4 point of the scaling run (driver protocol), plus bin/halo.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c9; mkdir -p $OUT
N=4
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 20 --warmup 5 2> $OUT/bench_n4.err | grep '^{' > $OUT/bench_n4.json; tail -c 300 $OUT/bench_n4.err; cut -c1-500 $OUT/bench_n4.json
timeout 90 bin/halo -n $N --json $OUT/halo_cli_n4.jsonl 2>&1 | tail -1 | cut -c1-250
echo "== r2 call9 done"
