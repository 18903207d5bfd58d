#!/usr/bin/env bash
# Round 2, call 10 (2 GPUs): L2 evict_first hint on the flagship's streaming traffic, N=1 and N=2, plus the halo tests.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c10; mkdir -p $OUT
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671 "$@"; }
N=2 run scripts/halo_tune.py --out $OUT/halo_l2hint_n2.jsonl --modes pull push --geometry 16x6 --l2-hint 0 1 0 1 --steps 50 2>&1 | grep '^{' | cut -c1-330
N=1 CUDA_VISIBLE_DEVICES=0 run scripts/halo_tune.py --out $OUT/halo_l2hint_n1.jsonl --modes pull push --geometry 16x6 --l2-hint 0 1 0 1 --steps 50 2>&1 | grep '^{' | cut -c1-330
N=2 run scripts/halo_tune.py --out $OUT/halo_l2hint_n2_rows1.jsonl --modes pull push --geometry 16x6 --rows 1 --l2-hint 0 1 --steps 50 2>&1 | grep '^{' | cut -c1-330
timeout 300 python -m pytest tests/test_gpu_halo.py -q --timeout 120 2>&1 | tail -4 | tee $OUT/pytest_halo.txt
for v in "" "--l2-hint" "--mode push" "--mode push --l2-hint"; do
  timeout 120 bin/halo -n 2 $v --json $OUT/halo_cli_n2.jsonl 2>&1 | tail -1 | cut -c1-250
done
echo "== r2 call10 done"
