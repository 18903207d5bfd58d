#!/usr/bin/env bash
# Round 2, first GPU call (1 GPU): the new flagship (K-halo) — exactness, driver-protocol bench, geometry sweep, ncu.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c1; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_halo.py -x -q --timeout 120 2>&1 | tail -15 | tee $OUT/pytest_halo.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 600 $OUT/bench_n1.err
cut -c1-1500 $OUT/bench_n1.json
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 5 --no-extras > $OUT/bench_n1_200.json 2>> $OUT/bench_n1.err
cut -c1-400 $OUT/bench_n1_200.json
timeout 600 python scripts/halo_tune.py --out $OUT/halo_tune_n1.jsonl 2>&1 | tail -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:halo_stencil -s 1 -c 1 -f -o $OUT/prof_halo_pull_n1 \
  python scripts/ncu_halo.py --mode pull > $OUT/ncu_halo.log 2>&1; tail -3 $OUT/ncu_halo.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== r2 call1 done"
