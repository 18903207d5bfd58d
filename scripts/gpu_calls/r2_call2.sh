#!/usr/bin/env bash
# Round 2, second GPU call (1 GPU): whole GPU suite, the never-run kernels (2-SM UMMA GEMM, GEMM+collective fusions on
# virtual ranks), driver-protocol bench with all extras, a finer K-halo geometry sweep, fused-mode concurrency verdicts.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c2; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
export HPCP_EXPERIMENTAL=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -k "2sm" -q --timeout 120 2>&1 | tail -8 | tee $OUT/pytest_2sm.txt
timeout 400 python -m pytest tests/test_gpu_kernels.py -k "l2_hint or halo_ctas or template" -q --timeout 120 2>&1 | tail -8 | tee $OUT/pytest_misc_experimental.txt
timeout 900 python -m pytest tests/test_gpu_gemm_collective.py -q --timeout 120 2>&1 | tail -25 | tee $OUT/pytest_gemm_collective.txt
timeout 300 python scripts/gemm_put_bench.py 2>/dev/null | grep '^{' | tee $OUT/gemm_put_bench.jsonl | cut -c1-400
timeout 300 python scripts/tp_bench.py --check --mlp --tokens 8192 --out-features 8192 --in-features 4096 2>&1 | tail -3 | tee $OUT/tp_bench_n1.txt | cut -c1-600
unset HPCP_EXPERIMENTAL
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 400 $OUT/bench_n1.err; cut -c1-2500 $OUT/bench_n1.json
timeout 600 python scripts/halo_tune.py --out $OUT/halo_tune_n1.jsonl --modes pull --geometry 16x6 16x7 12x8 12x9 8x6 8x8 24x6 32x6 2>&1 | grep '^{' | cut -c1-260
timeout 200 python scripts/halo_tune.py --out $OUT/halo_tune_n1.jsonl --modes push --geometry 16x6 16x7 8x6 2>&1 | grep '^{' | cut -c1-260
timeout 200 python scripts/halo_tune.py --out $OUT/halo_tune_n1_rows.jsonl --modes pull --geometry 16x6 --rows 1 3 6 2>&1 | grep '^{' | cut -c1-260
for mode in in_order fused; do
  for t in 1 0; do
    HPCP_FUSED_SIDE_THREADS=$t timeout 300 bin/concurency $mode --commands C C --commands C M2D --commands C D2M --commands M2D D2M --commands H2D D2H --json $OUT/concurency_${mode}_threads$t.jsonl 2>&1 | grep -E "^##|Speedup Rel" | tee -a $OUT/concurency_${mode}_threads$t.txt
    [ $mode = in_order ] && break
  done
done
echo "== r2 call2 done"
