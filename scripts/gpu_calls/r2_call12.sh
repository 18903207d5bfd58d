#!/usr/bin/env bash
# Round 2, final 1-GPU call: the tree as committed — whole GPU suite, smoke(), driver-protocol bench, reference-arm line.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c12; mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu_full.txt 2>&1; tail -6 $OUT/pytest_gpu_full.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 2> $OUT/bench_n1.err | grep '^{' > $OUT/bench_n1.json; cut -c1-400 $OUT/bench_n1.json
timeout 100 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference_arm.json 2>&1; cut -c1-300 $OUT/bench_reference_arm.json
echo "== r2 call12 done"
