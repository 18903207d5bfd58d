#!/usr/bin/env bash
# N=1 (loop-back, HBM-bound) sweep of the flagship kernel's tiling.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
mkdir -p gpurun_out
run() { local tag=$1; shift; timeout 100 python bench.py --gpus 1 --no-extras --e2e-steps 1 "$@" 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$tag', d['ms_per_step'], d['value'])"; }
run "tma default(16k x6)" --engine tma
run "tma 16k x3 (2/SM)" --engine tma --stages 3
run "tma 8k x6 (2/SM)" --engine tma --stage-kb 8 --stages 6
run "tma 8k x4 (3/SM)" --engine tma --stage-kb 8 --stages 4
run "tma 32k x3" --engine tma --stage-kb 32 --stages 3
run "tma 4k x8 (3/SM)" --engine tma --stage-kb 4 --stages 8
run "ldst default" --engine ldst
run "ldst c592 u2" --engine ldst --ctas 592
run "ldst c592 u4" --engine ldst --ctas 592 --unroll 4
run "ldst v32 c592" --engine ldst --ctas 592 --vec 32
run "ldst v32 c296" --engine ldst --vec 32
