#!/usr/bin/env bash
# Round 2, call 14 (2 GPUs): first runs of the instantiations still listed as UNVALIDATED — typed pull / two-slot rings,
# two-shot for > 8 ranks and the remaining type classes (thread-ranks), NVLS for 32-bit integers (needs 2 real GPUs).
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c14; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 120 \
  -k "ring_variants_every_type_class or two_shot_large_worlds or nvls_integer or profile_relaunch" > $OUT/pytest_variants.txt 2>&1
tail -6 $OUT/pytest_variants.txt
echo "== r2 call14 done"
