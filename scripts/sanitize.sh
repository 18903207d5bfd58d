#!/usr/bin/env bash
# Race / memory checking of the single-GPU kernels with compute-sanitizer (run on a GPU box).
# The reference has no sanitizer story at all (only -Werror); its known latent race is the shared
# send/receive buffer of peer2pear's bidirectional phase, which this suite removed by design.
#   scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck]...
set -u
cd "$(dirname "$0")/.."
OUT=${OUT:-gpurun_out}; mkdir -p "$OUT"
tools=("$@"); [ ${#tools[@]} -eq 0 ] && tools=(memcheck racecheck synccheck)
rc=0
for tool in "${tools[@]}"; do
  echo "== compute-sanitizer --tool $tool"
  # small problem sizes: the sanitizer slows kernels 10-100x
  timeout 900 compute-sanitizer --tool "$tool" --error-exitcode 9 --log-file "$OUT/sanitize_$tool.log" \
    python -m pytest tests/test_gpu_kernels.py -q -x --timeout 600 \
      -k "copy_matches_torch and 1024 or copy_signal or fill_and_verify or triad_put_matches and 4096 or fused_bench_commands or allreduce_building_blocks or tcgen05_tile_loop and 3-5" \
      > "$OUT/sanitize_$tool.pytest.log" 2>&1 || rc=1
  tail -3 "$OUT/sanitize_$tool.pytest.log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$OUT/sanitize_$tool.log" | tail -2
done
exit $rc
