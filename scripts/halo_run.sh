#!/usr/bin/env bash
# Sweep the fused halo exchange (bin/halo) over mode x rows x ranks next to its stock arm, one JSON row per point.
# Same role for `halo` as p2p/run.sh has for peer2pear in the reference (build, sweep the variants, one line each).
#   HPCP_HALO_JSON=out.jsonl scripts/halo_run.sh            # all GPUs of the node
#   HPCP_HALO_BYTES=8388608 HPCP_HALO_RANKS="2 4" scripts/halo_run.sh
set -x
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
make -C "$here" -j bin/halo >/dev/null || exit 1
cd "$here" || exit 1
ngpu=$(nvidia-smi -L 2>/dev/null | grep -c '^GPU ')
ranks=${HPCP_HALO_RANKS:-"1 2 $ngpu"}
bytes=${HPCP_HALO_BYTES:-188743680}
for n in $ranks; do
  [ "$n" -ge 1 ] || continue
  for rows in 1 8; do
    for variant in "--mode pull" "--mode push" "--mode pull --per-step" "--stock memcpy"; do
      # shellcheck disable=SC2086
      "$here/bin/halo" -n "$n" --rows "$rows" --bytes "$bytes" $variant ${HPCP_HALO_JSON:+--json "$HPCP_HALO_JSON"} | tail -1
    done
  done
done
