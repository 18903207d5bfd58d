#!/usr/bin/env bash
# Build topology + peer2pear and sweep mapping policy x affinity mechanism x transport x ranks.
# Role of p2p/run.sh in the reference (compact|spread|compact_plan x ZAM|ODS x {Isend/Irecv, Put} x n).
# Single-process (thread-per-rank) binary: the policy is applied inside with --mapping; the
# process-per-GPU variant goes through tile_mapping.sh + torchrun.
set -x
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
make -C "$here" -j bin/topology bin/peer2pear >/dev/null || exit 1
cd "$here" || exit 1
# HPCP_P2P_CPU=1: host-only plumbing run of the same sweep (ranks are threads, the transport is memcpy).
if [ -n "${HPCP_P2P_CPU:-}" ]; then
  ngpu=${HPCP_NUM_DEVICES:-4}; extra=(--cpu --bytes "${HPCP_P2P_BYTES:-1048576}" --iters 3)
else
  ngpu=$(nvidia-smi -L | grep -c '^GPU '); extra=()
  "$here/bin/topology" --matrix
fi

for mode in compact spread compact_plan; do
  for transport in sendrecv put get memcpy; do
    for n in 2 "$ngpu"; do
      [ "$n" -ge 2 ] || continue
      "$here/bin/peer2pear" "peer2pear_$transport $n $mode" -n "$n" --mapping "$mode" --transport "$transport" \
          "${extra[@]}" ${HPCP_P2P_JSON:+--json "$HPCP_P2P_JSON"}
    done
  done
done

[ -n "${HPCP_P2P_CPU:-}" ] && exit 0
# process-per-GPU flavour: rank -> GPU chosen by the wrapper, both mechanisms
for mode in compact spread compact_plan; do
  for mech in CVD SET; do
    [ "$mech" = CVD ] && continue   # CUDA IPC peer mappings need every GPU visible in every process
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --no-python \
        "$here/scripts/tile_mapping.sh" "$mode" "$mech" python -m hpc_patterns_b200.models.peer2pear \
        "torchrun 2 $mode $mech" --transport put
  done
done
