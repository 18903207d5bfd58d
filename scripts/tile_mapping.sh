#!/usr/bin/env bash
# Per-rank launcher: pick this rank's GPU by policy, export it, exec the program.
#   tile_mapping.sh <compact|spread|compact_plan> <CVD|SET> cmd [args...]
# Role of p2p/tile_mapping.sh in the reference (ZE_AFFINITY_MASK / ONEAPI_DEVICE_SELECTOR);
# here CVD = CUDA_VISIBLE_DEVICES, SET = HPCP_DEVICE (program calls cudaSetDevice).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"

rank="${LOCAL_RANK:-${OMPI_COMM_WORLD_LOCAL_RANK:-${PALS_LOCAL_RANKID:-${SLURM_LOCALID:-}}}}"
if [[ -z "$rank" ]]; then echo "tile_mapping.sh: no local rank in the environment" >&2; exit 2; fi
num_gpu="${HPCP_NUM_DEVICES:-$(nvidia-smi -L 2>/dev/null | grep -c '^GPU ' || true)}"
if [[ -z "$num_gpu" || "$num_gpu" -le 0 ]]; then echo "tile_mapping.sh: no GPU found" >&2; exit 2; fi
num_domain=2   # the two halves of an HGX baseboard (PCIe switch / NUMA domains)

policy="${1:-}"; shift || true
case "$policy" in
  compact)      gpu=$(( rank % num_gpu )) ;;
  spread)       # round-robin over the domains (contiguous blocks, the first n % d one GPU larger), skipping exhausted
                # ones — the rule of parallel/tile_mapping.py::_spread and topology_core.cpp::device_for_rank
                (( num_domain > num_gpu )) && num_domain=$num_gpu
                base=$(( num_gpu / num_domain )); extra=$(( num_gpu % num_domain ))
                r=$(( rank % num_gpu )); seen=0; level=0; gpu=-1
                while (( gpu < 0 )); do
                  for (( k = 0; k < num_domain; k++ )); do
                    size=$(( base + (k < extra ? 1 : 0) ))
                    if (( level < size )); then
                      if (( seen == r )); then gpu=$(( k * base + (k < extra ? k : extra) + level )); break; fi
                      seen=$(( seen + 1 ))
                    fi
                  done
                  level=$(( level + 1 ))
                done ;;
  compact_plan) gpu="$("$here/bin/topology" $(( rank % num_gpu )))" ;;
  *) echo "tile_mapping.sh: unknown policy '$policy' (compact|spread|compact_plan)" >&2; exit 2 ;;
esac

mechanism="${1:-}"; shift || true
export CUDA_DEVICE_ORDER=PCI_BUS_ID
case "$mechanism" in
  CVD|ZAM) export CUDA_VISIBLE_DEVICES="$gpu"; export HPCP_DEVICE=0 ;;   # ZAM/ODS: the reference's spellings
  SET|ODS) export HPCP_DEVICE="$gpu" ;;
  *) echo "WRONG AFFINITY MECHANISM EITHER CVD OR SET" >&2; exit 2 ;;
esac
exec "$@"
