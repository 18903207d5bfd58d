#!/usr/bin/env bash
# Sweep the concurrency benchmark on the host (CPU / OpenMP backend): the no-GPU plumbing
# configuration.  Role of concurency/run_omp.sh in the reference; the two compile-time
# builds (-DHOST_THREADS, -DNOWAIT) are run-time modes of one binary here.
set -o xtrace
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
make -C "$here" bin/omp_con >/dev/null || exit 1
work=$(mktemp -d tmp-omp-XXXX); cd "$work" || exit 1
rm -f omp.log

LCOMMANDS=("C C" "C M2D" "C D2M" "M2D D2M" "H2D D2H")
SIZE=${HPCP_OMP_ELEMS:-20000000}

for envs in "OMP_PROC_BIND=false" \
            "OMP_PROC_BIND=spread OMP_PLACES=cores" \
            "OMP_WAIT_POLICY=active"
do
    (
    export $envs
    for mode in "host_threads" "nowait"; do
        # shellcheck disable=SC2068
        "$here/bin/omp_con" "$mode" --globalsize_default_memory "$SIZE" ${LCOMMANDS[@]/#/--commands }
    done
    ) |& tee -a omp.log
done
PYTHONPATH="$here" python -m hpc_patterns_b200.utils.parse omp.log
