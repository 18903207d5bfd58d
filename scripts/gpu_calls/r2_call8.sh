#!/usr/bin/env bash
# 1-GPU: ncu of the flagship kernel at the benchmark's real shape and final geometry (rows 8 x 188 743 680 B, 2 CTAs/SM),
# pull and push, plus the launch list of a short bench run (device time per launch, shares only).
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out/r2c8; mkdir -p $OUT
export PYTHONPATH=$PWD
for mode in pull push; do
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:halo_stencil -s 2 -c 1 -f -o $OUT/prof_halo_${mode}_n1_full \
    python scripts/ncu_halo.py --mode $mode --rows 8 --bytes 188743680 --steps 4 > $OUT/ncu_$mode.log 2>&1; tail -1 $OUT/ncu_$mode.log
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 200 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 5 --warmup 3 --no-extras --blocks 2 --preheat-ms 20 --e2e-steps 1 > $OUT/bench_under_ncu.log 2>&1
tail -2 $OUT/bench_under_ncu.log | cut -c1-200
grep -c halo_stencil $OUT/launches.csv
echo "== r2 call8 done"
