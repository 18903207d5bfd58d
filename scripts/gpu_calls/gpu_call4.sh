#!/usr/bin/env bash
# 1-GPU call: tcgen05 payload checks + ncu, fused-concurrency diagnosis, sanitizer pass.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out; mkdir -p $OUT
t() { local secs=$1; shift; timeout "$secs" "$@"; }
t 300 python -m pytest tests/test_gpu_kernels.py -q --timeout 200 -k "tcgen05 or tensor_command or smoke" -s 2>&1 | tail -15 | tee $OUT/call4_tests.txt
# fused-mode copy rates per direction (single-command groups: serial = copy engine, // = fused TMA kernel)
t 200 ./bin/concurency fused --repetitions 5 --commands H2D --commands D2H --commands D2D --commands H2D D2H --commands T H2D --commands T C 2>&1 | grep -E "^##|Minimum Time|Total Time|Speedup|Param|tripcount|globalsize" | tee $OUT/call4_concurency.txt
HPCP_FUSED_COPY_ENGINE=ldst t 200 ./bin/concurency fused --repetitions 5 --commands H2D --commands D2H --commands H2D D2H 2>&1 | grep -E "^##|Total Time|Speedup" | tee -a $OUT/call4_concurency.txt
t 200 ./bin/concurency in_order --repetitions 5 --commands T H2D --commands T C --commands T A 2>&1 | grep -E "^##|Minimum Time|Total Time|Speedup" | tee -a $OUT/call4_concurency.txt
# ncu: tensor pipe utilisation of the T kernel
cat > /tmp/tc_run.py <<'PY'
import torch, hpc_patterns_b200 as h
from hpc_patterns_b200.ops import payload
C = h.native()
ops = payload.tc_operands()
for _ in range(3):
    out = payload.tc_busy(ops, 148, 20000)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = payload.tc_busy(ops, 148, 40000); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("tc_busy 148 CTAs x 40000 passes: %.3f ms -> %.0f TFLOP/s bf16" % (ms, 148*40000*2.0*128*256*64/ms/1e9))
print("exact:", bool(torch.equal(out[0], payload.tc_busy_reference(ops, 40000))))
PY
PYTHONPATH=. t 120 python /tmp/tc_run.py 2>&1 | tee $OUT/call4_tc.txt
PYTHONPATH=. t 300 ncu --set full --clock-control none --import-source on -k regex:tc_busy -s 2 -c 1 -f -o $OUT/prof_tc_busy python /tmp/tc_run.py > $OUT/ncu_tc.log 2>&1
# sanitizer (small sizes)
t 600 bash scripts/sanitize.sh memcheck racecheck 2>&1 | tail -12 | tee $OUT/call4_sanitize.txt
echo "== call4 done"
