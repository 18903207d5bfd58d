#!/usr/bin/env bash
# Final validation on 2 GPUs: whole GPU test-suite, bench at N=1/2, fused GEMM->put over NVLink, peer-letter groups.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -12 | tee $OUT/final_pytest.txt
timeout 200 python bench.py --gpus 1 | tee $OUT/final_bench_n1.json | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 2>/dev/null | grep '^{' | tee $OUT/final_bench_n2.json | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29592 scripts/gemm_put_bench.py 2>/dev/null | grep '^{' | tee $OUT/final_gemm_put_n2.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items()})"
timeout 200 ./bin/concurency fused --repetitions 5 --commands C D2P --commands D2P P2D --commands T D2P --commands A D2P --json $OUT/final_conc.jsonl 2>&1 | grep -E "^##|Speedup Rel|Total Time //" | tee $OUT/final_concurency.txt
HPCP_T_OUT=P HPCP_T_CLUSTER=2 timeout 100 ./bin/concurency in_order --repetitions 5 --commands T C 2>&1 | grep -E "^##|Speedup Rel" | tee -a $OUT/final_concurency.txt
echo "== final done"
