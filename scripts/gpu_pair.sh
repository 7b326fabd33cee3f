#!/bin/bash
# cta_group::2 forms of the GEMM / attention kernels: parity + whole-step timing + GEMM event trace
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pair_tests.log 2>&1; echo "rc=$?" >> gpurun_out/pair_tests.log
tail -3 gpurun_out/pair_tests.log
for cfg in 0,0 1,0 0,1 1,1; do
  g=${cfg%,*}; a=${cfg#*,}
  OG_GEMM_PAIR=$g OG_ATTN_PAIR=$a timeout 300 python bench.py --steps 4 --no-cpu-baseline > gpurun_out/pair_bench_g${g}_a${a}.json 2> gpurun_out/pair_bench_g${g}_a${a}.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/pair_bench_g${g}_a${a}.json'))
    print('gemm_pair=$g attn_pair=$a', round(d['value'], 1), 'pairs/s', round(d['ms_per_step'], 3), 'ms  attn', round(d['roofline']['ms_per_launch'], 4), 'ms')
except Exception as e:
    print('gemm_pair=$g attn_pair=$a failed', e)
PY
done
if [ -f openglue_b200/libopenglue_b200_trace.so ]; then
  OG_GEMM_PAIR=0 timeout 120 python scripts/trace_gemm.py > gpurun_out/trace_gemm_single.log 2>&1
  OG_GEMM_PAIR=1 timeout 120 python scripts/trace_gemm.py > gpurun_out/trace_gemm_pair5.log 2>&1
fi
