#!/bin/bash
# Round-2 visit B: first run of the fp16 hi/lo kernels.  Every group runs in its own process (a trap poisons the context).
cd "$GRAFT_REPO_ROOT" || exit 1
set +e
mkdir -p gpurun_out
run() { name=$1; shift; timeout 420 python -m pytest "$@" -q -s --timeout 300 2>&1 | grep -v "^$" | tail -40 > gpurun_out/b_$name.log; echo "== $name"; grep -E "^\[|passed|failed|error" gpurun_out/b_$name.log | tail -25; }
run split tests/test_gpu_f16.py -k "weight_split"
run probe tests/test_gpu_f16.py -k "layout_probe"
run gemm_y tests/test_gpu_f16.py -k "fp32_output"
run gemm_split tests/test_gpu_f16.py -k "split_outputs"
run attn tests/test_gpu_f16.py -k "attention_f16"
run path_small tests/test_gpu_parity.py -k "matches_oracle and fp16x3"
run path_big tests/test_gpu_parity.py -k "reference_big and fp16x3"
timeout 600 python bench.py --precision fp16x3 --no-cpu-baseline > gpurun_out/b_bench_f16.json 2> gpurun_out/b_bench_f16.err
tail -c 1500 gpurun_out/b_bench_f16.json; tail -5 gpurun_out/b_bench_f16.err
