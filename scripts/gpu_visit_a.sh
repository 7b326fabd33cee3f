#!/bin/bash
# new tests (pipelined submit, configs[4]-shaped parity) + the headline bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "submit or forward_matches_oracle or host_buffers or cuda_graph" > gpurun_out/visit_a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/visit_a_tests.log
tail -5 gpurun_out/visit_a_tests.log
timeout 600 python bench.py > gpurun_out/visit_a_bench.json 2> gpurun_out/visit_a_bench.err
tail -c 3000 gpurun_out/visit_a_bench.json
tail -5 gpurun_out/visit_a_bench.err
