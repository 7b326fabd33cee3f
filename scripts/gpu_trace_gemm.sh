#!/bin/bash
# event traces of the persistent GEMM, single-CTA and cta_group::2 forms (debug build with -DOG_TRACE)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
# libopenglue_b200_trace.so is built in the container (nvcc ... -DOG_TRACE -o openglue_b200/libopenglue_b200_trace.so) and travels with the snapshot
OG_GEMM_PAIR=0 timeout 120 python scripts/trace_gemm.py > gpurun_out/trace_gemm_single.log 2>&1
OG_GEMM_PAIR=1 timeout 120 python scripts/trace_gemm.py > gpurun_out/trace_gemm_pair4.log 2>&1
tail -n 12 gpurun_out/trace_gemm_single.log gpurun_out/trace_gemm_pair4.log
