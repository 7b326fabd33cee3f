#!/bin/bash
mkdir -p gpurun_out
OG_LIB=openglue_b200/libopenglue_b200_trace.so timeout 120 python scripts/trace_sink.py 16 2048 2048 100 2>&1 | tee gpurun_out/o_trace_sink_C3.txt
OG_SINK_OCC=1 OG_LIB=openglue_b200/libopenglue_b200_trace.so timeout 120 python scripts/trace_sink.py 16 2048 2048 100 2>&1 | tee gpurun_out/o_trace_sink_C3_occ1.txt
