#!/bin/bash
# visit L: does the evict_last subset reach L2?  dram bytes / L2 hit rate of the Sinkhorn launch with and without it
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_read_lookup_hit.sum,lts__t_sectors_op_read_lookup_miss.sum
for cfg in "0 0" "200 64" "0 64" "200 40"; do
  set -- $cfg
  echo "== PERSIST_MB=$1 L2_MB=$2"
  OG_PERSIST_MB=$1 OG_SINK_L2_MB=$2 timeout 300 ncu --metrics $M --clock-control none -k regex:sinkhorn_kernel -s 3 -c 1 python scripts/sink_l2_exp.py 16 2048 2048 100 2>&1 | grep -E "dram__|lts__|gpu__time|persisting"
done | tee gpurun_out/l_sink_l2_ncu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "oracle" 2>&1 | tail -3
