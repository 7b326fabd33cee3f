#!/bin/bash
set +e
mkdir -p gpurun_out; rm -f gpurun_out/sink_exp.log
for cfg in "0 16" "8 18" "6 24" "5 29" "4 37" "3 49" "2 74"; do
  set -- $cfg
  echo "== OG_SINK_PAIRS=$1 OG_SINK_SP=$2" >> gpurun_out/sink_exp.log
  OG_SINK_PAIRS=$1 OG_SINK_SP=$2 timeout 300 python - >> gpurun_out/sink_exp.log 2>&1 <<'PY'
import ctypes as C, sys, torch
sys.path.insert(0, '.')
from openglue_b200 import _cabi
lib = _cabi.lib(); dev='cuda:0'
B,n=16,2048
S=torch.randn(B,n,n,device=dev)*4; scores=torch.empty(B,n+1,n+1,device=dev); dust=torch.ones(1,device=dev)
wsb=lib.og_sinkhorn_workspace_bytes(B,n,n); ws=torch.empty(wsb,dtype=torch.uint8,device=dev)
st=C.c_void_p(torch.cuda.current_stream().cuda_stream); p=lambda t:C.c_void_p(t.data_ptr())
def run(): _cabi.check(lib.og_sinkhorn_fwd(p(S),n,n*n,p(dust),B,n,n,100,1.0,p(scores),p(ws),wsb,st),'sink')
run(); run(); torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
print('ms per launch-set', e0.elapsed_time(e1)/5, 'checksum', float(scores.double().sum()))
PY
done
cat gpurun_out/sink_exp.log
