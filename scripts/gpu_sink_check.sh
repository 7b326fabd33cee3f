#!/bin/bash
set +e
for cfg in "0 16" "0 40" "0 74"; do
  set -- $cfg
  echo "== OG_SINK_PAIRS=$1 OG_SINK_SP=$2"
  OG_SINK_PAIRS=$1 OG_SINK_SP=$2 timeout 300 python - <<'PY'
import ctypes as C, sys, torch
sys.path.insert(0, '.')
from openglue_b200 import _cabi
from oracle import superglue_oracle as O
lib = _cabi.lib(); dev='cuda:0'
B,n,m,T=2,2048,2048,100
g=torch.Generator().manual_seed(5)
S=4*torch.randn(B,n,m,generator=g)
dust=torch.tensor(1.3)
ref=O.matching_log_probs(S.double(),dust.double(),T,1.0)
dS=S.to(dev).contiguous(); scores=torch.empty(B,n+1,m+1,device=dev)
wsb=lib.og_sinkhorn_workspace_bytes(B,n,m); ws=torch.empty(wsb,dtype=torch.uint8,device=dev)
st=C.c_void_p(torch.cuda.current_stream().cuda_stream); p=lambda t:C.c_void_p(t.data_ptr())
_cabi.check(lib.og_sinkhorn_fwd(p(dS),m,n*m,p(dust.to(dev)),B,n,m,T,1.0,p(scores),p(ws),wsb,st),'sink')
torch.cuda.synchronize()
print('max err vs fp64 oracle', float((scores.cpu().double()-ref).abs().max()), 'checksum', float(scores.double().sum()), 'ref', float(ref.sum()))
PY
done
