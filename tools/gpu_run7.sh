#!/bin/bash
O=$GRAFT_REPO_ROOT/gpurun_out; mkdir -p $O; cd $GRAFT_REPO_ROOT
export DCS_DEBUG_TMA_PERSIST=0
for p in 0 1 2 4 8 16 17 3 12 15; do
  DCS_DEBUG_TMA_PROBE=$p timeout 300 python - > $O/probe_$p.txt 2>&1 <<PY
import sys, torch, json
sys.path.insert(0, '.')
import bench
from deepconvsep_b200.engine import Separator
cfg = bench.CONFIGS['dsd2048']
params = bench.synthetic_params('dsd', 1025, 0)
sep = bench.make_separator(cfg, params, 0)
L = 180*44100
x = bench.synth_clip_device(L, 1000, torch.device('cuda',0))
out = torch.empty((4, L), device='cuda')
for _ in range(3): sep.separate_device(x, out)
torch.cuda.synchronize()
sep.ctx.profile(True)
for _ in range(8): sep.separate_device(x, out)
torch.cuda.synchronize()
recs = sep.ctx.profile_read()
st = {}
for n,t in recs: st.setdefault(n, []).append(t)
print(json.dumps({k: round(sum(v)/len(v),4) for k,v in st.items()}))
PY
done
echo run7 done
