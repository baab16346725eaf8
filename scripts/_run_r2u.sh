#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest.txt 2>&1; tail -15 gpurun_out/r2u_pytest.txt
python - <<'PY' > gpurun_out/r2u_decided.txt 2>&1
import sys
sys.path.insert(0,'.')
from tests import common
from tests.test_gpu_parity import _first_bounce_rays
import numpy as np
for name in ["disk3D","trench","holes","sphere3D","plane"]:
    c=common.case(name); orc=common.make_oracle(c); ctx,src,st=common.make_gpu(c)
    rays=orc.source_rays(common.oracle_particle(c), orc.config(10**7,3),0,300000)
    for leg in range(3):
        r=ctx.debug_nb_shortcut(rays)
        print(name,"leg",leg,r,"decided %.3f"%(r["decided"]/max(r["examined"],1)))
        go,po_,to,_,_=ctx.debug_intersect(rays,nb_cap=1); keep=go==1
        rays=_first_bounce_rays(orc,rays[keep],po_[keep],to[keep],c,seed=leg)
    ctx.close()
PY
cat gpurun_out/r2u_decided.txt
for e in "" "VR_NB_SHORTCUT_OFF=1"; do
  echo "C4 both 256e6 [$e]: $(env $e python scripts/profile_step.py 256e6 both 2>&1 | tail -1 | cut -d' ' -f6-)"
  echo "C4 both 256e6 lanes1 [$e]: $(env $e VR_LANES=1 VR_TIME_KERNELS=1 python scripts/profile_step.py 256e6 both 2>&1 | grep phases | tail -1)"
  echo "C5 4e8 [$e]: $(env $e python scripts/profile_c5.py 4e8 2>&1 | grep 'rep 1')"
done > gpurun_out/r2u_timing.txt 2>&1
cat gpurun_out/r2u_timing.txt
