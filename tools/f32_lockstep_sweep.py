"""fp32-storage flavour of the QP kernels (host build of qp_split.cuh, tests/emu) against the fp64 oracle in lock step: every solve of 6 closed-loop
steps starts from the oracle's guess; counts status mismatches and accepted solves whose first control differs by more than 1e-3 relative.
Development tool (CPU only), cited in DESIGN.md section 0 row N1."""
import os
import sys

import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from tests.common import make_problem, start_states, rollout_guess
from tests.emu import kernel_source_oracle
from oracle.oracle import Oracle
from safe_mpc_b200 import abi
for vel,spread,scale,tag in ((0.5,0.15,2.0,'test-like'),(0.6,0.3,3.0,'aggressive')):
  tot=over=stdiff=stall=0; worst=0.0
  for ctrl,cost in (('naive','ext'),('st','ext'),('htwa','ext'),('receding','ext'),('constraint_everywhere','ext'),('zerovel','nls'),('stwa','ext')):
    for N in (10, 23, 45):
        for seed in (1,2,3):
            B=16
            prob, params, md = make_problem(ctrl, cost=cost, N=N); prob32,_,_ = make_problem(ctrl, cost=cost, N=N, precision='f32')
            a=Oracle(prob,B,0); b=kernel_source_oracle(prob32,B,0,f32=True)
            x0=start_states(B,seed=seed*7,vel=vel,spread=spread)
            xg,ug=rollout_guess(x0,N,params.dt,seed=seed*11,scale=scale)
            for e in (a,b):
                e.set_guess(xg,ug); e.reset_controller()
            x=x0.copy()
            for step in range(6):
                ua,aa=a.controller_step(x); ub,ab=b.controller_step(x)
                sa=a.get_state(abi.STATE_STATUS); sb=b.get_state(abi.STATE_STATUS)
                qb=b.get_state(abi.STATE_QP_STATUS)
                rel=np.abs(ua-ub).max(axis=1)/np.maximum(1.0,np.abs(ua).max(axis=1))
                ok=(sa==0)&(sb==0)
                tot+=B; stdiff+=int((sa!=sb).sum()); over+=int((rel[ok]>1e-3).sum()); stall+=int(((rel>1e-3)&ok&(qb==1)).sum()); worst=max(worst,float(rel[ok].max()) if ok.any() else 0.0)
                # keep the two in lock step: the fp32 handle continues from the oracle's guess
                b.set_guess(*a.get_guess())
                x,_=a.plant_step(x,ua)
  print(tag,'solves',tot,'status differs',stdiff,'accepted solves with rel err > 1e-3:',over,'of which stall exits',stall,'worst',worst)
