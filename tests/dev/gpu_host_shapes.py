import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from pyloo_b200 import engine
from oracle import psis_oracle as orc
rng=np.random.default_rng(5)
for S,N,reff in ((16000,3001,1.0),(8000,2500,0.5),(1000,9001,0.7)):
    ll=-1.4+rng.normal(size=(S,N))
    engine.profile(True)
    r=engine.loo_host(ll,reff,device=0)
    prof=engine.profile_read(); engine.profile(False)
    idx=np.arange(0,N,N//40)[:40]
    pw=orc.loo_pointwise(ll[:,idx],reff)
    print(S,N,reff,'max rel err',float(np.max(np.abs(r['elpd_i'][idx]-pw['elpd_i'])/np.abs(pw['elpd_i']))), float(np.max(np.abs(r['pareto_k'][idx]-pw['pareto_k']))), {k:v[1] for k,v in prof.items() if v[1]}, r['stats'].n)
