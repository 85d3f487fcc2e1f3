import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, pyloo_b200 as pl
ll = np.random.default_rng(0).normal(-1.4, 1.0, size=(4, 1000, 250_000))
idata = pl.from_dict(log_likelihood={"y": ll})
res = pl.loo(idata, pointwise=True, reff=1.0)
t0 = time.perf_counter(); res = pl.loo(idata, pointwise=True, reff=1.0); print("seconds", time.perf_counter() - t0)
print(res)
