"""Numerical prototype (NumPy) of the algebraic shortcuts the CUDA kernel uses, checked against
the oracle: product-form GPD profile, shifted posterior weights, exp-free tail sum, closed-form
elpd_i.  Dev tool; not imported by the product."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import psis_oracle as orc

EPS = np.finfo(float).eps

def gpdfit_fast(t, small_thr=2.0**-6, stats=None):
    n = len(t); m = 30 + int(n ** 0.5)
    j = np.arange(1, m + 1, dtype=float)
    b = 1 - np.sqrt(m / (j - 0.5)); b /= 3 * t[int(n / 4 + 0.5) - 1]; b += 1 / t[-1]
    # product form with exponent renormalisation
    f = 1.0 + (-b[:, None]) * t[None, :]          # fma in the kernel
    mant, ex = np.frexp(f)
    # sequential product emulation (chunks of 4 then frexp)
    P = np.ones(m); E = np.zeros(m)
    fmax = 1.0 + np.abs(b).max() * t[-1]
    R = 8 if fmax < 2.0**120 else (4 if fmax < 2.0**250 else 1)
    if stats is not None: stats.setdefault("R%d" % R, 0); stats["R%d" % R] += 1
    for i in range(n):
        P = P * f[:, i]
        if (i % R) == R - 1:
            mm, ee = np.frexp(P); P = mm; E += ee
    mm, ee = np.frexp(P); P = mm; E += ee
    ksum = np.log(P) + E * np.log(2.0)
    # guard: small |b| * sum(t) -> exact log1p path
    small = np.abs(b) * t.sum() < small_thr
    if stats is not None: stats["small"] += int(small.sum())
    if small.any():
        ksum[small] = np.log1p(-b[small, None] * t).sum(axis=1)
    kj = ksum / n
    L = n * (np.log(-(b / kj)) - kj - 1)
    if not np.all(np.isfinite(L)):
        if stats is not None: stats["nonfinite_L"] += 1
        w = 1 / np.exp(L - L[:, None]).sum(axis=1)
    else:
        e = np.exp(L - L.max()); w = e / e.sum()
    keep = w >= 10 * EPS
    w = w[keep]; bb = b[keep]; w = w / w.sum()
    bp = np.sum(bb * w)
    kp = np.log1p(-bp * t).mean()
    sigma = -kp / bp
    return (n * kp + 5) / (n + 10), sigma

def loo_row_fast(ll, M, stats):
    r = -ll
    mx = r.max(); x = r - mx
    srt = np.sort(x)
    c = max(srt[-M - 1], orc.CUTOFFMIN)
    ec = np.exp(c)
    tail = x > c
    n = int(tail.sum())
    xt = np.sort(x[tail])
    body_sum = np.exp(x[~tail]).sum()
    k = np.inf; smoothed = None
    if n > 4:
        t = np.exp(xt) - ec
        k, sigma = gpdfit_fast(t, stats=stats)
        if np.isfinite(k):
            p = (np.arange(n) + 0.5) / n
            q = orc.gpinv(p, k, sigma)
            y = q + ec
            smoothed = np.minimum(np.log(y), 0.0)
            ysum = np.minimum(y, 1.0).sum()
    if smoothed is None:
        smoothed = xt; ysum = np.exp(xt).sum()
    lse = np.log(body_sum + ysum)
    # elpd closed form
    vb = -mx - lse
    d = smoothed - xt
    dmax = max(0.0, d.max()) if n else 0.0
    tot = (len(x) - n) * np.exp(-dmax) + np.exp(d - dmax).sum()
    elpd = vb + dmax + np.log(tot)
    return elpd, k, lse

def main():
    rng = np.random.default_rng(5)
    stats = {"small": 0, "nonfinite_L": 0}
    for name, gen, S, reff, N in (("normal", lambda s: -1.4 + rng.normal(size=s), 4000, 0.9, 300),
                                  ("student", lambda s: -rng.standard_t(1.5, size=s), 8000, 1.0, 200),
                                  ("wide", lambda s: 3.0 * rng.normal(size=s), 2000, 1.0, 200),
                                  ("t3", lambda s: -rng.standard_t(3.0, size=s), 4000, 1.0, 200),
                                  ("lognorm", lambda s: -np.exp(rng.normal(size=s)), 1000, 0.7, 200),
                                  ("narrow", lambda s: -1000.3 + 1e-3 * rng.normal(size=s), 4000, 1.0, 100)):
        M = orc.tail_length(S, reff)
        worst = np.zeros(3)
        with np.errstate(all="ignore"):
            for i in range(N):
                ll = gen(S)
                lw, k = orc.psislw_row(-ll, M)
                elpd = orc.logsumexp_row(lw + ll)
                e2, k2, lse2 = loo_row_fast(ll, M, stats)
                if np.isfinite(k):
                    worst[0] = max(worst[0], abs(k2 - k) / abs(k))
                else:
                    assert not np.isfinite(k2)
                worst[1] = max(worst[1], abs(e2 - elpd) / abs(elpd))
        print(name, "max rel err k, elpd:", worst[:2], stats)

main()

def debug():
    rng = np.random.default_rng(5)
    S=8000; M=orc.tail_length(S,1.0)
    stats = {"small": 0, "nonfinite_L": 0}
    with np.errstate(all="ignore"):
        for i in range(200):
            ll = -rng.standard_t(1.5, size=S)
            lw, k = orc.psislw_row(-ll, M)
            e2,k2,_ = loo_row_fast(ll, M, stats)
            if np.isfinite(k) and abs(k2-k)/abs(k) > 1e-10:
                r=-ll; x=r-r.max(); srt=np.sort(x); c=max(srt[-M-1], orc.CUTOFFMIN)
                xt=np.sort(x[x>c]); t=np.exp(xt)-np.exp(c)
                print(i, k, k2, "n", len(t), "t range", t[0], t[int(len(t)/4+.5)-1], t[-1])
                n=len(t); m=30+int(n**.5)
                j=np.arange(1,m+1,dtype=float); b=1-np.sqrt(m/(j-.5)); b/=3*t[int(n/4+.5)-1]; b+=1/t[-1]
                f=1+(-b[:,None])*t[None,:]
                print(" f max", f.max(), "f min", f.min())
                break
