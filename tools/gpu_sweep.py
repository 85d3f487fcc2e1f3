"""Launch-shape sweep (dev tool): time psislw / loo for env-selected NT / NBUF / CAP."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from pyloo_b200 import engine
    N, S = 40000, 4000
    torch.manual_seed(0)
    x = torch.randn(N, S, dtype=torch.float64, device="cuda")
    out = torch.empty_like(x)
    res = {}
    info = engine.row_launch_info(S, 200, "psislw")
    for name, fn in (("psislw", lambda: engine.psislw_cuda(x, 0.9, out=out)), ("loo_rows", lambda: engine.loo_cuda(x.t(), 1.0))):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(4): fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = N / (e0.elapsed_time(e1) / 4) * 1e3 / 1e6
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("B2L_")}, "launch": info, "Mobs_per_s": res}))
else:
    for env in ({}, {"B2L_NBUF": "2", "B2L_NT": "256"}, {"B2L_NBUF": "1", "B2L_NT": "128"}, {"B2L_NBUF": "1", "B2L_NT": "128", "B2L_CAP": "512"},
                {"B2L_NBUF": "1", "B2L_NT": "256", "B2L_CAP": "512"}, {"B2L_NBUF": "2", "B2L_NT": "128", "B2L_CAP": "512"}, {"B2L_NBUF": "1", "B2L_NT": "512"}):
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-400:])
