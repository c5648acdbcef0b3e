import sys, os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np
import newtonkrylov_jl_b200 as nk, oracle as O, problems as P
ctx = nk.get_context(0)
for name, d, kw in [("bratu2d_32", P.bratu2d(32), {}), ("bratu1d_200", P.bratu1d(200), {}), ("b2d_64_fixed", P.bratu2d(64), dict(forcing=nk.Fixed(0.1)))]:
    for native in (False, True):
        F_, u, p, _ = P.device_setup(nk, ctx, d)
        hist = []
        fn = nk.newton_krylov_native_ if native else nk.newton_krylov_
        _, r = fn(F_, u, p, None, history=hist, **kw)
        po = P.oracle_problem(O, d)
        o = nk.host._newton_opts(1e-6, 1e-12, 50, kw.get("forcing", nk.EisenstatWalker()), "gmres", 20, 0, {})
        ur, sr, hr = O.newton(po, d["u0"], o)
        print(name, "native" if native else "host", r.solved, r.stats, sr["outer_iterations"], sr["inner_iterations"])
        for a, b in zip(hist, hr):
            print("   ", a["inner"], b["inner"], "%.15e %.15e %.2e" % (a["n_res"], b["n_res"], abs(a["n_res"]-b["n_res"])/b["n_res"]), a["eta"], b["eta"])
