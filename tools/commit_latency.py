"""Where a demo-size commit's wall time goes: device phases (engine events) vs wall, pageable vs pinned host input."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import plonky2_demo_b200 as p
from helpers import seeded_polys

p.init(0)
for w, lg_d, fv in [(135, 15, True), (135, 15, False), (16, 15, False), (135, 12, True)]:
    x = seeded_polys(w, 1 << lg_d)
    pin = torch.empty((w, 1 << lg_d), dtype=torch.int64, pin_memory=True)
    pin.numpy().view(np.uint64)[:] = x
    xp = pin.numpy().view(np.uint64)
    for name, arr in (("pageable", x), ("pinned", xp)):
        f = (lambda t: p.PolynomialBatch.from_values(arr, 3, False, 4, t)) if fv else (lambda t: p.PolynomialBatch.from_coeffs(arr, 3, False, 4, t))
        for _ in range(3):
            f(None).free()
        ts, phases = [], None
        for _ in range(7):
            t = {}
            t0 = time.perf_counter()
            b = f(None)
            ts.append(1e3 * (time.perf_counter() - t0))
            b.free()
        t = {}
        b = f(t); b.free()
        print(w, lg_d, "from_values" if fv else "from_coeffs", name, "wall ms median %.2f" % sorted(ts)[3], {k: round(v, 2) for k, v in t.items()})
