"""diagnostic: host-mirror HSD vs device-resident HSD on config 5 at full size (iteration log side by side)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import tlpb200_loader
pkg = tlpb200_loader.load()
from tulip_jl_b200 import hsd, lpgen

lp = lpgen.config(5)
refine = int(os.environ.get("TLPB200_DC_REFINE", "2"))
kh = pkg.setup(lp.A, pkg.K1(), pkg.Backend())
h = hsd.HSD(lp.A, lp.b, lp.c, lp.l, lp.u, kh)
h.optimize(max_iter=30)
kd = pkg.setup(lp.A, pkg.K1(), pkg.Backend())
d = pkg.DeviceHSD(kd, lp.b, lp.c, lp.l, lp.u, params=hsd.IPMOptions(IterationsLimit=30))
d.optimize()
print("dc_refine", refine, "host", h.status, h.niter, h.n_update, h.n_solve, "device", d.status, d.niter, d.n_update, d.n_solve)
hl, dl = h.log, d.log
for i in range(max(len(hl), len(dl))):
    a = hl[i] if i < len(hl) else None
    b = dl[i] if i < len(dl) else None
    fmt = lambda r: "%2d pobj %.10e dobj %.10e pf %.2e df %.2e mu %.1e" % (r[0], r[1], r[2], r[3], r[4], r[6]) if r else "-"
    print("H", fmt(a), "| D", fmt(b))

# state-leak checks: (a) device loop twice on the same fresh handle, (b) host IPM then device loop on the same handle
k2 = pkg.setup(lp.A, pkg.K1(), pkg.Backend())
d2 = pkg.DeviceHSD(k2, lp.b, lp.c, lp.l, lp.u, params=hsd.IPMOptions(IterationsLimit=40))
for rep in range(3):
    d2.optimize()
    print("fresh handle, device run", rep, d2.status, d2.niter, "%.10e" % d2.primal_objective)
k3 = pkg.setup(lp.A, pkg.K1(), pkg.Backend())
h3 = hsd.HSD(lp.A, lp.b, lp.c, lp.l, lp.u, k3)
h3.optimize(max_iter=40)
print("same handle: host run", h3.status, h3.niter)
d3 = pkg.DeviceHSD(k3, lp.b, lp.c, lp.l, lp.u, params=hsd.IPMOptions(IterationsLimit=40))
for rep in range(2):
    d3.optimize()
    print("same handle after host run, device run", rep, d3.status, d3.niter, "%.10e" % d3.primal_objective)
    for r in d3.log[14:22]:
        print("   ", r)
h4 = hsd.HSD(lp.A, lp.b, lp.c, lp.l, lp.u, k3)
h4.optimize(max_iter=40)
print("same handle: host run again", h4.status, h4.niter)
