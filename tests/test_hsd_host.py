"""CPU tests of the host-side caller mirror (tulip.jl_b200/hsd.py) against the oracle's
restatement of the reference driver (oracle/hsd_ref.py) and the reference's golden answers.
The KKT backend plugged in here is the oracle's CPU one -- this tests host logic only."""
import numpy as np
import pytest

import tlpb200_loader
from golden.lpex import LPEX
from oracle import hsd_ref, kkt_ref

pkg = tlpb200_loader.load()
from tulip_jl_b200 import hsd, lpgen  # noqa: E402

TOL = 100 * float(np.sqrt(np.finfo(float).eps))


def _std(lp):
    return hsd_ref.standard_form(**{k: v for k, v in lp.items() if k != "expect"})


@pytest.mark.parametrize("name", list(LPEX))
def test_golden_lps_same_trajectory(name):
    lp = LPEX[name]
    dat = _std(lp)
    ref = hsd_ref.HSDRef(dat, kkt_ref.SparseK2(dat.A))
    ref.optimize()
    h = hsd.HSD(dat.A, dat.b, dat.c, dat.l, dat.u, kkt_ref.SparseK2(dat.A), c0=dat.c0, objsense=dat.objsense)
    st = h.optimize()
    assert st == lp["expect"]["status"] == ref.status
    assert h.niter == ref.niter
    a = np.array([r[1:] for r in h.log]); b = np.array([r[1:] for r in ref.log])
    np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-9)
    if "obj" in lp["expect"]:
        assert abs(h.primal_objective - lp["expect"]["obj"]) <= TOL * (1 + abs(lp["expect"]["obj"]))


@pytest.mark.parametrize("cfg", [2, 3])
def test_mini_config_same_trajectory(cfg):
    lp = lpgen.config(cfg, mini=True)
    dat = hsd_ref.IPMData(lp.A, lp.b, True, lp.c, 0.0, lp.l, lp.u)
    ref = hsd_ref.HSDRef(dat, kkt_ref.SparseK2(lp.A))
    ref.optimize()
    h = hsd.HSD(lp.A, lp.b, lp.c, lp.l, lp.u, kkt_ref.SparseK2(lp.A))
    h.optimize()
    assert h.status == ref.status == "Trm_Optimal"
    assert h.niter == ref.niter
    assert abs(h.primal_objective - ref.primal_objective) <= 1e-8 * (1 + abs(ref.primal_objective))
    assert h.n_solve == ref.n_solve and h.n_update == ref.n_update
    assert 3 <= min(h.solves_per_iter) and max(h.solves_per_iter) <= 6


def test_regularisation_bump_on_failure():
    """step.jl:34-51: a PosDefException from update! multiplies the regularisations by 100."""
    lp = lpgen.config(2, mini=True)

    class Flaky(kkt_ref.SparseK2):
        calls = 0

        def update(self, th, rp, rd):
            Flaky.calls += 1
            if Flaky.calls == 2:
                raise kkt_ref.PosDefException("injected")
            super().update(th, rp, rd)

    h = hsd.HSD(lp.A, lp.b, lp.c, lp.l, lp.u, Flaky(lp.A))
    seen = []
    h.on_update = lambda th, rp, rd: seen.append(rp[0])
    h.optimize(max_iter=3)
    assert h.n_update == h.niter + 1
    assert seen[2] == pytest.approx(seen[1] * 100)
