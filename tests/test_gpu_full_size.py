"""BASELINE.json configs 4 and 5 at their full sizes (2^22 mrg32k3a paths) on the GPU: against the reference where the
CPU finishes in seconds (config 5), through size-independent properties where it does not (config 4: linearity of the
aggregate risk in the notionals, monotone and convex call ladders, antithetic pairs)."""
import numpy as np
import pytest

from conftest import rel_err
from test_gpu_europeans import config4, check_multi
from test_gpu_multi import config5, check_risks

pytestmark = pytest.mark.gpu
N = 1 << 22


def test_config5_full_size_vs_reference(cf, ref):
    """10-asset autocallable, 2^22 paths, price and 107 AAD risks: the reference's own run on the host cores."""
    config5(cf); config5(ref)
    pv, rv, risks = cf.aad_risk_one("dlm5", "auto5", N, sobol=False)
    pv_r, rv_r, risks_r = ref.aad_risk_one("dlm5", "auto5", N, sobol=False)
    assert abs(rv / rv_r - 1) < 1e-10 and risks.size == 107
    check_risks(risks, risks_r)


def test_config4_full_size_properties(cf):
    """720 European payoffs on the Dupire surface, 2^22 paths."""
    npay = config4(cf)
    values = cf.value("dup4", "eurs4", N, sobol=False).reshape(12, 60)           # [maturity][strike 70.5 .. 129.5]
    assert (np.diff(values, axis=1) < 0).all()                                    # calls decrease in strike
    assert (np.diff(values, n=2, axis=1) > -1e-12).all()                          # and are convex in strike, path by path
    assert (np.diff(values, axis=0) > 0).all()                                    # and increase in maturity (no rates)
    # a second run reproduces the first bit for bit (fixed-order reductions)
    assert (cf.value("dup4", "eurs4", N, sobol=False).reshape(12, 60) == values).all()
    # the aggregate risk is linear in the notionals
    rng = np.random.default_rng(7)
    w1, w2 = rng.normal(size=npay), rng.normal(size=npay)
    pv1, rv1, r1 = cf.aad_risk_aggregate("dup4", "eurs4", w1, N, sobol=False)
    pv2, rv2, r2 = cf.aad_risk_aggregate("dup4", "eurs4", w2, N, sobol=False)
    pv3, rv3, r3 = cf.aad_risk_aggregate("dup4", "eurs4", 2.0 * w1 - 0.5 * w2, N, sobol=False)
    assert rel_err(pv1, values.ravel()) < 1e-13 and rel_err(pv3, values.ravel()) < 1e-13
    assert abs(rv3 - (2.0 * rv1 - 0.5 * rv2)) < 1e-11 * (abs(rv1) + abs(rv2))
    scale = np.max(np.abs(r1)) + np.max(np.abs(r2))
    assert np.max(np.abs(r3 - (2.0 * r1 - 0.5 * r2))) < 1e-10 * scale
    assert abs(rv1 - float(w1 @ pv1)) < 1e-11 * np.abs(w1 * pv1).sum()


def test_config4_itemised_risk_matrix_vs_reference_at_65536_paths(cf, ref):
    """AADriskMulti of config 4 (1081 parameters x 720 payoffs) against the reference's own matrix on 2^16 mrg32k3a
    paths -- the size the reference finishes in seconds on the box's host cores (its sweep is one pass per payoff)."""
    npay = config4(cf); config4(ref)
    risks = check_multi(cf, ref, "dup4", "eurs4", 1 << 16, False, [0, 59, 5 * 60 + 31, 719])
    assert risks.shape == (1081, npay)
