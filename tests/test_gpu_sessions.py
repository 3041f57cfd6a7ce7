"""Resident sessions of the host entry points (cf_base.h: clone + tape + device plan kept per stored (model, product,
RNG)), the bit-reproducible itemised risk, and the resident-plan host-buffer runs of the C ABI."""
import ctypes as C

import numpy as np
import pytest

from conftest import config3_surface

pytestmark = pytest.mark.gpu


def _risk(cf, n=20_000, w=(0.7, 0.3)):
    v, d, vega = cf.dupire_aad_risk("ses_dup", "ses_uoc", list(w), 30, 36, n)
    return np.concatenate([[v, d], vega.ravel()])


def test_session_reuse_and_invalidation(cf, ref):
    spots, times, vols = config3_surface()
    cf.put_dupire(100.0, spots, times, vols, 0.25, "ses_dup")
    cf.put_barrier(120.0, 150.0, 3.0, 1.0 / 52, 0.01, False, "ses_uoc")
    first = _risk(cf)                       # builds the session
    again = _risk(cf)                       # reuses it: clone, tape and plan resident
    assert np.array_equal(first, again)
    other_w = _risk(cf, w=(0.0, 1.0))       # same session, other notionals
    assert not np.array_equal(first, other_w)
    # against the reference, through the reused session
    ref.put_dupire(100.0, spots, times, vols, 0.25, "ses_dup")
    ref.put_barrier(120.0, 150.0, 3.0, 1.0 / 52, 0.01, False, "ses_uoc")
    rv, rd, rvega = ref.dupire_aad_risk("ses_dup", "ses_uoc", [0.0, 1.0], 30, 36, 20_000)
    assert abs(other_w[0] / rv - 1) < 1e-10 and abs(other_w[1] / rd - 1) < 1e-8
    # putting the model again under the same name invalidates the session
    cf.put_dupire(100.0, spots, times, vols * 1.01, 0.25, "ses_dup")
    bumped = _risk(cf)
    assert abs(bumped[0] / first[0] - 1) > 1e-4
    cf.put_dupire(100.0, spots, times, vols, 0.25, "ses_dup")
    assert np.array_equal(_risk(cf), first)
    # and so does putting the product again
    cf.put_barrier(120.0, 155.0, 3.0, 1.0 / 52, 0.01, False, "ses_uoc")
    assert abs(_risk(cf)[0] / first[0] - 1) > 1e-4
    cf.put_barrier(120.0, 150.0, 3.0, 1.0 / 52, 0.01, False, "ses_uoc")
    assert np.array_equal(_risk(cf), first)
    # value() has its own session
    a = cf.value("ses_dup", "ses_uoc", 20_000)
    assert np.array_equal(a, cf.value("ses_dup", "ses_uoc", 20_000)) and abs(a[0] / first[0] - 1) < 1.0   # finite, same paths


def test_itemised_risk_is_bit_reproducible(cf):
    """AADriskMulti of config 4's shape: the strike-class tables are accumulated in fixed point with integer atomics."""
    spots, times, vols = config3_surface()
    cf.put_dupire(100.0, spots, times, vols, 0.25, "ses_dup4")
    mats = np.repeat(0.25 * np.arange(1, 13), 60)
    strikes = np.tile(70.5 + np.arange(60), 12)
    cf.put_europeans(mats, strikes, "ses_eurs")
    for sobol in (False, True):
        v1, r1 = cf.aad_risk_multi("ses_dup4", "ses_eurs", 1 << 14, sobol=sobol)
        v2, r2 = cf.aad_risk_multi("ses_dup4", "ses_eurs", 1 << 14, sobol=sobol)
        assert np.array_equal(v1, v2) and np.array_equal(r1, r2)
        assert np.all(np.isfinite(r1)) and np.max(np.abs(r1)) > 0


def test_itemised_risk_has_no_payoff_cap(cf, ref):
    """One sweep per payoff for pairs without a strike-class kernel: more than 64 payoffs (Baskets ladder)."""
    for api in (cf, ref):
        api.put_displaced([100.0, 90.0], [0.2, 0.25], [0.0, -0.05], 0.02, [0.0, 0.001], [0.5], np.full((1, 2), 0.01),
                          np.array([[1.0, 0.4], [0.4, 1.0]]), 0.25, "ses_dlm")
        api.put_baskets([0.5, 0.5], 1.0, 60.0 + np.arange(70.0), "ses_bask")
    v, r = cf.aad_risk_multi("ses_dlm", "ses_bask", 4096, sobol=True)
    assert r.shape[1] == 70
    for k in (0, 33, 69):
        pv, rv, one = ref.aad_risk_one("ses_dlm", "ses_bask", 4096, risk_payoff=k, sobol=True)
        scale = max(1e-6, float(np.max(np.abs(one))))
        assert np.max(np.abs(r[:, k] - one)) < 1e-8 * scale


def test_resident_plan_host_buffer_runs(eng):
    """cf_plan_run_value / cf_plan_run_aad: the one-shot runs without re-uploading the tables."""
    from test_gpu_parity import _config3_lowlevel
    tab, mdl, prd = _config3_lowlevel(eng)
    rng = eng.rng("sobol")
    n, w = 12_345, [1.0, 0.25]
    one = eng.run_aad(mdl, prd, rng, 7, n, w)
    plan = C.c_void_p()
    eng._chk(eng.lib.cf_plan_create(C.byref(mdl), C.byref(prd), C.byref(rng), C.byref(plan)))
    dp = C.POINTER(C.c_double)
    sums, agg, adj = np.zeros(2), np.zeros(1), np.zeros(one["table_adj"].size)
    wv = (C.c_double * 2)(*w)
    for _ in range(2):
        eng._chk(eng.lib.cf_plan_run_aad(plan, wv, 7, n, sums.ctypes.data_as(dp), agg.ctypes.data_as(dp), adj.ctypes.data_as(dp), None, None))
        assert np.array_equal(sums, one["payoff_sums"]) and agg[0] == one["agg_sum"] and np.array_equal(adj, one["table_adj"])
    vs = np.zeros(2)
    eng._chk(eng.lib.cf_plan_run_value(plan, 7, n, vs.ctypes.data_as(dp), None))
    assert np.array_equal(vs, eng.run_value(mdl, prd, rng, 7, n))
    eng.lib.cf_plan_destroy(plan)
    f, c = eng.shard_range(1 << 20, 3, 8)
    assert (f, c) == (3 << 17, 1 << 17)
