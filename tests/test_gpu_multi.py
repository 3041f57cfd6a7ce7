"""GPU parity of the multi-asset path (cf_dlm.cuh): MultiDisplaced x {MultiStats, Autocall, Baskets}
through the drop-in API, against (a) the reference's shipped golden spreadsheets, (b) the reference
compiled with g++ run live.  Tolerances: prices 1e-10 relative, AAD risks 1e-8 relative (entries below
1e-6 in absolute value: 1e-12 absolute)."""
import json
import os

import numpy as np
import pytest

from conftest import rel_err
from test_multi_oracle import X, put_dlm, named

pytestmark = pytest.mark.gpu
PRICE_TOL, RISK_TOL = 1e-10, 1e-8


def check_risks(got, want):
    got, want = np.asarray(got), np.asarray(want)
    big = np.abs(want) > 1e-6
    assert rel_err(got[big], want[big]) < RISK_TOL
    assert np.max(np.abs(got - want)[~big], initial=0.0) < 1e-12


def config5(api, model_id="dlm5", product_id="auto5", n_assets=10):
    a = np.arange(n_assets)
    spots = 100.0 + 5 * a
    atms = 0.20 + 0.02 * a
    skews = np.where(a % 3 == 0, 0.0, -0.05 * (a % 3))
    correl = np.full((n_assets, n_assets), 0.5) + 0.5 * np.eye(n_assets)
    api.put_displaced(spots, atms, skews, 0.02, 0.001 * a, [0.5, 1.5], np.full((2, n_assets), 0.01), correl, 0.25, model_id)
    api.put_autocall(spots, 3.0, 12, 1.0, 0.7, 0.10, 0.01, product_id)
    return spots


def test_shipped_golden_test_dlm_xlsx(cf):
    """27 moments of the displaced model with dividends, 10^6 Sobol paths (testDLM.xlsx U16:U42)."""
    g = X["test_dlm"]
    put_dlm(cf, g, "dlm_t")
    cf.put_multistats(3, g["fix_dates"], g["fwd_dates"], "stats_t")
    assert named(cf.payoff_labels("stats_t"), ["google", "amazon", "starbucks"]) == g["labels"]
    got = cf.value("dlm_t", "stats_t", g["n_paths"])
    assert rel_err(got, g["values"]) < PRICE_TOL


def test_shipped_golden_autocall_pricer_xlsx(cf):
    """Price and 25 AAD risks of the 3-asset autocallable, 100,000 Sobol paths (AutocallPricer.xlsx R15, R20:R45)."""
    g = X["autocall_pricer"]
    put_dlm(cf, g, "dlm_x")
    cf.put_autocall(g["spots"], g["maturity"], g["periods"], g["ko"], g["strike"], g["cpn"], g["smooth"], "auto_x")
    assert cf.payoff_labels("auto_x")[0] == g["payoff_label"]
    assert named(cf.param_labels("dlm_x"), ["uber", "lyft", "luckin"]) == g["risk_labels"]
    assert abs(cf.value("dlm_x", "auto_x", g["n_paths"])[0] / g["price"] - 1) < PRICE_TOL
    pv, rv, risks = cf.aad_risk_one("dlm_x", "auto_x", g["n_paths"])
    assert abs(rv / g["risk_value"] - 1) < PRICE_TOL
    check_risks(risks, g["risks"])


@pytest.mark.parametrize("sobol,n", [(True, 1 << 14), (False, 1 << 14), (False, 4097)])
def test_config5_autocall_10_assets_vs_reference(cf, ref, sobol, n):
    config5(cf); config5(ref)
    want = ref.value("dlm5", "auto5", n, sobol=sobol)
    assert rel_err(cf.value("dlm5", "auto5", n, sobol=sobol), want) < PRICE_TOL
    pv, rv, risks = cf.aad_risk_one("dlm5", "auto5", n, sobol=sobol)
    pv_r, rv_r, risks_r = ref.aad_risk_one("dlm5", "auto5", n, sobol=sobol)
    assert abs(rv / rv_r - 1) < PRICE_TOL
    assert risks.size == 107
    check_risks(risks, risks_r)


@pytest.mark.parametrize("n_assets", [1, 5, 8, 9, 13, 16])
def test_every_asset_count_bucket_vs_reference(cf, ref, n_assets):
    """The kernel is instantiated for up to 4 / 8 / 12 / 16 assets: asset counts at and inside the bucket limits
    (a full last bucket, a nearly empty one, a single asset), autocall and basket, value + AAD risks."""
    mid, pid, bid = f"dlm_b{n_assets}", f"auto_b{n_assets}", f"bsk_b{n_assets}"
    for api in (cf, ref):
        spots = config5(api, mid, pid, n_assets=n_assets)
        api.put_baskets(np.full(n_assets, 1.0 / n_assets), 2.0, [0.9 * spots.mean(), 1.1 * spots.mean()], bid)
    n = 4096
    assert rel_err(cf.value(mid, pid, n, sobol=False), ref.value(mid, pid, n, sobol=False)) < PRICE_TOL
    pv, rv, risks = cf.aad_risk_one(mid, pid, n, sobol=False)
    pv_r, rv_r, risks_r = ref.aad_risk_one(mid, pid, n, sobol=False)
    assert abs(rv / rv_r - 1) < PRICE_TOL and risks.size == risks_r.size
    check_risks(risks, risks_r)
    pv, rv, risks = cf.aad_risk_aggregate(mid, bid, [1.0, -0.5], n)
    pv_r, rv_r, risks_r = ref.aad_risk_aggregate(mid, bid, [1.0, -0.5], n)
    assert rel_err(pv, pv_r) < PRICE_TOL and abs(rv / rv_r - 1) < PRICE_TOL
    check_risks(risks, risks_r)


def _dlm_case(api, name):
    """Odd corners of the displaced model and the autocallable; returns (model id, product id)."""
    a = np.arange(4)
    spots, atms = 100.0 + 10 * a, 0.2 + 0.03 * a
    skews = np.array([0.0, -0.2, 0.15, -0.05])            # all four dynamics
    correl = np.full((4, 4), 0.3) + 0.7 * np.eye(4)
    kw = dict(spots=spots, atms=atms, skews=skews, rate=0.03, repo=0.002 * a, div_dates=[0.4, 1.1],
              divs=np.full((2, 4), 0.015), correl=correl, lam=0.2)
    auto = dict(refs=spots, maturity=2.0, periods=8, ko=1.0, strike=0.8, cpn=0.08, smooth=0.02)
    if name == "no_dividends": kw.update(div_dates=[], divs=np.zeros((0, 4)))
    elif name == "dividend_on_an_event_date": kw.update(div_dates=[0.25, 1.0])
    elif name == "independent_assets": kw.update(correl=np.eye(4), lam=0.0)
    elif name == "lambda_pushes_to_full_correlation": kw.update(lam=0.95)
    elif name == "zero_rates": kw.update(rate=0.0, repo=np.zeros(4))
    elif name == "one_period": auto.update(periods=1)
    elif name == "knock_out_far_away": auto.update(ko=3.0)
    elif name == "knocked_out_on_the_first_date": auto.update(ko=0.2)
    elif name == "thin_smoothing": auto.update(smooth=1e-6)
    elif name == "no_smoothing": auto.update(smooth=0.0)      # digital knock-out: the adjoint of the smoothing ratio is 0 / 0 unless skipped
    elif name == "strike_above_par": auto.update(strike=1.3)
    elif name == "references_off_the_spots": auto.update(refs=spots * np.array([0.8, 1.0, 1.2, 1.05]))
    api.put_displaced(kw["spots"], kw["atms"], kw["skews"], kw["rate"], kw["repo"], kw["div_dates"], kw["divs"], kw["correl"],
                      kw["lam"], "dlm_odd")
    api.put_autocall(auto["refs"], auto["maturity"], auto["periods"], auto["ko"], auto["strike"], auto["cpn"], auto["smooth"], "auto_odd")
    return "dlm_odd", "auto_odd"


DLM_CASES = ["no_dividends", "dividend_on_an_event_date", "independent_assets", "lambda_pushes_to_full_correlation", "zero_rates",
             "one_period", "knock_out_far_away", "knocked_out_on_the_first_date", "thin_smoothing", "no_smoothing", "strike_above_par",
             "references_off_the_spots"]


@pytest.mark.parametrize("case", DLM_CASES)
def test_odd_displaced_models_and_autocalls_vs_reference(cf, ref, case):
    for api in (cf, ref):
        mid, pid = _dlm_case(api, case)
    n = 4096 + 3
    got, want = cf.simul_paths(mid, pid, n, sobol=False), ref.simul_paths(mid, pid, n, sobol=False)
    assert np.max(np.abs(got - want)) < 1e-11
    pv, rv, risks = cf.aad_risk_one(mid, pid, n)
    pv_r, rv_r, risks_r = ref.aad_risk_one(mid, pid, n)
    assert abs(rv / rv_r - 1) < PRICE_TOL and risks.size == risks_r.size
    check_risks(risks, risks_r)


def test_per_path_payoffs_autocall(cf, ref):
    config5(cf); config5(ref)
    got = cf.simul_paths("dlm5", "auto5", 777, sobol=False)
    want = ref.simul_paths("dlm5", "auto5", 777, sobol=False)
    assert np.max(np.abs(got - want)) < 1e-12


def test_baskets_vs_reference(cf, ref):
    """Strike ladder on a weighted basket of 4 assets with all four dynamics, value + aggregate AAD risk."""
    spots = [100.0, 50.0, 80.0, 120.0]
    atms = [0.2, 0.3, 0.25, 0.2]
    skews = [0.0, -0.15, -0.1, 0.02]          # lognormal, normal (beta = 0), surnormal, surnormal
    correl = np.array([[1, .4, .2, .1], [.4, 1, .3, .0], [.2, .3, 1, -.2], [.1, .0, -.2, 1]])
    strikes = np.arange(60.0, 121.0, 5.0)
    for api in (cf, ref):
        api.put_displaced(spots, atms, skews, 0.03, [0.0, 0.005, 0.0, 0.01], [0.3], [[0.02, 0.0, 0.01, 0.0]], correl, 0.1, "dlm4")
        api.put_baskets([0.25, 0.5, 0.3125, 0.2], 1.0, strikes, "bsk")
    n = 1 << 14
    assert rel_err(cf.value("dlm4", "bsk", n), ref.value("dlm4", "bsk", n)) < PRICE_TOL
    notionals = np.linspace(1.0, 2.0, strikes.size)
    pv, rv, risks = cf.aad_risk_aggregate("dlm4", "bsk", notionals, n, sobol=False)
    pv_r, rv_r, risks_r = ref.aad_risk_aggregate("dlm4", "bsk", notionals, n, sobol=False)
    assert rel_err(pv, pv_r) < PRICE_TOL and abs(rv / rv_r - 1) < PRICE_TOL
    check_risks(risks, risks_r)


def test_multistats_has_no_device_aad(cf):
    g = X["test_dlm"]
    put_dlm(cf, g, "dlm_t")
    cf.put_multistats(3, g["fix_dates"], g["fwd_dates"], "stats_t")
    with pytest.raises(RuntimeError):
        cf.aad_risk_one("dlm_t", "stats_t", 1024)


@pytest.mark.parametrize("sobol", [False, True])
def test_arbitrary_shard_boundaries_add_up_displaced_model(cf, sobol):
    """Ranges of paths through the C ABI (first_path, n_paths) add up to the whole run whatever the boundaries: odd ones
    split antithetic pairs of mrg32k3a across ranges (the kernel's lanes then no longer share a generator) and cut Sobol
    windows; a range of one path; ranges longer than one batch."""
    config5(cf)
    n = 5000
    nadj = cf.describe("dlm5", "auto5", aad=True)["adjoint_size"]
    w = [1.0]
    whole = cf.run_range("dlm5", "auto5", 0, n, w, sobol=sobol, n_adjoints=nadj)
    cuts = [0, 1, 1001, 1002, 2345, 4999, n]
    parts = [cf.run_range("dlm5", "auto5", a, b - a, w, sobol=sobol, n_adjoints=nadj) for a, b in zip(cuts[:-1], cuts[1:])]
    assert abs(sum(p[0][0] for p in parts) / whole[0][0] - 1) < 1e-13
    assert abs(sum(p[1] for p in parts) / whole[1] - 1) < 1e-13
    tot = sum(p[2] for p in parts)
    assert np.max(np.abs(tot - whole[2])) < 1e-11 * np.max(np.abs(whole[2]))
    vals = [cf.run_range("dlm5", "auto5", a, b - a, sobol=sobol)[0][0] for a, b in zip(cuts[:-1], cuts[1:])]
    assert abs(sum(vals) / whole[0][0] - 1) < 1e-13
