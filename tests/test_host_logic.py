"""CPU tests of the host mirror (compfinance_b200/host): timelines, init() tables, the time map read
off the host tape, labels, store and error behaviour -- everything that happens before the mark."""
import numpy as np
import pytest

from oracle import restate as R
from conftest import config3_surface, put_config3


@pytest.fixture(scope="module")
def api(built):
    from compfinance_b200.api import CompFinance
    return CompFinance()


def test_dupire_tables_match_oracle(api):
    spots, times, vols = put_config3(api)
    ptl = R.uoc_timeline(3.0, 1.0 / 52)
    assert (api.product_timeline("uoc") == np.array(ptl)).all()
    tab = R.DupireTables(100.0, spots, times, vols, 0.25, ptl)
    for aad in (False, True):
        d = api.describe("dup", "uoc", aad=aad)
        assert (d["n_steps"], d["n_events"], d["n_knots"]) == (156, 157, 30)
        assert (d["tab_a"] == tab.interp_vols).all() and (d["tab_b"] == tab.log_spots).all()
        assert (d["is_event"] == np.array(tab.common, dtype=np.uint8)).all()
        assert d["smooth"] == float(np.exp(np.log(100.0)) * 0.01) and d["first_sample_is_today"]
    d = api.describe("dup", "uoc", aad=True)
    assert d["adjoint_size"] == 1081 and d["n_times"] == 36
    nT, c1, c2, w1, w2 = tab.time_map()
    _, k1, k2, v1, v2 = d["time_map"]
    # same linear map (a zero weight may sit on either column)
    dense_o, dense_h = np.zeros((156, nT)), np.zeros((156, nT))
    for i in range(156):
        dense_o[i, c1[i]] += w1[i]; dense_o[i, c2[i]] += w2[i]
        dense_h[i, k1[i]] += v1[i]; dense_h[i, k2[i]] += v2[i]
    assert np.max(np.abs(dense_o - dense_h)) < 1e-16


def test_dupire_fill_steps(api):
    spots, times, vols = config3_surface()
    api.put_dupire(100.0, spots, times, vols, 0.25, "dupc")
    api.put_european(110.0, 1.0, 1.0, "eur1y")
    d = api.describe("dupc", "eur1y")
    tab = R.DupireTables(100.0, spots, times, vols, 0.25, [1.0])
    assert d["n_steps"] == 4 and d["is_event"].tolist() == [0, 0, 0, 0, 1]
    assert (d["tab_a"] == tab.interp_vols).all()


def test_bs_tables_match_oracle(api):
    api.put_black_scholes(100, 0.15, False, 0.03, 0.01, "bs")
    api.put_european(100, 1.0, 1.25, "eur")
    tb = R.BSTables(100, 0.15, 0.03, 0.01, [1.0], [1.25], [1.25], [True])
    d = api.describe("bs", "eur", aad=True)
    assert (d["tab_a"] == tb.drifts).all() and (d["tab_b"] == tb.stds).all()
    assert d["numeraires"][0] == tb.numeraires[0] and d["fwd_factors"][0] == tb.fwd_factors[0]
    assert d["discounts"][0] == tb.discounts[0] and d["adjoint_size"] == 1 + 2 + 4      # spot, drift, std, numeraire, fwd factor, discount, libor
    api.put_barrier(100, 120, 1.0, 1.0 / 52, 0.01, False, "uoc1y")
    ptl = R.uoc_timeline(1.0, 1.0 / 52)
    tb2 = R.BSTables(100, 0.15, 0.03, 0.01, ptl, ptl, [None] * len(ptl), [False] * (len(ptl) - 1) + [True])
    d = api.describe("bs", "uoc1y")
    assert d["n_steps"] == 52 and d["n_events"] == 53 and d["is_event"][0] == 1
    assert (d["tab_a"] == tb2.drifts).all() and (d["numeraires"] == tb2.numeraires).all()
    assert d["smooth"] == 1.0


def test_labels_and_store(api, ref):
    put_config3(api)
    put_config3(ref)
    assert api.payoff_labels("uoc") == ref.labels("uoc")
    assert api.param_labels("dup") == ref.labels("dup", params=True)
    api.put_black_scholes(100, 0.15, False, 0.03, 0.01, "bs")
    ref.put_bs(100, 0.15, False, 0.03, 0.01, "bs")
    api.put_european(100, 1.0, 1.25, "eur")
    ref.put_european(100, 1.0, 1.25, "eur")
    assert api.payoff_labels("eur") == ref.labels("eur") and api.param_labels("bs") == ref.labels("bs", params=True)
    api.put_europeans([0.5, 0.5, 1.0], [90.0, 100.0, 100.0], "eurs")
    ref.put_europeans([0.5, 0.5, 1.0], [90.0, 100.0, 100.0], "eurs")
    assert api.payoff_labels("eurs") == ref.labels("eurs") and api.num_payoffs("eurs") == 3


def test_errors(api):
    from compfinance_b200.api import CfHostError
    with pytest.raises(CfHostError, match="not found|Could not retrieve"):
        api.value("nope", "nothing", 100)
    with pytest.raises(CfHostError):
        api.num_payoffs("nothing")
    api.put_black_scholes(100, 0.15, True, 0.0, 0.0, "bs_spot_measure")
    api.put_european(100, 1.0, 1.0, "eur")
    with pytest.raises(CfHostError, match="no device image"):
        api.describe("bs_spot_measure", "eur")


def test_dupire_calibration_matches_reference(api, ref):
    """dupireCalib (main.h:413, mcMdlDupire.h:289-388, ivs.h): host only, Merton surface of BASELINE config 4."""
    args = ([50.0, 100.0, 200.0], 5.0, [0.25, 3.0], 1.0 / 12, 100.0, 0.15, 0.05, -0.15, 0.10)
    s, t, lv = api.dupire_calib(*args)
    s_r, t_r, lv_r = ref.dupire_calib(*args)
    assert np.array_equal(s, s_r) and np.array_equal(t, t_r)
    assert lv.shape == lv_r.shape and np.max(np.abs(lv / lv_r - 1)) < 1e-12


def test_time_map_recorded_by_init_equals_the_one_read_off_the_tape(built):
    """Dupire::init() records the step -> (two time columns, weights) map itself; under CF_CHECK_TIME_MAP
    deviceImage() also derives it from the leaf gradients of the tape and throws on any difference.  Cases:
    config 3, a surface whose time knots coincide with simulation dates (t == 0: one parent only) and one
    whose knots start after today and end before maturity (flat extrapolation on both sides)."""
    import os
    import subprocess
    import sys
    code = r'''
import numpy as np
from compfinance_b200.api import CompFinance
cf = CompFinance()
spots = np.arange(55, 201, 5.0)
cases = {"cfg3": np.arange(1, 37) / 12.0, "on_dates": np.arange(0, 13) * 0.25, "flat_both": np.array([0.6, 1.1, 1.9])}
cf.put_barrier(120.0, 150.0, 3.0, 1.0 / 52, 0.01, False, "uoc")
cf.put_european(110.0, 2.0, 2.0, "eur")
for name, times in cases.items():
    vols = 0.15 + 0.10 * np.log(spots[:, None] / 100.0) ** 2 + 0.02 * times[None, :]
    cf.put_dupire(100.0, spots, times, vols, 0.25, name)
    for prd in ("uoc", "eur"):
        d = cf.describe(name, prd, aad=True)
        assert d["n_times"] == times.size and d["adjoint_size"] == 1 + spots.size * times.size, (name, prd)
        nT, c1, c2, w1, w2 = d["time_map"]
        assert (c1 <= c2).all() and (w2[c1 == c2] == 0).all()
        # the map reproduces the table: interpVols[i][j] = w1 vols[j][c1] + w2 vols[j][c2]
        rebuilt = w1[:, None] * vols[:, c1].T + w2[:, None] * vols[:, c2].T
        assert np.max(np.abs(rebuilt - d["tab_a"])) < 1e-15, (name, prd)
print("ok")
'''
    env = dict(os.environ, CF_CHECK_TIME_MAP="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
