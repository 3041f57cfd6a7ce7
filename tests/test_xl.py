"""The Excel-facing wrappers (xlExport.cpp:72-1255 mirrored in compfinance_b200/host/cf_xl.h) called the way Excel
calls them: XLOPER12 / FP12 in, XLOPER12 out, #N/A on any error.  The first group needs no GPU (marshalling, the
object store, host-side analytics, error behaviour); the second compares the simulations with the flat API."""
import numpy as np
import pytest

import xl
from conftest import config3_surface


@pytest.fixture(scope="module")
def api(built):
    from compfinance_b200.api import CompFinance
    return CompFinance()


@pytest.fixture(scope="module")
def X(api):
    return xl.bind(api.lib)


def test_wrappers_are_exported(X):
    assert len(xl.SIGS) == 24


def test_put_functions_store_and_report_errors(X, api):
    k = []
    assert xl.read(X.xPutBlackScholes(100, 0.2, 0, 0.03, 0.01, xl.xstr("xbs", k))) == "xbs"
    assert xl.read(X.xPutBlackScholes(100, 0.2, 0, 0.03, 0.01, xl.xstr("", k))) == xl.NA          # no id
    assert xl.read(X.xPutEuropean(110, 2.0, 0.0, xl.xstr("xeur", k))) == "xeur"
    assert xl.read(X.xPutBarrier(100, 130, 1.0, 1.0 / 12, 0.01, xl.xstr("put", k), xl.xstr("xuop", k))) == "xuop"
    assert xl.read(X.xPutContingent(0.02, 2.0, 0.5, 0.01, xl.xstr("xcb", k))) == "xcb"
    assert xl.read(X.xPayoffIds(xl.xstr("xeur", k))) == [["call 110.00 2.00"]]
    ids = xl.read(X.xPayoffIds(xl.xstr("xuop", k)))
    assert len(ids) == 2 and ids[1][0].startswith("put ") and "up and out" in ids[0][0]
    assert xl.read(X.xPayoffIds(xl.xstr("nobody", k))) == xl.NA
    params = xl.read(X.xParameters(xl.xstr("xbs", k)))
    assert [r[0] for r in params] == list(api.param_labels("xbs")) and [r[1] for r in params] == [100, 0.2, 0.03, 0.01]
    # Europeans: blanks (<= EPS) are dropped pairwise, shapes must agree
    mats, strikes = [1.0, 1.0, 0.0, 2.0], [90.0, 110.0, 100.0, 0.0]
    assert xl.read(X.xPutEuropeans(xl.fp12(mats, k), xl.fp12(strikes, k), xl.xstr("xeurs", k))) == "xeurs"
    assert [r[0] for r in xl.read(X.xPayoffIds(xl.xstr("xeurs", k)))] == list(api.payoff_labels("xeurs")) and api.num_payoffs("xeurs") == 2
    assert xl.read(X.xPutEuropeans(xl.fp12(mats, k), xl.fp12(strikes[:3], k), xl.xstr("bad", k))) == xl.NA
    # Dupire: the vols range must be spots x times
    spots, times, vols = config3_surface()
    assert xl.read(X.xPutDupire(100, xl.fp12(spots, k), xl.fp12(times, k), xl.fp12(vols, k), 0.25, xl.xstr("xdup", k))) == "xdup"
    assert xl.read(X.xPutDupire(100, xl.fp12(spots, k), xl.fp12(times, k), xl.fp12(vols[:, :5], k), 0.25, xl.xstr("xdup2", k))) == xl.NA
    assert xl.read(X.xPutDupire(100, xl.fp12(spots, k), xl.fp12(times, k), xl.fp12(vols, k), 0.0, xl.xstr("xdup2", k))) == xl.NA
    assert len(xl.read(X.xParameters(xl.xstr("xdup", k)))) == 1081
    assert X.xRestartThreadPool(3.0) == 3.0


def test_multi_asset_puts(X, api):
    k = []
    assets = xl.xstrs(["a", "b", "c"], k)
    correl = np.full((3, 3), 0.5) + 0.5 * np.eye(3)
    args = (assets, xl.fp12([100, 90, 110], k), xl.fp12([0.2, 0.25, 0.15], k), xl.fp12([-0.05, 0.0, -0.1], k), 0.02,
            xl.fp12([0.0, 0.001, 0.002], k), xl.fp12([0.5, 1.5], k), xl.fp12(np.full((2, 3), 0.01), k), xl.fp12(correl, k), 0.25)
    assert xl.read(X.xPutDLM(*args, xl.xstr("xdlm", k))) == "xdlm"
    assert len(xl.read(X.xParameters(xl.xstr("xdlm", k)))) == api.num_params("xdlm")
    bad = list(args); bad[1] = xl.fp12([100, 90], k)
    assert xl.read(X.xPutDLM(*bad, xl.xstr("xdlm2", k))) == xl.NA
    assert xl.read(X.xPutAutocall(assets, xl.fp12([100, 90, 110], k), 3.0, 6.0, 1.0, 0.7, 0.1, 0.01, xl.xstr("xauto", k))) == "xauto"
    assert xl.read(X.xPutAutocall(assets, xl.fp12([100, 90], k), 3.0, 6.0, 1.0, 0.7, 0.1, 0.01, xl.xstr("xauto2", k))) == xl.NA
    assert xl.read(X.xPutBaskets(assets, xl.fp12([0.3, 0.3, 0.4], k), 2.0, xl.fp12([90, 100, 110], k), xl.xstr("xbask", k))) == "xbask"
    assert len(xl.read(X.xPayoffIds(xl.xstr("xbask", k)))) == 3
    assert xl.read(X.xPutMultiStats(assets, xl.fp12([1.0, 2.0, 0.0], k), xl.fp12([1.0, 2.5, 0.0], k), xl.xstr("xstats", k))) == "xstats"
    assert xl.read(X.xPutMultiStats(assets, xl.fp12([2.0, 1.0], k), xl.fp12([2.0, 1.0], k), xl.xstr("xstats2", k))) == xl.NA   # not increasing


def test_host_side_analytics(X, api):
    k = []
    assert X.xMerton(100, 0.15, 2.0, 110, 0.05, -0.15, 0.1) > 0
    spots, times = [50.0, 100.0, 200.0], [0.25, 3.0]
    got = xl.read(X.xDupireCalib(100, 0.15, 0.05, -0.15, 0.10, xl.fp12(spots, k), 5.0, xl.fp12(times, k), 1.0 / 12))
    cs, ct, lv = api.dupire_calib(spots, 5.0, times, 1.0 / 12, 100, 0.15, 0.05, -0.15, 0.10)
    assert got[0][0] == "" and got[0][1:] == list(ct) and [r[0] for r in got[1:]] == list(cs)
    assert np.array_equal(np.array([r[1:] for r in got[1:]]), lv)
    assert not X.xDupireCalib(100, 0.15, 0.05, -0.15, 0.10, xl.fp12(spots, k), 0.0, xl.fp12(times, k), 1.0 / 12)     # NULL, as the reference


@pytest.mark.gpu
def test_simulation_wrappers_match_the_flat_api(X, cf):
    k = []
    X = xl.bind(cf.lib)
    spots, times, vols = config3_surface()
    X.xPutDupire(100, xl.fp12(spots, k), xl.fp12(times, k), xl.fp12(vols, k), 0.25, xl.xstr("xdup", k))
    X.xPutBarrier(120, 150, 3.0, 1.0 / 52, 0.01, xl.xstr("call", k), xl.xstr("xuoc", k))
    n = 1 << 14
    labels = list(cf.payoff_labels("xuoc"))
    got = xl.read(X.xValue(xl.xstr("xdup", k), xl.xstr("xuoc", k), 1.0, 0.0, 0.0, float(n), 1.0))
    assert [r[0] for r in got] == labels and np.array_equal([r[1] for r in got], cf.value("xdup", "xuoc", n))
    assert xl.read(X.xValue(xl.xstr("xdup", k), xl.xstr("xuoc", k), 1.0, 0.0, 0.0, 0.0, 1.0)) == xl.NA            # no paths
    assert xl.read(X.xValue(xl.xstr("nobody", k), xl.xstr("xuoc", k), 1.0, 0.0, 0.0, float(n), 1.0)) == xl.NA
    timed = xl.read(X.xValueTime(xl.xstr("xdup", k), xl.xstr("xuoc", k), 1.0, 0.0, 0.0, float(n), 1.0))
    assert len(timed) == 3 and timed[0][0] == got[0][1] and timed[2][0] >= 0
    # one risk payoff / an aggregate of payoffs
    pv, v, risks = cf.aad_risk_one("xdup", "xuoc", n, risk_payoff=1)
    got = xl.read(X.xAADrisk(xl.xstr("xdup", k), xl.xstr("xuoc", k), xl.xstr(labels[1], k), 1.0, 0.0, 0.0, float(n), 1.0))
    assert got[0] == ["value", v] and len(got) == 1082 and np.array_equal([r[1] for r in got[1:]], risks)
    val, delta, vega = cf.dupire_aad_risk("xdup", "xuoc", [1.0, 0.5], 30, 36, n)
    got = xl.read(X.xAADriskAggregate(xl.xstr("xdup", k), xl.xstr("xuoc", k), xl.xstrs(labels + [""], k), xl.fp12([1.0, 0.5, 7.0], k),
                                      1.0, 0.0, 0.0, float(n), 1.0))
    assert got[0] == ["value", val] and got[1][1] == delta and np.array_equal(np.array([r[1] for r in got[2:]]).reshape(30, 36), vega)
    assert xl.read(X.xAADriskAggregate(xl.xstr("xdup", k), xl.xstr("xuoc", k), xl.xstrs(labels, k), xl.fp12([1.0], k),
                                       1.0, 0.0, 0.0, float(n), 1.0)) == xl.NA                                      # shapes differ
    # itemised risk: shown at once, or stored and displayed by payoff; bumps next to it
    X.xPutBlackScholes(100, 0.2, 0, 0.02, 0.0, xl.xstr("xbs", k))
    now = xl.read(X.xAADriskMulti(xl.xstr("xbs", k), xl.xstr("xuoc", k), 1.0, 0.0, 0.0, 4096.0, 1.0, 1.0, xl.xstr("", k)))
    assert now[0] == [""] + labels and now[1][0] == "value" and [r[0] for r in now[2:]] == list(cf.param_labels("xbs"))
    assert xl.read(X.xAADriskMulti(xl.xstr("xbs", k), xl.xstr("xuoc", k), 1.0, 0.0, 0.0, 4096.0, 1.0, 0.0, xl.xstr("rep", k))) == "rep"
    shown = xl.read(X.xDisplayRisk(xl.xstr("rep", k), xl.xstr(labels[1], k)))
    assert shown[0] == ["", labels[1]] and [r[1] for r in shown[1:]] == [r[2] for r in now[1:]]
    assert xl.read(X.xDisplayRisk(xl.xstr("rep", k), xl.xstr("no such payoff", k))) == xl.NA
    assert xl.read(X.xDisplayRisk(xl.xstr("no such report", k), xl.xstr(labels[1], k))) == xl.NA
    bumps = xl.read(X.xBumprisk(xl.xstr("xbs", k), xl.xstr("xuoc", k), 1.0, 0.0, 0.0, 4096.0, 1.0, 1.0, xl.xstr("", k)))
    assert np.allclose(np.array([r[1:] for r in bumps[2:]], dtype=float), np.array([r[1:] for r in now[2:]], dtype=float), rtol=1e-2, atol=1e-4)      # 1e-8 bumps through a smoothed barrier: a coarse cross-check
    # Sobol points straight from the device generator, with their antithetics
    pts = np.array(xl.read(X.xSobolPoints(6.0, 3.0, 1.0, 5.0)))
    seq = cf.rng_sequence(True, 3, 5, 3, False)
    assert np.array_equal(pts[0::2], seq) and np.array_equal(pts[1::2], 1 - seq)


@pytest.mark.gpu
def test_superbucket_wrapper(X, cf):
    k = []
    X = xl.bind(cf.lib)
    X.xPutEuropean(110, 2.0, 0.0, xl.xstr("xeur", k))
    strikes, mats = [90.0, 100.0, 110.0, 120.0], [1.0, 2.0]
    args = (100, 0.15, 0.05, -0.15, 0.10, xl.fp12(strikes, k), xl.fp12(mats, k), xl.fp12([50.0, 100.0, 200.0], k), 5.0,
            xl.fp12([0.25, 3.0], k), 1.0 / 12, 0.25, xl.xstr("xeur", k), xl.xstr("call 110.00 2.00", k), xl.fp12([1.0], k),
            1.0, 0.0, 0.0, float(1 << 14), 1.0, 0.0)
    got = xl.read(X.xDupireSuperbucket(*args))
    value, delta, vega = cf.dupire_superbucket(100, 0.25, "xeur", [1.0], [50.0, 100.0, 200.0], 5.0, [0.25, 3.0], 1.0 / 12,
                                               strikes, mats, 0.15, 0.05, -0.15, 0.10, 1 << 14)
    assert got[0][:2] == ["value", value] and got[1][:2] == ["delta", delta]
    assert got[2][:2] == ["vega", "mats"] and got[2][2:] == mats and [row[1] for row in got[4:]] == strikes
    assert np.array_equal(np.array([row[2:] for row in got[4:]], dtype=float), vega)
