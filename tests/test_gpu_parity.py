"""GPU parity of the path kernels through the drop-in API (cf_main.h entry points behind
libcf_host.so) and the low-level C ABI, against (a) committed golden vectors generated from the
reference and (b) the reference compiled with g++ run live on the same inputs.

Tolerances (BASELINE.json north_star): prices 1e-10 relative, AAD risks 1e-8 relative.  Vega entries
are compared relatively where |vega| > 1e-6 and absolutely (1e-12) elsewhere: the reference itself
only reproduces the small entries to ~1e-13 absolute from run to run (thread summation order)."""
import json
import os

import numpy as np
import pytest

from oracle import restate as R
from conftest import config3_surface, put_config3, rel_err

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_r1.json")))
PRICE_TOL, RISK_TOL = 1e-10, 1e-8


def check_vega(got, want):
    got, want = np.asarray(got), np.asarray(want)
    big = np.abs(want) > 1e-6
    assert rel_err(got[big], want[big]) < RISK_TOL
    assert np.max(np.abs(got - want)) < 1e-11


# ---- config 1: Black-Scholes European ---------------------------------------------------------
def test_config1_bs_european(cf):
    g = GOLD["config1"]
    cf.put_black_scholes(100, 0.15, False, 0.0, 0.0, "bs1")
    cf.put_european(100, 1.0, 1.0, "eur1")
    assert abs(cf.value("bs1", "eur1", g["n"])[0] / g["value"] - 1) < PRICE_TOL
    assert abs(cf.value("bs1", "eur1", g["n"])[0] / 5.9777943646922012 - 1) < PRICE_TOL     # BASELINE.md pin
    pv, rv, risks = cf.aad_risk_one("bs1", "eur1", g["n"])
    assert abs(rv / g["value"] - 1) < PRICE_TOL
    assert rel_err(risks, g["risks"]) < RISK_TOL
    assert rel_err(risks[:2], [0.5298707170844229, 39.775613482124662]) < RISK_TOL           # delta, vega pins


# ---- config 2: Black-Scholes barrier ----------------------------------------------------------
@pytest.mark.parametrize("name,sobol", [("config2_sobol", True), ("config2_mrg", False)])
def test_config2_bs_barrier(cf, name, sobol):
    g = GOLD[name]
    cf.put_black_scholes(100, 0.15, False, 0.03, 0.01, "bs2")
    cf.put_barrier(100, 120, 1.0, 1.0 / 52, 0.01, False, "uoc2")
    assert rel_err(cf.value("bs2", "uoc2", g["n"], sobol=sobol), g["values"]) < PRICE_TOL
    pv, rv, risks = cf.aad_risk_one("bs2", "uoc2", g["n"], sobol=sobol)
    assert rel_err(pv, g["values"]) < PRICE_TOL and abs(rv / g["risk_value"] - 1) < PRICE_TOL
    assert rel_err(risks, g["risks"]) < RISK_TOL


def test_config2_full_size(cf):
    g = GOLD["config2_full"]
    cf.put_black_scholes(100, 0.15, False, 0.0, 0.0, "bs1")
    cf.put_barrier(100, 120, 1.0, 1.0 / 52, 0.01, False, "uoc2")
    pv, rv, risks = cf.aad_risk_one("bs1", "uoc2", g["n"])
    assert rel_err(pv, g["values"]) < PRICE_TOL
    assert rel_err(pv, [2.1240722492984334, 5.9779041811635834]) < PRICE_TOL                 # BASELINE.md pins
    assert rel_err(risks[:2], g["risks"][:2]) < RISK_TOL
    assert np.max(np.abs(risks[2:] - np.array(g["risks"][2:]))) < 1e-8 * max(1.0, np.max(np.abs(g["risks"])))


BS_CASES = {
    # name: ((spot, vol, rate, div), product kind, product args)
    "late_settlement": ((100.0, 0.2, 0.03, 0.01), "european", (105.0, 1.0, 1.5)),
    "negative_rates": ((100.0, 0.2, -0.01, 0.0), "european", (95.0, 2.0, 2.0)),
    "dividends_above_rates": ((100.0, 0.25, 0.01, 0.06), "barrier", (100.0, 125.0, 1.5, 1.0 / 52, 0.01, False)),
    "high_vol_daily_barrier": ((100.0, 0.6, 0.02, 0.0), "barrier", (100.0, 160.0, 1.0, 1.0 / 252, 0.01, False)),
    "tiny_vol": ((100.0, 1e-4, 0.05, 0.0), "barrier", (100.0, 104.0, 1.0, 1.0 / 12, 0.01, False)),
    "put_barrier": ((100.0, 0.2, 0.03, 0.02), "barrier", (110.0, 130.0, 1.0, 1.0 / 52, 0.02, True)),
    "short_expiry": ((100.0, 0.2, 0.03, 0.0), "european", (100.0, 1.0 / 365, 1.0 / 365)),
    "ladder": ((100.0, 0.2, 0.03, 0.01), "europeans", ([0.5, 0.5, 1.0, 1.0, 2.0], [90.0, 110.0, 100.0, 100.0, 120.0])),
}


@pytest.mark.parametrize("case", sorted(BS_CASES))
@pytest.mark.parametrize("sobol", [True, False])
def test_odd_black_scholes_cases_vs_reference(cf, ref, case, sobol):
    """Black-Scholes corners: settlement after exercise (forward and discount factors), negative rates, dividends above
    rates, daily monitoring at high vol, almost no vol, a put barrier, a one-day option, a strike ladder with repeated
    dates and strikes: per-path payoffs and the four AAD risks against the reference."""
    (spot, vol, rate, div), kind, args = BS_CASES[case]
    for api in (cf, ref):
        (api.put_black_scholes if hasattr(api, "put_black_scholes") else api.put_bs)(spot, vol, False, rate, div, "bs_odd")
        getattr(api, "put_" + kind)(*args, "prd_odd")
    n = 2048 + 11
    got, want = cf.simul_paths("bs_odd", "prd_odd", n, sobol=sobol), ref.simul_paths("bs_odd", "prd_odd", n, sobol=sobol)
    assert np.max(np.abs(got - want)) < 1e-9
    last = got.shape[1] - 1
    pv, rv, risks = cf.aad_risk_one("bs_odd", "prd_odd", n, risk_payoff=last, sobol=sobol)
    pv_r, rv_r, risks_r = ref.aad_risk_one("bs_odd", "prd_odd", n, risk_payoff=last, sobol=sobol)
    assert abs(rv - rv_r) < PRICE_TOL * max(1.0, abs(rv_r))
    assert np.max(np.abs(risks - risks_r)) < RISK_TOL * max(1.0, np.max(np.abs(risks_r)))


# ---- config 3: Dupire barrier, risk to the whole local-vol surface ---------------------------------
@pytest.mark.parametrize("name,sobol", [("config3_sobol_16k", True), ("config3_mrg_16k", False)])
def test_config3_dupire_barrier_golden(cf, name, sobol):
    g = GOLD[name]
    put_config3(cf, "dup3", "uoc3")
    assert rel_err(cf.value("dup3", "uoc3", g["n"], sobol=sobol), g["values"]) < PRICE_TOL
    pp = cf.simul_paths("dup3", "uoc3", 64, sobol=sobol)
    assert np.max(np.abs(pp - np.array(g["first_paths"]))) < 1e-9
    val, delta, vega = cf.dupire_aad_risk("dup3", "uoc3", g["notionals"], 30, 36, g["n"], sobol=sobol)
    assert abs(val / g["value"] - 1) < PRICE_TOL and abs(delta / g["delta"] - 1) < RISK_TOL
    check_vega(vega, g["vega"])


def test_config3_full_size_north_star(cf):
    """2^20 Sobol paths x 156 steps, 1081 risks: the headline workload, against the reference's own numbers."""
    g = GOLD["config3_full"]
    put_config3(cf, "dup3", "uoc3")
    assert rel_err(cf.value("dup3", "uoc3", g["n"]), g["values"]) < PRICE_TOL
    val, delta, vega = cf.dupire_aad_risk("dup3", "uoc3", [1.0, 0.0], 30, 36, g["n"])
    assert abs(val / g["value"] - 1) < PRICE_TOL and abs(val / 0.96926107424976005 - 1) < PRICE_TOL
    assert abs(delta / g["delta"] - 1) < RISK_TOL and abs(delta / 0.020804057371458962 - 1) < RISK_TOL
    check_vega(vega, g["vega"])
    assert abs(vega.sum() / -5.4262407757753115 - 1) < RISK_TOL and abs(vega[13][11] / -0.024065608611340886 - 1) < RISK_TOL


def test_dupire_european_and_put_barrier(cf):
    put_config3(cf, "dup3", "uoc3")
    cf.put_european(110, 1.0, 1.0, "eur110")
    g = GOLD["dupire_european_16k"]
    val, delta, vega = cf.dupire_aad_risk("dup3", "eur110", [1.0], 30, 36, g["n"])
    assert abs(val / g["value"] - 1) < PRICE_TOL and abs(delta / g["delta"] - 1) < RISK_TOL
    check_vega(vega, g["vega"])
    cf.put_barrier(90, 130, 2.0, 1.0 / 12, 0.02, True, "uop")
    g = GOLD["dupire_uop_mrg_8k"]
    val, delta, vega = cf.dupire_aad_risk("dup3", "uop", [0.5, 0.5], 30, 36, g["n"], sobol=False)
    assert abs(val / g["value"] - 1) < PRICE_TOL and abs(delta / g["delta"] - 1) < RISK_TOL
    check_vega(vega, g["vega"])


# ---- live comparison with the compiled reference on odd shapes --------------------------------------
@pytest.mark.parametrize("n,sobol", [(1, True), (255, True), (257, False), (1000, True), (4097, False)])
def test_ragged_path_counts_vs_reference(cf, ref, n, sobol):
    put_config3(cf, "dup3", "uoc3")
    put_config3(ref, "dup3", "uoc3")
    assert np.max(np.abs(cf.simul_paths("dup3", "uoc3", n, sobol=sobol) - ref.simul_paths("dup3", "uoc3", n, sobol=sobol))) < 1e-9
    val, delta, vega = cf.dupire_aad_risk("dup3", "uoc3", [1.0, 0.5], 30, 36, n, sobol=sobol)
    rval, rdelta, rvega = ref.dupire_aad_risk("dup3", "uoc3", [1.0, 0.5], 30, 36, n, sobol=sobol)
    assert abs(val - rval) < 1e-10 * max(1.0, abs(rval)) and abs(delta - rdelta) < 1e-8 * max(abs(rdelta), 1e-3)
    assert np.max(np.abs(vega - rvega)) < 1e-8 * max(np.max(np.abs(rvega)), 1e-3)


@pytest.mark.parametrize("weeks", [1, 2, 3, 5, 6, 7, 9, 53, 157, 158, 159])
def test_step_counts_around_the_chunk_size_vs_reference(cf, ref, weeks):
    """The north-star kernels work on chunks of 4 steps (Gaussians, history sectors, reverse groups): step counts
    1, 2, 3 (less than a chunk), 5-7 and 157-159 (every remainder), with an odd path count, against the reference."""
    for api in (cf, ref):
        put_config3(api, "dupw", "uocw")
        api.put_barrier(105.0, 125.0, weeks / 52.0, 1.0 / 52, 0.01, False, f"uoc_w{weeks}")
    n = 2048 + 17
    assert cf.describe("dupw", f"uoc_w{weeks}")["n_steps"] == weeks
    want = ref.value("dupw", f"uoc_w{weeks}", n)
    assert rel_err(cf.value("dupw", f"uoc_w{weeks}", n), want) < PRICE_TOL
    val, delta, vega = cf.dupire_aad_risk("dupw", f"uoc_w{weeks}", [1.0, 0.5], 30, 36, n)
    rval, rdelta, rvega = ref.dupire_aad_risk("dupw", f"uoc_w{weeks}", [1.0, 0.5], 30, 36, n)
    assert abs(val / rval - 1) < PRICE_TOL and abs(delta / rdelta - 1) < RISK_TOL
    check_vega(vega, rvega)


@pytest.mark.parametrize("m,n_t", [(2, 1), (3, 2), (7, 36), (29, 3), (30, 1), (31, 5), (40, 36)])
def test_surface_shapes_vs_reference(cf, ref, m, n_t):
    """Local-vol surfaces from 2 x 1 knots up to more spot knots than the north-star kernels hold (30: beyond that the
    generic kernel takes over), a single time column (flat in time everywhere), against the reference."""
    spots = np.linspace(60.0, 180.0, m)
    times = np.linspace(0.1, 1.4, n_t) if n_t > 1 else np.array([0.5])
    vols = 0.16 + 0.08 * np.log(spots[:, None] / 100.0) ** 2 + 0.03 * times[None, :]
    for api in (cf, ref):
        api.put_dupire(100.0, spots, times, vols, 0.25, f"dup_{m}_{n_t}")
        api.put_barrier(105.0, 135.0, 1.0, 1.0 / 52, 0.01, False, "uoc_1y")
    n = 4096 + 33
    mid = f"dup_{m}_{n_t}"
    assert rel_err(cf.value(mid, "uoc_1y", n), ref.value(mid, "uoc_1y", n)) < PRICE_TOL
    val, delta, vega = cf.dupire_aad_risk(mid, "uoc_1y", [1.0, 0.25], m, n_t, n)
    rval, rdelta, rvega = ref.dupire_aad_risk(mid, "uoc_1y", [1.0, 0.25], m, n_t, n)
    assert abs(val / rval - 1) < PRICE_TOL and abs(delta / rdelta - 1) < RISK_TOL
    check_vega(vega, rvega)


@pytest.mark.parametrize("freq,maturity,steps", [(1.0 / 252, 3.0, 756), (1.0 / 52, 10.0, 520), (1.0 / 252, 1.0, 252)])
def test_long_schedules_vs_reference(cf, ref, freq, maturity, steps):
    """Daily monitoring over three years and weekly over ten: n_steps x n_knots tables beyond shared memory (the generic
    kernel reads table A from global memory and accumulates the table adjoints in its block's row), ten years also run
    past the last time knot of the surface.  Value and full AAD risk against the reference, both generators."""
    for api in (cf, ref):
        put_config3(api, "dupL", "uocL0")
        api.put_barrier(120.0, 150.0, maturity, freq, 0.01, False, "uocL")
    assert cf.describe("dupL", "uocL")["n_steps"] == steps
    n = 2048 + 5
    for sobol in ((True, False) if steps <= 1101 else (False,)):
        assert rel_err(cf.value("dupL", "uocL", n, sobol=sobol), ref.value("dupL", "uocL", n, sobol=sobol)) < PRICE_TOL
        val, delta, vega = cf.dupire_aad_risk("dupL", "uocL", [1.0, 0.5], 30, 36, n, sobol=sobol)
        rval, rdelta, rvega = ref.dupire_aad_risk("dupL", "uocL", [1.0, 0.5], 30, 36, n, sobol=sobol)
        assert abs(val / rval - 1) < PRICE_TOL and abs(delta / rdelta - 1) < RISK_TOL
        check_vega(vega, rvega)


ODD_CASES = {
    # name: (spots, times, max_dt, barrier args (strike, barrier, maturity, freq, smooth, put))
    "one_knot": (np.array([100.0]), np.array([1.0]), 0.25, (105.0, 130.0, 1.0, 1.0 / 52, 0.01, False)),
    "one_spot_knot_three_times": (np.array([90.0]), np.array([0.2, 0.6, 1.5]), 0.25, (105.0, 130.0, 1.0, 1.0 / 52, 0.01, False)),
    "fill_steps_between_events": (None, None, 0.01, (105.0, 130.0, 1.0, 1.0 / 52, 0.01, False)),
    "barrier_below_spot": (None, None, 0.25, (80.0, 95.0, 1.0, 1.0 / 52, 0.01, False)),
    "barrier_at_spot": (None, None, 0.25, (80.0, 100.0, 1.0, 1.0 / 52, 0.01, False)),
    "no_smoothing": (None, None, 0.25, (105.0, 130.0, 1.0, 1.0 / 52, 0.0, False)),
    "wide_smoothing": (None, None, 0.25, (105.0, 130.0, 1.0, 1.0 / 52, 0.25, False)),
    "strike_above_barrier": (None, None, 0.25, (140.0, 130.0, 1.0, 1.0 / 52, 0.01, False)),
    "put_monitored_monthly": (None, None, 0.25, (95.0, 125.0, 2.0, 1.0 / 12, 0.02, True)),
    "single_monitoring_date": (None, None, 0.25, (105.0, 130.0, 0.5, 1.0, 0.01, False)),
}


@pytest.mark.parametrize("case", sorted(ODD_CASES))
@pytest.mark.parametrize("sobol", [True, False])
def test_odd_products_and_surfaces_vs_reference(cf, ref, case, sobol):
    """Degenerate surfaces (a single knot), simulation steps between event dates, barriers at or below the spot
    (every path dies on the first sample), no smoothing and very wide smoothing, strike above the barrier, a put,
    one monitoring date: value, per-path payoffs and AAD risks against the reference."""
    spots, times, max_dt, bar = ODD_CASES[case]
    if spots is None:
        spots, times, _ = config3_surface()
    vols = 0.15 + 0.10 * np.log(spots[:, None] / 100.0) ** 2 + 0.02 * times[None, :]
    for api in (cf, ref):
        api.put_dupire(100.0, spots, times, vols, max_dt, "dup_odd")
        api.put_barrier(*bar, "uoc_odd")
    n = 2048 + 9
    got, want = cf.simul_paths("dup_odd", "uoc_odd", n, sobol=sobol), ref.simul_paths("dup_odd", "uoc_odd", n, sobol=sobol)
    assert np.max(np.abs(got - want)) < 1e-9
    val, delta, vega = cf.dupire_aad_risk("dup_odd", "uoc_odd", [1.0, 0.5], spots.size, times.size, n, sobol=sobol)
    rval, rdelta, rvega = ref.dupire_aad_risk("dup_odd", "uoc_odd", [1.0, 0.5], spots.size, times.size, n, sobol=sobol)
    assert abs(val - rval) < PRICE_TOL * max(1.0, abs(rval))
    assert abs(delta - rdelta) < RISK_TOL * max(abs(rdelta), 1e-3)
    check_vega(vega, rvega)


def test_live_path_regimes_vs_reference(cf, ref):
    """The reverse kernel sweeps only the paths with a non-zero payoff adjoint (the reference's tape skips
    zero-adjoint nodes, AADNode.h:76).  Regimes: many live paths per block (the European payoff carries weight:
    several passes of 4 paths per thread), none at all (deep out of the money: every risk is exactly zero),
    and everything in between on a shard that starts in the middle of the sequence."""
    put_config3(cf, "dup3", "uoc3")
    put_config3(ref, "dup3", "uoc3")
    n = 1 << 19
    val, delta, vega = cf.dupire_aad_risk("dup3", "uoc3", [0.3, 0.7], 30, 36, n)
    rval, rdelta, rvega = ref.dupire_aad_risk("dup3", "uoc3", [0.3, 0.7], 30, 36, n)
    assert abs(val / rval - 1) < PRICE_TOL and abs(delta / rdelta - 1) < RISK_TOL
    check_vega(vega, rvega)
    cf.put_barrier(400.0, 500.0, 3.0, 1.0 / 52, 0.01, False, "otm"); ref.put_barrier(400.0, 500.0, 3.0, 1.0 / 52, 0.01, False, "otm")
    val, delta, vega = cf.dupire_aad_risk("dup3", "otm", [1.0, 1.0], 30, 36, 1 << 16)
    rval, rdelta, rvega = ref.dupire_aad_risk("dup3", "otm", [1.0, 1.0], 30, 36, 1 << 16)
    assert val == 0.0 and rval == 0.0 and delta == 0.0 and rdelta == 0.0
    assert not np.any(vega) and not np.any(rvega)


@pytest.mark.parametrize("sobol", [True, False])
def test_contingent_bond_black_scholes(cf, ref, sobol):
    """ContingentBond (mcPrd.h:404-574) under Black-Scholes: libors, numeraires on every payment date, a smoothed
    digital on consecutive samples.  Value, per-path payoffs, AAD risks (spot, vol, rate, div) and bumps."""
    cf.put_black_scholes(100, 0.2, False, 0.03, 0.01, "bsc"); ref.put_bs(100, 0.2, False, 0.03, 0.01, "bsc")
    cf.put_contingent(0.02, 3.0, 0.25, 0.01, "cb"); ref.put_contingent(0.02, 3.0, 0.25, 0.01, "cb")
    assert np.array_equal(cf.product_timeline("cb"), ref.product_timeline("cb")) and len(cf.product_timeline("cb")) == 13
    n = 1 << 14
    assert np.max(np.abs(cf.simul_paths("bsc", "cb", n, sobol=sobol) - ref.simul_paths("bsc", "cb", n, sobol=sobol))) < 1e-9
    assert rel_err(cf.value("bsc", "cb", n, sobol=sobol), ref.value("bsc", "cb", n, sobol=sobol)) < PRICE_TOL
    pv, v, risks = cf.aad_risk_one("bsc", "cb", n, sobol=sobol)
    rpv, rv, rrisks = ref.aad_risk_one("bsc", "cb", n, sobol=sobol)
    assert abs(v / rv - 1) < PRICE_TOL and rel_err(risks, rrisks) < RISK_TOL
    # no smoothing: the digital has no fuzzy zone left, only the libor / numeraire chains carry risk
    cf.put_contingent(0.05, 1.0, 0.5, 0.0, "cb0"); ref.put_contingent(0.05, 1.0, 0.5, 0.0, "cb0")
    pv, v, risks = cf.aad_risk_one("bsc", "cb0", 4097, sobol=sobol)
    rpv, rv, rrisks = ref.aad_risk_one("bsc", "cb0", 4097, sobol=sobol)
    assert abs(v / rv - 1) < PRICE_TOL and np.max(np.abs(np.asarray(risks) - np.asarray(rrisks))) < 1e-8 * max(1.0, np.max(np.abs(rrisks)))


def test_per_path_aad_results_vs_reference(cf, ref):
    """mcSimulAAD's per-path outputs (payoffs, aggregated) through the mirrored free function."""
    cf.put_black_scholes(100, 0.2, False, 0.02, 0.0, "bsx"); ref.put_bs(100, 0.2, False, 0.02, 0.0, "bsx")
    cf.put_barrier(95, 125, 0.5, 1.0 / 52, 0.01, False, "uocx"); ref.put_barrier(95, 125, 0.5, 1.0 / 52, 0.01, False, "uocx")
    n = 3000
    pays, agg, risks = cf.simul_aad_paths("bsx", "uocx", n, risk_payoff=1)
    rp = ref.simul_paths("bsx", "uocx", n)
    assert np.max(np.abs(pays - rp)) < 1e-9 and np.max(np.abs(agg - rp[:, 1])) < 1e-9
    pv, rv, rrisks = ref.aad_risk_one("bsx", "uocx", n, risk_payoff=1)
    assert rel_err(risks, rrisks) < RISK_TOL


def test_deep_out_of_range_spots_flat_extrapolation(cf, ref):
    """Spot far outside the surface: every path sits in the flat-extrapolation region of interp()."""
    spots, times, vols = config3_surface()
    for api in (cf, ref):
        api.put_dupire(20.0, spots, times, vols, 0.25, "dlow")
        api.put_dupire(400.0, spots, times, vols, 0.25, "dhigh")
        api.put_european(20.0, 1.0, 1.0, "e20")
        api.put_european(380.0, 1.0, 1.0, "e380")
    for m, p in [("dlow", "e20"), ("dhigh", "e380")]:
        val, delta, vega = cf.dupire_aad_risk(m, p, [1.0], 30, 36, 2048)
        rval, rdelta, rvega = ref.dupire_aad_risk(m, p, [1.0], 30, 36, 2048)
        assert abs(val / rval - 1) < PRICE_TOL and abs(delta / rdelta - 1) < RISK_TOL
        check_vega(vega, rvega)


# ---- low-level C ABI: shards, fallback kernel, determinism --------------------------------------------
def _config3_lowlevel(eng, time_map=True):
    spots, times, vols = config3_surface()
    ptl = R.uoc_timeline(3.0, 1.0 / 52)
    tab = R.DupireTables(100.0, spots, times, vols, 0.25, ptl)
    mdl = eng.dupire_model(100.0, tab.log_spots, tab.interp_vols, tab.common, len(ptl),
                           time_map=tab.time_map() if time_map else None)
    prd = eng.uoc(120.0, 150.0, float(np.exp(np.log(100.0)) * 0.01), len(ptl))
    return tab, mdl, prd


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("rng,first", [("sobol", (1 << 32) - 1 - 3000), ("mrg", (1 << 32) - 3000)])
def test_top_of_the_path_index_range_vs_oracle(eng, rng, first, fast):
    """Maximum sizes: the last 3000 paths the 32-bit path index of the reference can address (Sobol point 2^32 - 1,
    mrg32k3a pair 2^31 - 1), value + AAD through the C ABI against the numpy restatement, north-star kernels and
    generic kernel; the padding lanes of the last batch lie past the end of the sequence.  One path more is refused."""
    from compfinance_b200.capi import CfError
    tab, mdl, prd = _config3_lowlevel(eng, time_map=fast)
    n, w = 3000, [1.0, 0.5]
    r = ("sobol",) if rng == "sobol" else ("mrg", 12345, 12346)
    o = R.dupire_uoc_run(tab, dict(strike=120.0, barrier=150.0, smooth=0.01), r, first, n, w)
    g = eng.run_aad(mdl, prd, eng.rng(rng), first, n, w, per_path=True)
    assert np.max(np.abs(g["payoffs"] - o["payoffs"])) < 1e-9
    assert abs(g["agg_sum"] / o["agg"].sum() - 1) < PRICE_TOL
    assert abs(g["table_adj"][0] / o["spot_adj"] - 1) < RISK_TOL
    if fast:      # adjoints of the local vols (init() folded into the sweep)
        _, v_o = tab.param_risks(o["spot_adj"], o["ybar"], n)
        check_vega(g["table_adj"][1:].reshape(30, 36) / n, v_o)
    else:         # adjoints of the interpolated table
        check_vega(g["table_adj"][1:].reshape(tab.n_steps, -1) / n, o["ybar"] / n)
    with pytest.raises(CfError, match="2\\^32"):
        eng.run_aad(mdl, prd, eng.rng(rng), first + 1, n, w)


@pytest.mark.parametrize("rng,first", [("sobol", 0), ("sobol", 777), ("sobol", (1 << 32) - 1 - 2500),
                                       ("mrg", 0), ("mrg", 12345), ("mrg", (1 << 32) - 2500)])
def test_black_scholes_shards_through_the_c_abi_vs_oracle(eng, rng, first):
    """Black-Scholes x UOC with rates and dividends through the C ABI (first_path / n_paths as a rank of a multi-GPU
    run passes them): shards from the start, from an odd path (mrg32k3a: the antithetic half of a pair) and at the top
    of the index range; per-path payoffs, aggregate and every table adjoint against the numpy restatement."""
    ptl = R.uoc_timeline(1.0, 1.0 / 52)
    tb = R.BSTables(100, 0.15, 0.03, 0.01, ptl, ptl, [None] * len(ptl), [False] * (len(ptl) - 1) + [True])
    mdl = eng.bs_model(100.0, tb.drifts, tb.stds, tb.is_event, tb.numeraires, tb.fwd_factors, tb.discounts)
    smooth = float(100.0 * tb.fwd_factors[0] * 0.01)
    prd = eng.uoc(100.0, 120.0, smooth, len(ptl))
    n, w = 2500, [1.0, 0.25]
    r = ("sobol",) if rng == "sobol" else ("mrg", 12345, 12346)
    o = R.bs_run(tb, "uoc", dict(strike=100, barrier=120, smooth=0.01), r, first, n, w)
    g = eng.run_aad(mdl, prd, eng.rng(rng), first, n, w, per_path=True)
    assert np.max(np.abs(g["payoffs"] - o["payoffs"])) < 1e-10
    assert abs(g["agg_sum"] / o["agg"].sum() - 1) < PRICE_TOL
    D, E = tb.n_steps, len(ptl)
    got, want = g["table_adj"], o["table_adj"]
    assert got.size == 1 + 2 * D + 4 * E and not np.any(got[1 + 2 * D + 3 * E:])       # no libors in a barrier's samples
    scale = np.max(np.abs(want))
    assert np.max(np.abs(got[:want.size] - want)) < RISK_TOL * scale
    risks = tb.param_risks(got[:want.size], n)
    assert rel_err(risks, tb.param_risks(want, n)) < RISK_TOL


def test_shards_add_up(eng):
    """Disjoint skip-ahead blocks (the multi-GPU partition) sum to the single run."""
    tab, mdl, prd = _config3_lowlevel(eng)
    n, w = 1 << 15, [1.0, 0.25]
    for rng in ("sobol", "mrg"):
        whole = eng.run_aad(mdl, prd, eng.rng(rng), 0, n, w)
        parts = [eng.run_aad(mdl, prd, eng.rng(rng), k * n // 4, n // 4, w) for k in range(4)]
        assert abs(sum(p["agg_sum"] for p in parts) / whole["agg_sum"] - 1) < 1e-13
        tot = sum(p["table_adj"] for p in parts)
        assert np.max(np.abs(tot - whole["table_adj"])) < 1e-9 * np.max(np.abs(whole["table_adj"]))


def test_runs_longer_than_one_launch_and_mid_size_shards(eng):
    """A run of more than 2^21 paths is cut into launches that accumulate into the same per-warp tables; a mid-size
    shard takes the 1-path-per-thread forward shape with every SM busy.  Both must add up to the same sums."""
    tab, mdl, prd = _config3_lowlevel(eng)
    w = [1.0, 0.5]
    n = (1 << 21) + 4097
    whole = eng.run_aad(mdl, prd, eng.rng("sobol"), 0, n, w)
    cuts = [0, 100_000, 100_000 + (1 << 20), n]
    parts = [eng.run_aad(mdl, prd, eng.rng("sobol"), a, b - a, w) for a, b in zip(cuts[:-1], cuts[1:])]
    assert abs(sum(p["agg_sum"] for p in parts) / whole["agg_sum"] - 1) < 1e-13
    assert np.max(np.abs(sum(p["payoff_sums"] for p in parts) / whole["payoff_sums"] - 1)) < 1e-13
    tot = sum(p["table_adj"] for p in parts)
    assert np.max(np.abs(tot - whole["table_adj"])) < 1e-11 * np.max(np.abs(whole["table_adj"]))


def test_generic_kernel_matches_fast_kernel(eng):
    tab, mdl_fast, prd = _config3_lowlevel(eng, True)
    _, mdl_gen, _ = _config3_lowlevel(eng, False)
    n, w = 1 << 13, [0.7, 0.3]
    a = eng.run_aad(mdl_fast, prd, eng.rng("sobol"), 0, n, w)
    b = eng.run_aad(mdl_gen, prd, eng.rng("sobol"), 0, n, w)
    assert a["agg_sum"] == pytest.approx(b["agg_sum"], rel=1e-14)
    _, v = tab.param_risks(b["table_adj"][0], b["table_adj"][1:].reshape(tab.n_steps, -1), 1)
    assert np.max(np.abs(a["table_adj"][1:].reshape(30, 36) - v)) < 1e-10 * np.max(np.abs(v))


def test_bitwise_deterministic(eng):
    tab, mdl, prd = _config3_lowlevel(eng)
    a = eng.run_aad(mdl, prd, eng.rng("sobol"), 0, 1 << 16, [1.0, 0.0])
    b = eng.run_aad(mdl, prd, eng.rng("sobol"), 0, 1 << 16, [1.0, 0.0])
    assert (a["table_adj"] == b["table_adj"]).all() and a["agg_sum"] == b["agg_sum"] and (a["payoff_sums"] == b["payoff_sums"]).all()


def test_aad_matches_bumps(cf):
    """Independent check of the adjoint kernels: finite differences by re-running value() (bumpRisk, main.h:316)."""
    cf.put_black_scholes(100, 0.15, False, 0.03, 0.01, "bs2")
    cf.put_barrier(100, 120, 1.0, 1.0 / 52, 0.05, False, "uocw")
    n = 1 << 15
    pv, rv, risks = cf.aad_risk_one("bs2", "uocw", n)
    values, bumps = cf.bump_risk("bs2", "uocw", n)
    # the spot is skipped: the smoothing half-width double(S0 * smooth) moves with a spot bump but is, by
    # design, not differentiated on the tape (mcPrd.h:247), so the reference's AAD delta differs from its bump delta
    assert np.max(np.abs(bumps[1:, 0] - risks[1:])) < 1e-5 * np.max(np.abs(risks))


def test_invalid_descriptors(eng):
    from compfinance_b200.capi import CfError
    tab, mdl, prd = _config3_lowlevel(eng)
    prd_bad = eng.uoc(120.0, 150.0, 1.0, 10)
    with pytest.raises(CfError, match="event"):
        eng.run_value(mdl, prd_bad, eng.rng("sobol"), 0, 100)
    with pytest.raises(CfError, match="n_paths"):
        eng.run_value(mdl, prd, eng.rng("sobol"), 0, 0)


def test_black_scholes_barrier_fast_path_large_and_ragged_runs(cf, ref):
    """The Black-Scholes barrier path of its own (shared forward kernel with the log-normal step, lane-parallel reverse
    over contiguous live ranges): a run longer than one launch (2^21 + 4097 paths), a ragged small one and a mid-size
    one (one path per thread), against the reference with both generators; puts, a discounted late settlement."""
    for api in (cf, ref):
        (api.put_black_scholes if api is cf else api.put_bs)(100.0, 0.22, False, 0.03, 0.01, "bsfp")
        api.put_barrier(105.0, 135.0, 1.0, 1.0 / 52, 0.5, False, "uocfp")
        api.put_barrier(95.0, 140.0, 1.0, 1.0 / 52, 1.0, True, "uopfp")
    for prd, n, sobol in [("uocfp", (1 << 21) + 4097, True), ("uocfp", 4097, False), ("uopfp", 150_001, False), ("uopfp", 257, True)]:
        pv, rv, risks = cf.aad_risk_one("bsfp", prd, n, risk_payoff=0, sobol=sobol)
        pv_r, rv_r, risks_r = ref.aad_risk_one("bsfp", prd, n, risk_payoff=0, sobol=sobol)
        assert rel_err(pv, pv_r) < PRICE_TOL and abs(rv / rv_r - 1) < PRICE_TOL
        assert rel_err(risks, risks_r) < RISK_TOL
        assert rel_err(cf.value("bsfp", prd, n, sobol=sobol), ref.value("bsfp", prd, n, sobol=sobol)) < PRICE_TOL


@pytest.mark.parametrize("sobol", [False, True])
def test_arbitrary_shard_boundaries_add_up_black_scholes_and_europeans(cf, sobol):
    """cf_run_value / cf_run_aad over ranges with odd boundaries (antithetic pairs split, Sobol windows cut, a single path,
    several batches per block) add up to the whole run: the Black-Scholes barrier path of its own and the generic
    kernel with contiguous batches per block (incremental mrg32k3a jumps) on a ladder of Europeans."""
    spots, times, vols = config3_surface()
    cf.put_black_scholes(100.0, 0.2, False, 0.03, 0.01, "bsr")
    cf.put_barrier(100.0, 130.0, 1.0, 1.0 / 52, 0.5, False, "uocr")
    cf.put_dupire(100.0, spots, times, vols, 0.25, "dupr")
    cf.put_europeans([0.5, 0.5, 1.0, 1.0, 1.0], [95.0, 105.0, 90.0, 100.0, 110.0], "eursr")
    for model, product, w, n in [("bsr", "uocr", [1.0, 0.5], 70_001), ("dupr", "eursr", [1.0, -0.5, 0.25, 2.0, 1.0], 150_001)]:
        nadj = cf.describe(model, product, aad=True)["adjoint_size"]
        whole = cf.run_range(model, product, 0, n, w, sobol=sobol, n_adjoints=nadj)
        cuts = [0, 1, 1001, 1002, 33_333, n - 1, n]
        parts = [cf.run_range(model, product, a, b - a, w, sobol=sobol, n_adjoints=nadj) for a, b in zip(cuts[:-1], cuts[1:])]
        assert np.max(np.abs(sum(p[0] for p in parts) / whole[0] - 1)) < 1e-12
        assert abs(sum(p[1] for p in parts) / whole[1] - 1) < 1e-12
        tot = sum(p[2] for p in parts)
        assert np.max(np.abs(tot - whole[2])) < 1e-10 * np.max(np.abs(whole[2]))
        vals = sum(cf.run_range(model, product, a, b - a, sobol=sobol)[0] for a, b in zip(cuts[:-1], cuts[1:]))
        assert np.max(np.abs(vals / whole[0] - 1)) < 1e-12
