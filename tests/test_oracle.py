"""CPU tests of the oracle: the numpy restatement (oracle/restate.py) against the known answers of
SURVEY.md section 8c and against the reference itself compiled with g++ (oracle/_ref)."""
import numpy as np
import pytest

from oracle import restate as R
from conftest import config3_surface, put_config3, rel_err


def test_direction_numbers_kat():
    d = R.sobol_direction_numbers()
    assert d.shape == (32, 1101)
    assert list(d[0, :4]) == [2147483648] * 4
    assert list(d[1, :4]) == [1073741824, 3221225472, 1073741824, 3221225472]
    assert d[20, 155] == 3977750528 and d[31, 1100] == 1996488755


def test_direction_numbers_vs_reference(ref):
    d = R.sobol_direction_numbers()
    for bit in range(32):
        for dim in list(range(0, 1101, 7)) + [1100]:
            assert ref.sobol_dirnum(bit, dim) == d[bit, dim]


def test_sobol_states_kat():
    s = R.sobol_states(2, 0, 3)
    assert s.tolist() == [[2147483648, 2147483648], [3221225472, 1073741824], [1073741824, 3221225472]]
    s = R.sobol_states(156, 123456, 1)
    assert s[0, 155] == 4105338880
    assert R.sobol_uniforms(156, 123456, 1)[0, 155] == 0.95584869384765769
    assert R.sobol_uniforms(1, 0, 1)[0, 0] == 0.50000000000000078      # ONEOVER2POW32 is not 2^-32
    assert (R.sobol_states(9, 0, 5000) == R.sobol_sequential(9, 5000)).all()


def test_sobol_vs_reference(ref):
    for dim, first, n in [(5, 0, 300), (156, 65536 - 10, 40), (1101, 12345, 3)]:
        u = ref.rng_draw(True, dim, first, n, False)
        assert (u == R.sobol_uniforms(dim, first, n)).all()


def test_mrg_kat():
    u = R.mrg32k3a_uniforms(12345, 12346, 3, 0, 3)
    assert u[0].tolist() == [0.12720739293357769, 0.87427024726015778, 0.58164476276890154]
    assert np.allclose(u[1], 1.0 - u[0], rtol=0, atol=0)
    assert u[2].tolist() == [0.08016046711089489, 0.2417626074251305, 0.90836507825645063]
    assert R.mrg32k3a_numerators(12345, 12346, 4, 0, 1)[0].tolist() == [546351566, 3754961938, 2498145113, 344286568]
    assert R.mrg32k3a_uniforms(12345, 12346, 12, 6400, 1)[0, 0] == 0.85964409839500966


def test_mrg_vs_reference(ref):
    for dim, first, n, s1, s2 in [(3, 0, 20, 12345, 12346), (12, 6400, 64, 12345, 12346), (120, 128, 10, 7, 11)]:
        u = ref.rng_draw(False, dim, first, n, False, seed1=s1, seed2=s2)
        assert (u == R.mrg32k3a_uniforms(s1, s2, dim, first, n)).all()


def test_inv_normal_kat():
    p = [1e-9, 0.001, 0.08, 0.3, 0.5, 0.50000000000000078, 0.92]
    want = [-5.9978070148919898, -3.0902323063275299, -1.4050715603096318, -0.52440051190665271, 0.0,
            1.9480414694550416e-15, 1.4050715603096322]
    got = R.inv_normal_cdf(p)
    assert np.allclose(got, want, rtol=2e-15, atol=1e-30)


def test_inv_normal_vs_reference(ref):
    p = np.concatenate([np.linspace(2.4e-10, 1 - 2.4e-10, 20001), [0.08, 0.92, 0.0799999, 0.9200001]])
    assert np.max(np.abs(R.inv_normal_cdf(p) - ref.inv_normal(p))) < 1e-14


def test_timeline_and_tables_vs_reference(ref):
    put_config3(ref)
    ptl = R.uoc_timeline(3.0, 1.0 / 52)
    assert len(ptl) == 157 and (np.array(ptl) == ref.product_timeline("uoc")).all()
    spots, times, vols = config3_surface()
    tab = R.DupireTables(100.0, spots, times, vols, 0.25, ptl)
    assert tab.n_steps == 156 and all(tab.common)
    # a coarse product timeline forces fillData to insert steps
    tab2 = R.DupireTables(100.0, spots, times, vols, 0.25, [1.0])
    assert tab2.n_steps == 4 and tab2.common == [False, False, False, False, True]


@pytest.mark.parametrize("rng", [("sobol",), ("mrg32k3a", 12345, 12346)])
def test_dupire_uoc_restatement_vs_reference(ref, rng):
    spots, times, vols = put_config3(ref)
    n = 1 << 12
    tab = R.DupireTables(100.0, spots, times, vols, 0.25, R.uoc_timeline(3.0, 1.0 / 52))
    w = [0.7, 0.3]
    out = R.dupire_uoc_run(tab, dict(strike=120.0, barrier=150.0, smooth=0.01), rng, 0, n, w)
    sobol = rng[0] == "sobol"
    pp = ref.simul_paths("dup", "uoc", n, sobol=sobol)
    assert np.max(np.abs(pp - out["payoffs"])) < 1e-9
    val, delta, vega = ref.dupire_aad_risk("dup", "uoc", w, 30, 36, n, sobol=sobol)
    d, v = tab.param_risks(out["spot_adj"], out["ybar"], n)
    assert abs(out["agg"].mean() / val - 1) < 1e-12
    assert abs(d / delta - 1) < 1e-10
    big = np.abs(vega) > 1e-6
    assert rel_err(v[big], vega[big]) < 1e-8 and np.max(np.abs(v - vega)) < 1e-12


def test_bs_restatement_vs_reference(ref):
    n = 1 << 12
    ref.put_bs(100, 0.15, False, 0.03, 0.01, "bs")
    ref.put_european(100, 1.0, 1.25, "eur")
    ref.put_barrier(100, 120, 1.0, 1.0 / 52, 0.01, False, "uoc1y")
    tb = R.BSTables(100, 0.15, 0.03, 0.01, [1.0], [1.25], [1.25], [True])
    o = R.bs_run(tb, "european", dict(strike=100), ("sobol",), 0, n, [1.0])
    pv, rv, risks = ref.aad_risk_one("bs", "eur", n)
    assert abs(o["payoffs"].mean() / pv[0] - 1) < 1e-13
    assert rel_err(tb.param_risks(o["table_adj"], n), risks) < 1e-10
    ptl = R.uoc_timeline(1.0, 1.0 / 52)
    tb2 = R.BSTables(100, 0.15, 0.03, 0.01, ptl, ptl, [None] * len(ptl), [False] * (len(ptl) - 1) + [True])
    o = R.bs_run(tb2, "uoc", dict(strike=100, barrier=120, smooth=0.01), ("mrg32k3a", 12345, 12346), 0, n, [1.0, 0.0])
    pv, rv, risks = ref.aad_risk_one("bs", "uoc1y", n, sobol=False)
    assert rel_err(o["payoffs"].mean(0), pv) < 1e-12
    assert rel_err(tb2.param_risks(o["table_adj"], n), risks) < 1e-9


def test_reference_reproduces_survey_pins(ref):
    """The g++ build of the reference gives the numbers pinned in SURVEY.md 8c / BASELINE.md (config 1)."""
    ref.put_bs(100, 0.15, False, 0, 0, "bs0")
    ref.put_european(100, 1.0, 1.0, "eur0")
    assert abs(ref.value("bs0", "eur0", 1 << 16)[0] / 5.9777943646922012 - 1) < 1e-14
    pv, rv, risks = ref.aad_risk_one("bs0", "eur0", 1 << 16)
    want = [0.5298707170844229, 39.775613482124662, 47.009277343749957, -52.987071708442201]
    assert rel_err(risks, want) < 1e-12


def test_the_reference_moves_with_the_compiler(ref):
    """Pins the noise floor of the superbucket chain (main.h:453-569 through ivs.h:119-138): the same reference sources
    built with and without fused multiply-adds disagree with each other at the 1e-4 level -- in the calibrated local
    vols already -- so 1e-8 agreement with "the reference" is not defined for this entry point; prices and AAD risks of
    the Monte-Carlo path itself move by ~1e-13 only."""
    from oracle import build_ref, refapi
    try:
        build_ref.build("/root/reference", verbose=False, variant="fma")
        fma = refapi.get("fma")
    except (FileNotFoundError, OSError):
        pytest.skip("second reference build not available")
    fma.start_pool(-1)
    for r in (ref, fma):
        r.put_barrier(120.0, 150.0, 1.0, 1.0 / 52, 0.01, False, "uoc_noise")
    args = dict(spot=100.0, max_dt=0.25, product="uoc_noise", notionals=[1.0, 0.0], incl_spots=[50.0, 100.0, 200.0], max_ds=10.0,
                incl_times=[0.25, 1.0], max_dt_vol=0.25, strikes=[80.0, 100.0, 120.0, 140.0], mats=[0.5, 1.0], vol=0.15,
                jmp_intens=0.05, jmp_avg=-0.15, jmp_std=0.10, n_path=1 << 11)
    va, da, ga = ref.dupire_superbucket(**args)
    vb, db, gb = fma.dupire_superbucket(**args)
    spread = np.max(np.abs(ga - gb)) / np.max(np.abs(ga))
    assert 1e-7 < spread < 1e-2 and 1e-7 < abs(va / vb - 1) < 1e-2
    la = ref.dupire_calib([50.0, 100.0, 200.0], 10.0, [0.25, 1.0], 0.25, 100.0, 0.15, 0.05, -0.15, 0.10)[2]
    lb = fma.dupire_calib([50.0, 100.0, 200.0], 10.0, [0.25, 1.0], 0.25, 100.0, 0.15, 0.05, -0.15, 0.10)[2]
    assert 1e-8 < np.max(np.abs(la - lb)) < 1e-3
    # the Monte-Carlo path on a fixed model does not have that problem
    spots = np.arange(55, 201, 5.0); times = np.arange(1, 37) / 12.0
    vols = 0.15 + 0.10 * np.log(spots[:, None] / 100.0) ** 2 + 0.02 * times[None, :]
    for r in (ref, fma):
        r.put_dupire(100.0, spots, times, vols, 0.25, "dup_noise")
    xa = ref.dupire_aad_risk("dup_noise", "uoc_noise", [1.0, 0.0], 30, 36, 1 << 11)
    xb = fma.dupire_aad_risk("dup_noise", "uoc_noise", [1.0, 0.0], 30, 36, 1 << 11)
    assert abs(xa[0] / xb[0] - 1) < 1e-11 and abs(xa[1] / xb[1] - 1) < 1e-9
