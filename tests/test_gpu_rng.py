"""GPU parity of the RNG kernels (called through the C ABI): integer streams bit-exact with the
oracle, uniforms bit-exact, Gaussians within 1e-14 (device log/div are not glibc's)."""
import json
import os

import numpy as np
import pytest

from oracle import restate as R

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_r1.json")))


@pytest.mark.parametrize("dim,first,n", [(1, 0, 1), (7, 0, 1000), (156, 123456, 700), (3, 255, 513), (2, 256, 256),
                                         (1101, 1 << 20, 300), (16, (1 << 32) - 600, 599), (5, 1, 0x10001)])
def test_sobol_states_bit_exact(eng, dim, first, n):
    assert (eng.sobol_states(dim, first, n) == R.sobol_states(dim, first, n)).all()


def test_sobol_limits(eng):
    from compfinance_b200.capi import CfError
    with pytest.raises(CfError, match="1101"):
        eng.sobol_states(1102, 0, 10)
    with pytest.raises(CfError, match="2\\^32"):
        eng.sobol_states(4, (1 << 32) - 5, 10)


@pytest.mark.parametrize("dim,first,n,s1,s2", [(3, 0, 10, 12345, 12346), (12, 6400, 100, 12345, 12346),
                                               (120, 100001, 50, 12345, 12346), (1, 7, 33, 1234, 1235),
                                               (156, 1 << 21, 40, 99, 100)])
def test_mrg_numerators_bit_exact(eng, dim, first, n, s1, s2):
    got = eng.mrg_numerators(eng.rng("mrg", s1, s2), dim, first, n)
    assert (got == R.mrg32k3a_numerators(s1, s2, dim, first, n)).all()


def test_mrg_top_of_the_index_range_and_limit(eng):
    """Path indices up to 2^32 - 1 (the reference's skipTo takes an unsigned): skip-ahead over 2^31 pairs x dim numbers."""
    from compfinance_b200.capi import CfError
    first, n = (1 << 32) - 301, 301
    got = eng.mrg_numerators(eng.rng("mrg", 12345, 12346), 12, first, n)
    assert (got == R.mrg32k3a_numerators(12345, 12346, 12, first, n)).all()
    with pytest.raises(CfError, match="2\\^32"):
        eng.mrg_numerators(eng.rng("mrg", 12345, 12346), 12, (1 << 32) - 5, 10)


def test_uniforms_bit_exact_and_golden(eng):
    u = eng.rng_draw(eng.rng("sobol"), 4, 1000, 4, False)
    assert (u == np.array(GOLD["sobol_uniforms_dim4_first1000_n4"])).all()
    u = eng.rng_draw(eng.rng("mrg"), 5, 6400, 4, False)
    assert (u == np.array(GOLD["mrg_uniforms_dim5_first6400_n4"])).all()
    u = eng.rng_draw(eng.rng("mrg"), 9, 3, 1001, False)          # odd start: antithetic partner first
    assert (u == R.mrg32k3a_uniforms(12345, 12346, 9, 3, 1001)).all()


def test_gaussians_vs_golden(eng):
    g = eng.rng_draw(eng.rng("sobol"), 156, 123456, 2, True)
    assert np.max(np.abs(g - np.array(GOLD["sobol_gauss_dim156_first123456_n2"]))) < 1e-14
    g = eng.rng_draw(eng.rng("mrg"), 12, 64, 3, True)
    assert np.max(np.abs(g - np.array(GOLD["mrg_gauss_dim12_first64_n3"]))) < 1e-14
    assert (g[1] == -g[0]).all()                                  # antithetic pair


def test_inv_normal(eng):
    p = np.concatenate([np.linspace(2.4e-10, 1 - 2.4e-10, 200001), [0.5, 0.50000000000000078, 0.08, 0.92, 0.0799999999]])
    got, want = eng.inv_normal(p), R.inv_normal_cdf(p)
    assert np.max(np.abs(got - want)) < 1e-14
    assert got[200001] == 0.0 and abs(got[200002] - 1.9480414694550416e-15) < 1e-29


def test_gaussians_vs_reference_live(eng, ref):
    for sobol, dim, first, n in [(True, 52, 4096, 300), (False, 52, 4096, 300)]:
        g = eng.rng_draw(eng.rng("sobol" if sobol else "mrg"), dim, first, n, True)
        assert np.max(np.abs(g - ref.rng_draw(sobol, dim, first, n, True))) < 1e-14


def test_sequential_rng_interface(cf):
    """RNG::init / skipTo / nextG of the host facade serves the device stream."""
    g = cf.rng_sequence(True, 6, 1000, 1500, True)
    assert np.max(np.abs(g - R.gaussians(("sobol",), 6, 1000, 1500))) < 1e-14
    u = cf.rng_sequence(False, 3, 0, 5, False, seed1=12345, seed2=12346)
    assert (u == R.mrg32k3a_uniforms(12345, 12346, 3, 0, 5)).all()


def test_lean_uniform_quotient_is_the_ieee_quotient_for_every_numerator(eng):
    """The path kernels compute u = z / (m1 + 1) with a reciprocal + residual correction instead of the IEEE division
    routine; all 2^32 numerators are compared on the device."""
    import ctypes as C
    bad = C.c_uint64(1)
    eng._chk(eng.lib.cf_selftest_mrg_uniform(C.byref(bad)))
    assert bad.value == 0
