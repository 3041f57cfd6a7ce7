"""ctypes binding of oracle/_ref/libcfref.so (the reference itself, compiled by build_ref.py).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (compfinance_b200/) never imports this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_ref", "libcfref.so")

_dp = C.POINTER(C.c_double)


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


class RefLib:
    """Thin, stateful wrapper: the reference keeps models/products in global stores (store.h)."""

    def __init__(self, path=SO_PATH):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run python oracle/build_ref.py")
        self.lib = C.CDLL(path)
        self.lib.ref_last_error.restype = C.c_char_p
        self.lib.ref_sobol_dirnum.restype = C.c_uint
        self.lib.ref_set_system_time.argtypes = [C.c_double]

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())

    # -- pool
    def start_pool(self, n=-1):
        self._chk(self.lib.ref_start_pool(C.c_int(n)))
        return self.lib.ref_pool_threads()

    def stop_pool(self):
        self._chk(self.lib.ref_stop_pool())

    def pool_threads(self):
        return self.lib.ref_pool_threads()

    # -- rng
    def rng_draw(self, sobol, dim, first, n, gaussian, seed1=12345, seed2=12346):
        out = np.empty((n, dim), dtype=np.float64)
        self._chk(self.lib.ref_rng_draw(C.c_int(int(sobol)), C.c_int(seed1), C.c_int(seed2), C.c_int(dim),
                                        C.c_uint(first), C.c_int(n), C.c_int(int(gaussian)),
                                        out.ctypes.data_as(_dp)))
        return out

    def inv_normal(self, p):
        p, pp = _d(p)
        out = np.empty_like(p)
        self.lib.ref_inv_normal(pp, out.ctypes.data_as(_dp), C.c_int(p.size))
        return out

    def sobol_dirnum(self, bit, dim):
        return int(self.lib.ref_sobol_dirnum(C.c_int(bit), C.c_int(dim)))

    # -- store
    def put_bs(self, spot, vol, spot_measure, rate, div, id_):
        self._chk(self.lib.ref_put_bs(C.c_double(spot), C.c_double(vol), C.c_int(int(spot_measure)),
                                      C.c_double(rate), C.c_double(div), id_.encode()))

    def put_dupire(self, spot, spots, times, vols, max_dt, id_):
        spots, ps = _d(spots)
        times, pt = _d(times)
        vols, pv = _d(vols)
        assert vols.shape == (spots.size, times.size)
        self._chk(self.lib.ref_put_dupire(C.c_double(spot), ps, C.c_int(spots.size), pt, C.c_int(times.size),
                                          pv, C.c_double(max_dt), id_.encode()))

    def put_displaced(self, spots, atms, skews, disc_rate, repo_spreads, div_dates, divs, correl, lam, id_):
        spots, p1 = _d(spots)
        atms, p2 = _d(atms)
        skews, p3 = _d(skews)
        repo_spreads, p4 = _d(repo_spreads)
        div_dates, p5 = _d(div_dates)
        divs, p6 = _d(divs)
        correl, p7 = _d(correl)
        n = spots.size
        self._chk(self.lib.ref_put_displaced(C.c_int(n), p1, p2, p3, C.c_double(disc_rate), p4, p5,
                                             C.c_int(div_dates.size), p6, p7, C.c_double(lam), id_.encode()))

    def put_european(self, strike, exercise, settlement, id_):
        self._chk(self.lib.ref_put_european(C.c_double(strike), C.c_double(exercise), C.c_double(settlement),
                                            id_.encode()))

    def put_contingent(self, coupon, maturity, pay_freq, smooth, id_):
        self._chk(self.lib.ref_put_contingent(C.c_double(coupon), C.c_double(maturity), C.c_double(pay_freq),
                                              C.c_double(smooth), id_.encode()))

    def put_barrier(self, strike, barrier, maturity, freq, smooth, call_put, id_):
        self._chk(self.lib.ref_put_barrier(C.c_double(strike), C.c_double(barrier), C.c_double(maturity),
                                           C.c_double(freq), C.c_double(smooth), C.c_int(int(call_put)),
                                           id_.encode()))

    def put_europeans(self, maturities, strikes, id_):
        m, pm = _d(maturities)
        k, pk = _d(strikes)
        assert m.size == k.size
        self._chk(self.lib.ref_put_europeans(pm, pk, C.c_int(m.size), id_.encode()))

    def put_multistats(self, n_assets, fix_dates, fwd_dates, id_):
        a, pa = _d(fix_dates)
        b, pb = _d(fwd_dates)
        self._chk(self.lib.ref_put_multistats(C.c_int(n_assets), pa, pb, C.c_int(a.size), id_.encode()))

    def put_baskets(self, weights, maturity, strikes, id_):
        w, pw = _d(weights)
        k, pk = _d(strikes)
        self._chk(self.lib.ref_put_baskets(C.c_int(w.size), pw, C.c_double(maturity), pk, C.c_int(k.size),
                                           id_.encode()))

    def put_autocall(self, refs, maturity, periods, ko, strike, cpn, smooth, id_):
        r, pr = _d(refs)
        self._chk(self.lib.ref_put_autocall(C.c_int(r.size), pr, C.c_double(maturity), C.c_int(periods),
                                            C.c_double(ko), C.c_double(strike), C.c_double(cpn),
                                            C.c_double(smooth), id_.encode()))

    def num_payoffs(self, product):
        return self.lib.ref_num_payoffs(product.encode())

    def num_params(self, model):
        return self.lib.ref_num_params(model.encode())

    def product_timeline(self, product):
        n = self.lib.ref_product_timeline(product.encode(), None, C.c_int(0))
        out = np.empty(n)
        self.lib.ref_product_timeline(product.encode(), out.ctypes.data_as(_dp), C.c_int(n))
        return out

    def labels(self, id_, params=False):
        n = self.lib.ref_labels(id_.encode(), C.c_int(int(params)), None, C.c_int(0))
        buf = C.create_string_buffer(n)
        self.lib.ref_labels(id_.encode(), C.c_int(int(params)), buf, C.c_int(n))
        return buf.value.decode().split("\n")[:-1]

    # -- entry points
    def value(self, model, product, n_path, sobol=True, parallel=True, seed1=12345, seed2=12346):
        out = np.empty(self.num_payoffs(product))
        self._chk(self.lib.ref_value(model.encode(), product.encode(), C.c_int(int(parallel)),
                                     C.c_int(int(sobol)), C.c_int(n_path), C.c_int(seed1), C.c_int(seed2),
                                     out.ctypes.data_as(_dp)))
        return out

    def simul_paths(self, model, product, n_path, sobol=True, parallel=True, seed1=12345, seed2=12346):
        out = np.empty((n_path, self.num_payoffs(product)))
        self._chk(self.lib.ref_simul_paths(model.encode(), product.encode(), C.c_int(int(parallel)),
                                           C.c_int(int(sobol)), C.c_int(n_path), C.c_int(seed1),
                                           C.c_int(seed2), out.ctypes.data_as(_dp)))
        return out

    def aad_risk_one(self, model, product, n_path, risk_payoff=-1, sobol=True, parallel=True, seed1=12345,
                     seed2=12346):
        pv = np.empty(self.num_payoffs(product))
        risks = np.empty(self.num_params(model))
        rv = C.c_double()
        self._chk(self.lib.ref_aad_risk_one(model.encode(), product.encode(), C.c_int(risk_payoff),
                                            C.c_int(int(parallel)), C.c_int(int(sobol)), C.c_int(n_path),
                                            C.c_int(seed1), C.c_int(seed2), pv.ctypes.data_as(_dp),
                                            C.byref(rv), risks.ctypes.data_as(_dp)))
        return pv, rv.value, risks

    def aad_risk_aggregate(self, model, product, notionals, n_path, sobol=True, parallel=True, seed1=12345,
                           seed2=12346):
        nots, pn = _d(notionals)
        pv = np.empty(self.num_payoffs(product))
        risks = np.empty(self.num_params(model))
        rv = C.c_double()
        self._chk(self.lib.ref_aad_risk_aggregate(model.encode(), product.encode(), pn, C.c_int(int(parallel)),
                                                  C.c_int(int(sobol)), C.c_int(n_path), C.c_int(seed1),
                                                  C.c_int(seed2), pv.ctypes.data_as(_dp), C.byref(rv),
                                                  risks.ctypes.data_as(_dp)))
        return pv, rv.value, risks

    def aad_risk_multi(self, model, product, n_path, sobol=True, parallel=True, seed1=12345, seed2=12346):
        npay, npar = self.num_payoffs(product), self.num_params(model)
        values = np.empty(npay)
        risks = np.empty((npar, npay))
        self._chk(self.lib.ref_aad_risk_multi(model.encode(), product.encode(), C.c_int(int(parallel)),
                                              C.c_int(int(sobol)), C.c_int(n_path), C.c_int(seed1),
                                              C.c_int(seed2), values.ctypes.data_as(_dp),
                                              risks.ctypes.data_as(_dp)))
        return values, risks

    def bump_risk(self, model, product, n_path, sobol=True, parallel=True, seed1=12345, seed2=12346):
        npay, npar = self.num_payoffs(product), self.num_params(model)
        values = np.empty(npay)
        risks = np.empty((npar, npay))
        self._chk(self.lib.ref_bump_risk(model.encode(), product.encode(), C.c_int(int(parallel)),
                                         C.c_int(int(sobol)), C.c_int(n_path), C.c_int(seed1), C.c_int(seed2),
                                         values.ctypes.data_as(_dp), risks.ctypes.data_as(_dp)))
        return values, risks

    def dupire_aad_risk(self, model, product, notionals, n_spots, n_times, n_path, sobol=True, parallel=True,
                        seed1=12345, seed2=12346):
        nots, pn = _d(notionals)
        vega = np.empty((n_spots, n_times))
        v, d = C.c_double(), C.c_double()
        self._chk(self.lib.ref_dupire_aad_risk(model.encode(), product.encode(), pn, C.c_int(int(parallel)),
                                               C.c_int(int(sobol)), C.c_int(n_path), C.c_int(seed1),
                                               C.c_int(seed2), C.byref(v), C.byref(d),
                                               vega.ctypes.data_as(_dp)))
        return v.value, d.value, vega

    def dupire_calib(self, incl_spots, max_ds, incl_times, max_dt, spot, vol, jmp_intens=0.0, jmp_avg=0.0,
                     jmp_std=0.0, cap=1 << 20):
        a, pa = _d(incl_spots)
        b, pb = _d(incl_times)
        spots, times, lv = np.empty(cap), np.empty(cap), np.empty(cap)
        ns, nt = C.c_int(), C.c_int()
        self._chk(self.lib.ref_dupire_calib(pa, C.c_int(a.size), C.c_double(max_ds), pb, C.c_int(b.size),
                                            C.c_double(max_dt), C.c_double(spot), C.c_double(vol),
                                            C.c_double(jmp_intens), C.c_double(jmp_avg), C.c_double(jmp_std),
                                            C.byref(ns), C.byref(nt), spots.ctypes.data_as(_dp),
                                            times.ctypes.data_as(_dp), lv.ctypes.data_as(_dp), C.c_int(cap)))
        ns, nt = ns.value, nt.value
        return spots[:ns].copy(), times[:nt].copy(), lv[:ns * nt].reshape(ns, nt).copy()

    def dupire_superbucket(self, spot, max_dt, product, notionals, incl_spots, max_ds, incl_times, max_dt_vol,
                           strikes, mats, vol, jmp_intens, jmp_avg, jmp_std, n_path, sobol=True, parallel=True,
                           seed1=12345, seed2=12346, bump=False):
        """dupireSuperbucket (main.h:453); bump=True: the reference's finite-difference driver dupireSuperbucketBump (main.h:575)."""
        nots, pn = _d(notionals)
        a, pa = _d(incl_spots)
        b, pb = _d(incl_times)
        k, pk = _d(strikes)
        m, pm = _d(mats)
        vega = np.empty((k.size, m.size))
        v, d = C.c_double(), C.c_double()
        fn = self.lib.ref_dupire_superbucket_bump if bump else self.lib.ref_dupire_superbucket
        self._chk(fn(
            C.c_double(spot), C.c_double(max_dt), product.encode(), pn, pa, C.c_int(a.size), C.c_double(max_ds),
            pb, C.c_int(b.size), C.c_double(max_dt_vol), pk, C.c_int(k.size), pm, C.c_int(m.size),
            C.c_double(vol), C.c_double(jmp_intens), C.c_double(jmp_avg), C.c_double(jmp_std),
            C.c_int(int(parallel)), C.c_int(int(sobol)), C.c_int(n_path), C.c_int(seed1), C.c_int(seed2),
            C.byref(v), C.byref(d), vega.ctypes.data_as(_dp)))
        return v.value, d.value, vega


_singletons = {}


def get(variant=""):
    """variant "": the checker (no FMA contraction, like the reference's /fp:precise build); "fma": the same sources
    compiled with -ffp-contract=fast, used only to measure how much the reference's own results move with the compiler."""
    if variant not in _singletons:
        path = SO_PATH if not variant else SO_PATH.replace("libcfref.so", f"libcfref_{variant}.so")
        _singletons[variant] = RefLib(path)
    return _singletons[variant]
