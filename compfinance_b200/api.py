"""Python binding of the host API library (compfinance_b200/lib/libcf_host.so = cf_export.cpp over
cf_main.h): the reference's main.h entry points and store, running on the CUDA engine.

Plumbing only (ctypes); there is no CPU fallback: the library must be built
(`python -m compfinance_b200.build`) and a GPU must be present for any simulation call.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "lib", "libcf_host.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

EXPORTED = [
    "cfx_last_error", "cfx_init", "cfx_init_devices", "cfx_drop_sessions", "cfx_set_system_time", "cfx_put_black_scholes", "cfx_put_dupire",
    "cfx_put_european", "cfx_put_barrier", "cfx_put_contingent", "cfx_put_europeans", "cfx_put_displaced", "cfx_put_multistats",
    "cfx_put_baskets", "cfx_put_autocall", "cfx_num_payoffs", "cfx_num_params",
    "cfx_payoff_labels", "cfx_param_labels", "cfx_product_timeline", "cfx_value", "cfx_simul_paths",
    "cfx_aad_risk_one", "cfx_simul_aad_paths", "cfx_simul_aad_multi_paths", "cfx_run_range", "cfx_aad_risk_aggregate", "cfx_aad_risk_multi", "cfx_bump_risk", "cfx_dupire_aad_risk", "cfx_dupire_calib", "cfx_dupire_superbucket",
    "cfx_describe", "cfx_rng_sequence",
]


class CfHostError(RuntimeError):
    pass


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


class CompFinance:
    """Stateful facade: models and products live in the library's global stores (store.h semantics)."""

    def __init__(self, device=None, devices=None):
        if not os.path.exists(HOST_LIB_PATH):
            raise ImportError(f"{HOST_LIB_PATH} is missing: build it with `python -m compfinance_b200.build`")
        self.lib = C.CDLL(HOST_LIB_PATH)
        self.lib.cfx_last_error.restype = C.c_char_p
        self.lib.cfx_set_system_time.argtypes = [C.c_double]
        if devices is not None:
            ids = (C.c_int * len(devices))(*[int(d) for d in devices])
            self._chk(self.lib.cfx_init_devices(C.c_int(len(devices)), ids))
        elif device is not None:
            self._chk(self.lib.cfx_init(C.c_int(int(device))))

    def _chk(self, rc):
        if rc != 0:
            raise CfHostError(self.lib.cfx_last_error().decode())

    # -- store ---------------------------------------------------------------------------------
    def put_black_scholes(self, spot, vol, spot_measure, rate, div, id_):
        self._chk(self.lib.cfx_put_black_scholes(C.c_double(spot), C.c_double(vol), C.c_int(int(spot_measure)),
                                                 C.c_double(rate), C.c_double(div), id_.encode()))

    def put_dupire(self, spot, spots, times, vols, max_dt, id_):
        spots, ps = _d(spots)
        times, pt = _d(times)
        vols, pv = _d(vols)
        assert vols.shape == (spots.size, times.size)
        self._chk(self.lib.cfx_put_dupire(C.c_double(spot), ps, C.c_int(spots.size), pt, C.c_int(times.size), pv,
                                          C.c_double(max_dt), id_.encode()))

    def put_european(self, strike, exercise, settlement, id_):
        self._chk(self.lib.cfx_put_european(C.c_double(strike), C.c_double(exercise), C.c_double(settlement), id_.encode()))

    def put_contingent(self, coupon, maturity, pay_freq, smooth, id_):
        """xPutContingent (xlExport.cpp:320): contingent floater, Black-Scholes only on the device."""
        self._chk(self.lib.cfx_put_contingent(C.c_double(coupon), C.c_double(maturity), C.c_double(pay_freq),
                                              C.c_double(smooth), id_.encode()))

    def put_barrier(self, strike, barrier, maturity, freq, smooth, call_put, id_):
        self._chk(self.lib.cfx_put_barrier(C.c_double(strike), C.c_double(barrier), C.c_double(maturity),
                                           C.c_double(freq), C.c_double(smooth), C.c_int(int(call_put)), id_.encode()))

    def put_europeans(self, maturities, strikes, id_):
        m, pm = _d(maturities)
        k, pk = _d(strikes)
        self._chk(self.lib.cfx_put_europeans(pm, pk, C.c_int(m.size), id_.encode()))

    def put_displaced(self, spots, atms, skews, disc_rate, repo_spreads, div_dates, divs, correl, lam, id_):
        """Multi-asset displaced-lognormal model (xPutDLM, xlExport.cpp:186); assets are named a0, a1, ..."""
        spots, ps = _d(spots); atms, pa = _d(atms); skews, pk = _d(skews); repo, pr = _d(repo_spreads)
        dd, pdd = _d(div_dates); dv, pdv = _d(divs); co, pc = _d(correl)
        n = spots.size
        assert co.shape == (n, n) and (dd.size == 0 or dv.shape == (dd.size, n))
        self._chk(self.lib.cfx_put_displaced(C.c_int(n), ps, pa, pk, C.c_double(disc_rate), pr, pdd, C.c_int(dd.size), pdv, pc,
                                             C.c_double(lam), id_.encode()))

    def put_multistats(self, n_assets, fix_dates, fwd_dates, id_):
        f, pf = _d(fix_dates); w, pw = _d(fwd_dates)
        self._chk(self.lib.cfx_put_multistats(C.c_int(n_assets), pf, pw, C.c_int(f.size), id_.encode()))

    def put_baskets(self, weights, maturity, strikes, id_):
        w, pw = _d(weights); k, pk = _d(strikes)
        self._chk(self.lib.cfx_put_baskets(C.c_int(w.size), pw, C.c_double(maturity), pk, C.c_int(k.size), id_.encode()))

    def put_autocall(self, refs, maturity, periods, ko, strike, cpn, smooth, id_):
        r, pr = _d(refs)
        self._chk(self.lib.cfx_put_autocall(C.c_int(r.size), pr, C.c_double(maturity), C.c_int(periods), C.c_double(ko),
                                            C.c_double(strike), C.c_double(cpn), C.c_double(smooth), id_.encode()))

    def num_payoffs(self, product):
        n = self.lib.cfx_num_payoffs(product.encode())
        if n < 0:
            raise CfHostError("product not found")
        return n

    def num_params(self, model):
        n = self.lib.cfx_num_params(model.encode())
        if n < 0:
            raise CfHostError("model not found")
        return n

    def _labels(self, fn, id_):
        n = fn(id_.encode(), None, C.c_int(0))
        if n < 0:
            raise CfHostError("not found")
        buf = C.create_string_buffer(n)
        fn(id_.encode(), buf, C.c_int(n))
        return buf.value.decode().split("\n")[:-1]

    def payoff_labels(self, product):
        return self._labels(self.lib.cfx_payoff_labels, product)

    def param_labels(self, model):
        return self._labels(self.lib.cfx_param_labels, model)

    def product_timeline(self, product):
        n = self.lib.cfx_product_timeline(product.encode(), None, C.c_int(0))
        out = np.empty(n)
        self.lib.cfx_product_timeline(product.encode(), out.ctypes.data_as(_dp), C.c_int(n))
        return out

    # -- host-only inspection ----------------------------------------------------------------------
    def describe(self, model, product, aad=False):
        dims = (C.c_int * 6)()
        self._chk(self.lib.cfx_describe(model.encode(), product.encode(), C.c_int(int(aad)), dims, *([None] * 11)))
        D, E, m, nT, nadj, today = list(dims)
        is_event = np.zeros(D + 1, dtype=np.uint8)
        na = D * m if m > 0 else D
        nb = m if m > 0 else D
        tab_a, tab_b = np.zeros(na), np.zeros(nb)
        has_ev = m == 0
        num, ff, disc = (np.zeros(E), np.zeros(E), np.zeros(E)) if has_ev else (None, None, None)
        c1 = np.zeros(D, dtype=np.int32); c2 = np.zeros(D, dtype=np.int32); w1 = np.zeros(D); w2 = np.zeros(D)
        pc = np.zeros(3)
        p = lambda a: a.ctypes.data_as(_dp) if a is not None else None   # noqa: E731
        self._chk(self.lib.cfx_describe(model.encode(), product.encode(), C.c_int(int(aad)), dims,
                                        is_event.ctypes.data_as(C.POINTER(C.c_ubyte)), p(tab_a), p(tab_b), p(num), p(ff),
                                        p(disc), c1.ctypes.data_as(_ip), c2.ctypes.data_as(_ip), p(w1), p(w2), p(pc)))
        return dict(n_steps=D, n_events=E, n_knots=m, n_times=nT, adjoint_size=nadj, first_sample_is_today=bool(today),
                    is_event=is_event, tab_a=tab_a.reshape(D, m) if m > 0 else tab_a, tab_b=tab_b, numeraires=num,
                    fwd_factors=ff, discounts=disc, time_map=(nT, c1, c2, w1, w2) if nT > 0 else None,
                    strike=pc[0], barrier=pc[1], smooth=pc[2])

    # -- entry points -----------------------------------------------------------------------------
    def value(self, model, product, n_path, sobol=True, parallel=True, seed1=12345, seed2=12346):
        out = np.empty(self.num_payoffs(product))
        self._chk(self.lib.cfx_value(model.encode(), product.encode(), C.c_int(int(sobol)), C.c_int(seed1), C.c_int(seed2),
                                     C.c_int(n_path), C.c_int(int(parallel)), out.ctypes.data_as(_dp)))
        return out

    def simul_paths(self, model, product, n_path, sobol=True, parallel=True, seed1=12345, seed2=12346):
        out = np.empty((n_path, self.num_payoffs(product)))
        self._chk(self.lib.cfx_simul_paths(model.encode(), product.encode(), C.c_int(int(sobol)), C.c_int(seed1),
                                           C.c_int(seed2), C.c_int(n_path), C.c_int(int(parallel)), out.ctypes.data_as(_dp)))
        return out

    def aad_risk_one(self, model, product, n_path, risk_payoff=-1, sobol=True, parallel=True, seed1=12345, seed2=12346):
        pv = np.empty(self.num_payoffs(product))
        risks = np.empty(self.num_params(model))
        rv = C.c_double()
        self._chk(self.lib.cfx_aad_risk_one(model.encode(), product.encode(), C.c_int(risk_payoff), C.c_int(int(sobol)),
                                            C.c_int(seed1), C.c_int(seed2), C.c_int(n_path), C.c_int(int(parallel)),
                                            pv.ctypes.data_as(_dp), C.byref(rv), risks.ctypes.data_as(_dp)))
        return pv, rv.value, risks

    def simul_aad_paths(self, model, product, n_path, risk_payoff=-1, sobol=True, parallel=True, seed1=12345, seed2=12346):
        npay = self.num_payoffs(product)
        pays, agg, risks = np.empty((n_path, npay)), np.empty(n_path), np.empty(self.num_params(model))
        self._chk(self.lib.cfx_simul_aad_paths(model.encode(), product.encode(), C.c_int(risk_payoff), C.c_int(int(sobol)),
                                               C.c_int(seed1), C.c_int(seed2), C.c_int(n_path), C.c_int(int(parallel)),
                                               pays.ctypes.data_as(_dp), agg.ctypes.data_as(_dp), risks.ctypes.data_as(_dp)))
        return pays, agg, risks

    def aad_risk_aggregate(self, model, product, notionals, n_path, sobol=True, parallel=True, seed1=12345, seed2=12346):
        nots, pn = _d(notionals)
        pv = np.empty(self.num_payoffs(product))
        risks = np.empty(self.num_params(model))
        rv = C.c_double()
        self._chk(self.lib.cfx_aad_risk_aggregate(model.encode(), product.encode(), pn, C.c_int(int(sobol)), C.c_int(seed1),
                                                  C.c_int(seed2), C.c_int(n_path), C.c_int(int(parallel)),
                                                  pv.ctypes.data_as(_dp), C.byref(rv), risks.ctypes.data_as(_dp)))
        return pv, rv.value, risks

    def simul_aad_multi_paths(self, model, product, n_path, sobol=True, parallel=True, seed1=12345, seed2=12346):
        """mcSimulAADMulti (mcBase.h:776): per-path payoffs [nPath][nPay] and risks [nParam][nPay] (sums over paths / nPath)."""
        npay, npar = self.num_payoffs(product), self.num_params(model)
        pays, risks = np.empty((n_path, npay)), np.empty((npar, npay))
        self._chk(self.lib.cfx_simul_aad_multi_paths(model.encode(), product.encode(), C.c_int(int(sobol)), C.c_int(seed1),
                                                     C.c_int(seed2), C.c_int(n_path), C.c_int(int(parallel)),
                                                     pays.ctypes.data_as(_dp), risks.ctypes.data_as(_dp)))
        return pays, risks

    def run_range(self, model, product, first_path, n_paths, weights=None, sobol=True, seed1=12345, seed2=12346, n_adjoints=0):
        """Sums over the paths [first_path, first_path + n_paths) through cf_run_value (weights None) or cf_run_aad:
        (payoff sums,) or (payoff sums, aggregate sum, table adjoints[n_adjoints])."""
        npay = self.num_payoffs(product)
        sums = np.empty(npay)
        if weights is None:
            self._chk(self.lib.cfx_run_range(model.encode(), product.encode(), C.c_int(0), C.c_int(int(sobol)), C.c_int(seed1), C.c_int(seed2),
                                             C.c_ulonglong(first_path), C.c_ulonglong(n_paths), None, sums.ctypes.data_as(_dp), None, None))
            return (sums,)
        w, pw = _d(weights)
        agg, adj = C.c_double(), np.empty(n_adjoints)
        self._chk(self.lib.cfx_run_range(model.encode(), product.encode(), C.c_int(1), C.c_int(int(sobol)), C.c_int(seed1), C.c_int(seed2),
                                         C.c_ulonglong(first_path), C.c_ulonglong(n_paths), pw, sums.ctypes.data_as(_dp), C.byref(agg),
                                         adj.ctypes.data_as(_dp)))
        return sums, agg.value, adj

    def aad_risk_multi(self, model, product, n_path, sobol=True, parallel=True, seed1=12345, seed2=12346):
        """AADriskMulti (main.h:269): values [nPay], risks [nParam][nPay]."""
        npay, npar = self.num_payoffs(product), self.num_params(model)
        values, risks = np.empty(npay), np.empty((npar, npay))
        self._chk(self.lib.cfx_aad_risk_multi(model.encode(), product.encode(), C.c_int(int(sobol)), C.c_int(seed1), C.c_int(seed2),
                                              C.c_int(n_path), C.c_int(int(parallel)), values.ctypes.data_as(_dp),
                                              risks.ctypes.data_as(_dp)))
        return values, risks

    def bump_risk(self, model, product, n_path, sobol=True, parallel=True, seed1=12345, seed2=12346):
        npay, npar = self.num_payoffs(product), self.num_params(model)
        values, risks = np.empty(npay), np.empty((npar, npay))
        self._chk(self.lib.cfx_bump_risk(model.encode(), product.encode(), C.c_int(int(sobol)), C.c_int(seed1), C.c_int(seed2),
                                         C.c_int(n_path), C.c_int(int(parallel)), values.ctypes.data_as(_dp),
                                         risks.ctypes.data_as(_dp)))
        return values, risks

    def dupire_aad_risk(self, model, product, notionals, n_spots, n_times, n_path, sobol=True, parallel=True,
                        seed1=12345, seed2=12346):
        nots, pn = _d(notionals)
        vega = np.empty((n_spots, n_times))
        v, d = C.c_double(), C.c_double()
        self._chk(self.lib.cfx_dupire_aad_risk(model.encode(), product.encode(), pn, C.c_int(int(sobol)), C.c_int(seed1),
                                               C.c_int(seed2), C.c_int(n_path), C.c_int(int(parallel)), C.byref(v),
                                               C.byref(d), vega.ctypes.data_as(_dp)))
        return v.value, d.value, vega

    def dupire_calib(self, incl_spots, max_ds, incl_times, max_dt, spot, vol, jmp_intens=0.0, jmp_avg=0.0, jmp_std=0.0):
        """dupireCalib (main.h:413): (spots, times, lvols[nSpots][nTimes]) calibrated to a Merton surface.  Host only."""
        s, ps = _d(incl_spots); t, pt = _d(incl_times)
        cap = 1 << 16
        spots, times, lv = np.empty(cap), np.empty(cap), np.empty(cap)
        ns, nt = C.c_int(), C.c_int()
        self._chk(self.lib.cfx_dupire_calib(ps, C.c_int(s.size), C.c_double(max_ds), pt, C.c_int(t.size), C.c_double(max_dt),
                                            C.c_double(spot), C.c_double(vol), C.c_double(jmp_intens), C.c_double(jmp_avg),
                                            C.c_double(jmp_std), C.byref(ns), C.byref(nt), spots.ctypes.data_as(_dp),
                                            times.ctypes.data_as(_dp), lv.ctypes.data_as(_dp), C.c_int(cap)))
        return spots[:ns.value].copy(), times[:nt.value].copy(), lv[:ns.value * nt.value].reshape(ns.value, nt.value).copy()

    def dupire_superbucket(self, spot, max_dt, product, notionals, incl_spots, max_ds, incl_times, max_dt_vol, strikes, mats,
                           vol, jmp_intens, jmp_avg, jmp_std, n_path, sobol=True, parallel=True, seed1=12345, seed2=12346,
                           bump=False):
        """dupireSuperbucket (main.h:453) / dupireSuperbucketBump (main.h:575): value, delta, vega[nStrikes][nMats]."""
        nots, pn = _d(notionals); s, ps = _d(incl_spots); t, pt = _d(incl_times); k, pk = _d(strikes); m, pm = _d(mats)
        vega = np.empty((k.size, m.size))
        v, d = C.c_double(), C.c_double()
        self._chk(self.lib.cfx_dupire_superbucket(C.c_double(spot), C.c_double(max_dt), product.encode(), pn, ps, C.c_int(s.size),
                                                  C.c_double(max_ds), pt, C.c_int(t.size), C.c_double(max_dt_vol), pk,
                                                  C.c_int(k.size), pm, C.c_int(m.size), C.c_double(vol), C.c_double(jmp_intens),
                                                  C.c_double(jmp_avg), C.c_double(jmp_std), C.c_int(int(sobol)), C.c_int(seed1),
                                                  C.c_int(seed2), C.c_int(n_path), C.c_int(int(parallel)), C.c_int(int(bump)),
                                                  C.byref(v), C.byref(d), vega.ctypes.data_as(_dp)))
        return v.value, d.value, vega

    def rng_sequence(self, sobol, dim, skip, n, gaussian, seed1=12345, seed2=12346):
        out = np.empty((n, dim))
        self._chk(self.lib.cfx_rng_sequence(C.c_int(int(sobol)), C.c_int(seed1), C.c_int(seed2), C.c_int(dim),
                                            C.c_uint(skip), C.c_int(n), C.c_int(int(gaussian)), out.ctypes.data_as(_dp)))
        return out
