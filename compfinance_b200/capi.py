"""ctypes binding of compfinance_b200/lib/libcf_b200.so (include/cf_b200.h).

Plumbing only: Python here plays the role of a foreign-language caller of the C ABI (tests, bench).
There is no CPU fallback: if the CUDA library is missing, import fails loudly; if no GPU is present,
every compute call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CF_B200_LIB", os.path.join(_HERE, "lib", "libcf_b200.so"))   # override: kernel experiments only

CF_RNG_SOBOL, CF_RNG_MRG32K3A = 0, 1
CF_MODEL_BS, CF_MODEL_DUPIRE, CF_MODEL_DISPLACED = 0, 1, 2
CF_PRODUCT_EUROPEAN, CF_PRODUCT_UOC, CF_PRODUCT_EUROPEANS = 0, 1, 2
CF_PRODUCT_BASKETS, CF_PRODUCT_AUTOCALL, CF_PRODUCT_MULTISTATS = 3, 4, 5

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)


class cf_rng(C.Structure):
    _fields_ = [("kind", C.c_int32), ("seed1", C.c_uint32), ("seed2", C.c_uint32)]


class cf_model(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n_assets", C.c_int32), ("n_steps", C.c_int32), ("n_events", C.c_int32),
        ("is_event", _u8p), ("spot", C.c_double),
        ("bs_drifts", _dp), ("bs_stds", _dp),
        ("numeraires", _dp), ("fwd_factors", _dp), ("discounts", _dp), ("libors", _dp),
        ("n_knots", C.c_int32), ("log_spots", _dp), ("interp_vols", _dp),
        ("n_times", C.c_int32), ("time_col1", _ip), ("time_col2", _ip), ("time_w1", _dp), ("time_w2", _dp),
        ("dlm_spots", _dp), ("dlm_chol", _dp), ("dlm_alphas", _dp), ("dlm_dynamics", _ip),
        ("dlm_dyn_fwd", _dp), ("dlm_drifts", _dp), ("dlm_stds", _dp), ("dlm_fwd_factors", _dp),
    ]


class cf_product(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n_events", C.c_int32), ("n_payoffs", C.c_int32), ("is_put", C.c_int32),
        ("strike", C.c_double), ("barrier", C.c_double), ("smooth", C.c_double), ("coupon", C.c_double),
        ("strike_offsets", _ip), ("strikes", _dp), ("weights", _dp), ("event_dt", _dp),
    ]


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m compfinance_b200.build` "
                          "(the engine has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.cf_last_error.restype = C.c_char_p
    lib.cf_launch_count.restype = C.c_uint64
    lib.cf_table_adjoint_size.restype = C.c_size_t
    lib.cf_table_adjoint_size.argtypes = [C.POINTER(cf_model), C.POINTER(cf_product)]
    lib.cf_plan_out_size.restype = C.c_size_t
    lib.cf_plan_out_size.argtypes = [C.c_void_p, C.c_int]
    lib.cf_sobol_direction_number.restype = C.c_uint32
    lib.cf_plan_create.argtypes = [C.POINTER(cf_model), C.POINTER(cf_product), C.POINTER(cf_rng), C.POINTER(C.c_void_p)]
    lib.cf_plan_destroy.argtypes = [C.c_void_p]
    lib.cf_plan_destroy.restype = None
    lib.cf_plan_launch_value.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.cf_plan_launch_aad.argtypes = [C.c_void_p, _dp, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.cf_plan_kernel_ms.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_int)]
    lib.cf_plan_debug_times.argtypes = [C.c_void_p, C.c_void_p]
    lib.cf_plan_run_value.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, _dp, _dp]
    lib.cf_plan_run_aad.argtypes = [C.c_void_p, _dp, C.c_uint64, C.c_uint64, _dp, _dp, _dp, _dp, _dp]
    lib.cf_plan_run_aad_multi.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, _dp, _dp]
    lib.cf_last_run_kernel_ms.restype = C.c_double
    lib.cf_comm_create.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_void_p]
    lib.cf_comm_connect.argtypes = [C.c_void_p]
    lib.cf_comm_enable.argtypes = [C.c_int]
    lib.cf_comm_info.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.cf_shard_range.argtypes = [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.cf_run_value.argtypes = [C.POINTER(cf_model), C.POINTER(cf_product), C.POINTER(cf_rng), C.c_uint64,
                                 C.c_uint64, _dp, _dp]
    lib.cf_run_aad.argtypes = [C.POINTER(cf_model), C.POINTER(cf_product), C.POINTER(cf_rng), C.c_uint64,
                               C.c_uint64, _dp, _dp, _dp, _dp, _dp, _dp]
    lib.cf_run_aad_multi.argtypes = [C.POINTER(cf_model), C.POINTER(cf_product), C.POINTER(cf_rng), C.c_uint64,
                                     C.c_uint64, _dp, _dp]
    lib.cf_sobol_states.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
    lib.cf_rng_draw.argtypes = [C.POINTER(cf_rng), C.c_int, C.c_uint64, C.c_uint64, C.c_int, _dp]
    lib.cf_mrg_numerators.argtypes = [C.POINTER(cf_rng), C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
    lib.cf_inv_normal.argtypes = [_dp, _dp, C.c_uint64]
    lib.cf_init.argtypes = [C.c_int, C.POINTER(C.c_int)]
    return lib


EXPORTED = [
    "cf_init", "cf_shutdown", "cf_last_error", "cf_launch_count", "cf_table_adjoint_size", "cf_run_value",
    "cf_run_aad", "cf_run_aad_multi", "cf_plan_create", "cf_plan_destroy", "cf_plan_launch_value", "cf_plan_launch_aad",
    "cf_plan_out_size", "cf_plan_kernel_ms", "cf_plan_run_value", "cf_plan_run_aad", "cf_plan_run_aad_multi", "cf_last_run_kernel_ms", "cf_plan_debug_times", "cf_device_count", "cf_context_generation",
    "cf_comm_create", "cf_comm_connect", "cf_comm_enable", "cf_comm_destroy", "cf_comm_info", "cf_comm_status", "cf_shard_range",
    "cf_sobol_states", "cf_sobol_direction_number", "cf_sobol_max_dim",
    "cf_rng_draw", "cf_mrg_numerators", "cf_inv_normal", "cf_selftest_mrg_uniform", "cf_measure_fp64_peak", "cf_device_sm_count",
]


class CfError(RuntimeError):
    pass


class Engine:
    """Convenience wrapper over the C ABI taking numpy arrays."""

    def __init__(self, device=None, devices=None):
        """device: one CUDA ordinal; devices: a list of ordinals -- a single-process multi-device context (runs are
        sharded over them and summed over peer memory)."""
        self.lib = load()
        self._keep = []
        if devices is not None:
            dev = (C.c_int * len(devices))(*[int(d) for d in devices])
            self._chk(self.lib.cf_init(len(devices), dev))
        elif device is not None:
            dev = (C.c_int * 1)(int(device))
            self._chk(self.lib.cf_init(1, dev))

    COMM_HANDLE_BYTES = 64

    def comm_create(self, world, rank, capacity):
        """This process's receive block of the rank sum; returns its CUDA IPC handle (bytes) for the launcher to gather."""
        buf = C.create_string_buffer(self.COMM_HANDLE_BYTES)
        self._chk(self.lib.cf_comm_create(world, rank, capacity, buf))
        return buf.raw

    def comm_connect(self, handles):
        """handles: the participants' handles in rank order."""
        blob = b"".join(handles)
        self._chk(self.lib.cf_comm_connect(C.c_char_p(blob)))

    def comm_enable(self, on):
        self._chk(self.lib.cf_comm_enable(1 if on else 0))

    def shard_range(self, n_paths, rank, world):
        f, c = C.c_uint64(), C.c_uint64()
        self._chk(self.lib.cf_shard_range(n_paths, rank, world, C.byref(f), C.byref(c)))
        return f.value, c.value

    def _chk(self, rc):
        if rc != 0:
            raise CfError(self.lib.cf_last_error().decode())

    @staticmethod
    def _d(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        return a, a.ctypes.data_as(_dp)

    def rng(self, kind, seed1=12345, seed2=12346):
        return cf_rng(CF_RNG_SOBOL if kind == "sobol" else CF_RNG_MRG32K3A, seed1, seed2)

    def dupire_model(self, spot, log_spots, interp_vols, is_event, n_events, time_map=None):
        """time_map = (n_times, col1[D], col2[D], w1[D], w2[D]) folds Dupire::init() into the device sweep."""
        keep = []
        ls, pls = self._d(log_spots); keep.append(ls)
        iv, piv = self._d(interp_vols); keep.append(iv)
        ev = np.ascontiguousarray(is_event, dtype=np.uint8); keep.append(ev)
        m = cf_model()
        m.kind, m.n_assets, m.n_steps, m.n_events = CF_MODEL_DUPIRE, 1, iv.shape[0], n_events
        m.is_event = ev.ctypes.data_as(_u8p)
        m.spot, m.n_knots, m.log_spots, m.interp_vols = spot, ls.size, pls, piv
        if time_map is not None:
            nt, c1, c2, w1, w2 = time_map
            c1 = np.ascontiguousarray(c1, dtype=np.int32); c2 = np.ascontiguousarray(c2, dtype=np.int32)
            w1, pw1 = self._d(w1); w2, pw2 = self._d(w2)
            keep += [c1, c2, w1, w2]
            m.n_times, m.time_col1, m.time_col2 = nt, c1.ctypes.data_as(_ip), c2.ctypes.data_as(_ip)
            m.time_w1, m.time_w2 = pw1, pw2
        m._keep = keep
        return m

    def bs_model(self, spot, drifts, stds, is_event, numeraires, fwd_factors, discounts):
        keep = []
        dr, pdr = self._d(drifts); keep.append(dr)
        st, pst = self._d(stds); keep.append(st)
        nu, pnu = self._d(numeraires); keep.append(nu)
        ff, pff = self._d(fwd_factors); keep.append(ff)
        di, pdi = self._d(discounts); keep.append(di)
        ev = np.ascontiguousarray(is_event, dtype=np.uint8); keep.append(ev)
        m = cf_model()
        m.kind, m.n_assets, m.n_steps, m.n_events = CF_MODEL_BS, 1, dr.size, nu.size
        m.is_event = ev.ctypes.data_as(_u8p)
        m.spot, m.bs_drifts, m.bs_stds = spot, pdr, pst
        m.numeraires, m.fwd_factors, m.discounts = pnu, pff, pdi
        m._keep = keep
        return m

    def european(self, strike):
        p = cf_product()
        p.kind, p.n_events, p.n_payoffs, p.strike = CF_PRODUCT_EUROPEAN, 1, 1, strike
        return p

    def uoc(self, strike, barrier, smooth_abs, n_events, is_put=False):
        p = cf_product()
        p.kind, p.n_events, p.n_payoffs, p.is_put = CF_PRODUCT_UOC, n_events, 2, int(is_put)
        p.strike, p.barrier, p.smooth = strike, barrier, smooth_abs
        return p

    def run_value(self, mdl, prd, rng, first, n, per_path=False):
        sums = np.zeros(prd.n_payoffs)
        pp = np.zeros((n, prd.n_payoffs)) if per_path else None
        self._chk(self.lib.cf_run_value(C.byref(mdl), C.byref(prd), C.byref(rng), first, n,
                                        sums.ctypes.data_as(_dp), pp.ctypes.data_as(_dp) if per_path else None))
        return (sums, pp) if per_path else sums

    def run_aad(self, mdl, prd, rng, first, n, weights, per_path=False):
        w, pw = self._d(weights)
        sums = np.zeros(prd.n_payoffs)
        agg = np.zeros(1)
        nadj = self.lib.cf_table_adjoint_size(C.byref(mdl), C.byref(prd))
        adj = np.zeros(nadj)
        pp = np.zeros((n, prd.n_payoffs)) if per_path else None
        pa = np.zeros(n) if per_path else None
        self._chk(self.lib.cf_run_aad(C.byref(mdl), C.byref(prd), C.byref(rng), first, n, pw,
                                      sums.ctypes.data_as(_dp), agg.ctypes.data_as(_dp), adj.ctypes.data_as(_dp),
                                      pp.ctypes.data_as(_dp) if per_path else None,
                                      pa.ctypes.data_as(_dp) if per_path else None))
        out = dict(payoff_sums=sums, agg_sum=float(agg[0]), table_adj=adj)
        if per_path:
            out["payoffs"], out["agg"] = pp, pa
        return out

    def sobol_states(self, dim, first, n):
        out = np.zeros((n, dim), dtype=np.uint32)
        self._chk(self.lib.cf_sobol_states(dim, first, n, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def rng_draw(self, rng, dim, first, n, gaussian):
        out = np.zeros((n, dim))
        self._chk(self.lib.cf_rng_draw(C.byref(rng), dim, first, n, int(gaussian), out.ctypes.data_as(_dp)))
        return out

    def mrg_numerators(self, rng, dim, first, n):
        out = np.zeros((n, dim), dtype=np.uint32)
        self._chk(self.lib.cf_mrg_numerators(C.byref(rng), dim, first, n, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def inv_normal(self, p):
        p, pp = self._d(p)
        out = np.zeros_like(p)
        self._chk(self.lib.cf_inv_normal(pp, out.ctypes.data_as(_dp), p.size))
        return out

    def fp64_peak_tflops(self):
        t, ms = C.c_double(), C.c_double()
        self._chk(self.lib.cf_measure_fp64_peak(C.byref(t), C.byref(ms)))
        return t.value

    def direction_number(self, bit, dim):
        return int(self.lib.cf_sobol_direction_number(bit, dim))
