"""ctypes view of the Excel C API data model of compfinance_b200/host/cf_xlcall.h (XLOPER12, FP12), for tests that call
the x... wrappers the way Excel does: counted UTF-16 strings, row-major ranges, results owned by the library."""
import ctypes as C

import numpy as np

XLTYPE_NUM, XLTYPE_STR, XLTYPE_ERR, XLTYPE_MULTI = 0x1, 0x2, 0x10, 0x40
XLERR_NA = 42


class XLOPER12(C.Structure):
    pass


class _Array(C.Structure):
    _fields_ = [("lparray", C.POINTER(XLOPER12)), ("rows", C.c_int32), ("columns", C.c_int32)]


class _Val(C.Union):
    _fields_ = [("num", C.c_double), ("str", C.POINTER(C.c_uint16)), ("err", C.c_int32), ("array", _Array), ("pad", C.c_ubyte * 24)]


XLOPER12._fields_ = [("val", _Val), ("xltype", C.c_uint32)]
assert C.sizeof(XLOPER12) == 32
LPX = C.POINTER(XLOPER12)


def xstr(s, keep):
    """XLOPER12 string; `keep` collects the buffers that must outlive the call."""
    buf = (C.c_uint16 * (len(s) + 1))(len(s), *[ord(c) for c in s])
    x = XLOPER12()
    x.xltype = XLTYPE_STR
    x.val.str = C.cast(buf, C.POINTER(C.c_uint16))
    keep.extend([buf, x])
    return C.byref(x)


def xstrs(strings, keep, cols=1):
    """xltypeMulti range of strings (row major)."""
    n = len(strings)
    arr = (XLOPER12 * n)()
    for i, s in enumerate(strings):
        buf = (C.c_uint16 * (len(s) + 1))(len(s), *[ord(c) for c in s])
        arr[i].xltype = XLTYPE_STR
        arr[i].val.str = C.cast(buf, C.POINTER(C.c_uint16))
        keep.append(buf)
    x = XLOPER12()
    x.xltype = XLTYPE_MULTI
    x.val.array.lparray = C.cast(arr, LPX)
    x.val.array.rows, x.val.array.columns = n // cols, cols
    keep.extend([arr, x])
    return C.byref(x)


def fp12(a, keep):
    """FP12 from a 1-D (one column) or 2-D array."""
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    rows, cols = (a.shape[0], 1) if a.ndim == 1 else a.shape
    n = max(rows * cols, 1)

    class FP(C.Structure):
        _fields_ = [("rows", C.c_int32), ("columns", C.c_int32), ("array", C.c_double * n)]

    f = FP(rows, cols, (C.c_double * n)(*a.ravel()))
    keep.append(f)
    return C.cast(C.byref(f), C.c_void_p)


def _cell(x):
    if x.xltype == XLTYPE_NUM:
        return x.val.num
    if x.xltype == XLTYPE_STR:
        n = x.val.str[0]
        return "".join(chr(x.val.str[1 + k]) for k in range(n))
    if x.xltype == XLTYPE_ERR:
        return ("#ERR", x.val.err)
    raise ValueError(f"unexpected xltype {x.xltype}")


def read(p):
    """Result of a wrapper -> python: number, str, ('#ERR', code) or a list of rows."""
    if not p:
        return None
    x = p.contents
    if x.xltype != XLTYPE_MULTI:
        return _cell(x)
    r, c = x.val.array.rows, x.val.array.columns
    return [[_cell(x.val.array.lparray[i * c + j]) for j in range(c)] for i in range(r)]


NA = ("#ERR", XLERR_NA)
D = C.c_double
SIGS = {   # name: (restype, argtypes)
    "xRestartThreadPool": (D, [D]),
    "xPutDupire": (LPX, [D, C.c_void_p, C.c_void_p, C.c_void_p, D, LPX]),
    "xPutBlackScholes": (LPX, [D, D, D, D, D, LPX]),
    "xPutDLM": (LPX, [LPX, C.c_void_p, C.c_void_p, C.c_void_p, D, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, D, LPX]),
    "xPutEuropean": (LPX, [D, D, D, LPX]),
    "xPutBarrier": (LPX, [D, D, D, D, D, LPX, LPX]),
    "xPutContingent": (LPX, [D, D, D, D, LPX]),
    "xPutEuropeans": (LPX, [C.c_void_p, C.c_void_p, LPX]),
    "xPutMultiStats": (LPX, [LPX, C.c_void_p, C.c_void_p, LPX]),
    "xPutBaskets": (LPX, [LPX, C.c_void_p, D, C.c_void_p, LPX]),
    "xPutAutocall": (LPX, [LPX, C.c_void_p, D, D, D, D, D, D, LPX]),
    "xPayoffIds": (LPX, [LPX]),
    "xParameters": (LPX, [LPX]),
    "xValue": (LPX, [LPX, LPX, D, D, D, D, D]),
    "xValueTime": (LPX, [LPX, LPX, D, D, D, D, D]),
    "xAADrisk": (LPX, [LPX, LPX, LPX, D, D, D, D, D]),
    "xAADriskAggregate": (LPX, [LPX, LPX, LPX, C.c_void_p, D, D, D, D, D]),
    "xBumprisk": (LPX, [LPX, LPX, D, D, D, D, D, D, LPX]),
    "xAADriskMulti": (LPX, [LPX, LPX, D, D, D, D, D, D, LPX]),
    "xDisplayRisk": (LPX, [LPX, LPX]),
    "xDupireCalib": (LPX, [D, D, D, D, D, C.c_void_p, D, C.c_void_p, D]),
    "xDupireSuperbucket": (LPX, [D, D, D, D, D, C.c_void_p, C.c_void_p, C.c_void_p, D, C.c_void_p, D, D, LPX, LPX, C.c_void_p,
                                 D, D, D, D, D, D]),
    "xMerton": (D, [D, D, D, D, D, D, D]),
    "xSobolPoints": (LPX, [D, D, D, D]),
}


def bind(lib):
    """Declare the signatures of the 24 wrappers on a loaded libcf_host.so; raises AttributeError if one is missing."""
    for name, (res, args) in SIGS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib
