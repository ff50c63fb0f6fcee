"""ctypes driver of mex/_build/libekfslam_mextest.so = mex/ekfslam_mex.c + the minimal mxArray runtime
(mex/stub/mexrt.c) + libekfslam.so: builds mxArrays from Python values, calls mexFunction like Octave would,
converts the outputs back.  TEST INFRASTRUCTURE."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "mex", "_build", "libekfslam_mextest.so")
_lib = None
_P = C.c_void_p


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(SO)
        sig = {"mxCreateDoubleMatrix": (_P, [C.c_size_t, C.c_size_t, C.c_int]), "mxCreateString": (_P, [C.c_char_p]),
               "mxCreateStructMatrix": (_P, [C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_char_p)]),
               "mxGetPr": (C.POINTER(C.c_double), [_P]), "mxGetM": (C.c_size_t, [_P]), "mxGetN": (C.c_size_t, [_P]),
               "mxIsStruct": (C.c_int, [_P]), "mxIsChar": (C.c_int, [_P]), "mxArrayToString": (_P, [_P]),
               "mxGetNumberOfFields": (C.c_int, [_P]), "mxGetFieldNameByNumber": (C.c_char_p, [_P, C.c_int]),
               "mxGetFieldByNumber": (_P, [_P, C.c_size_t, C.c_int]), "mxSetFieldByNumber": (None, [_P, C.c_size_t, C.c_int, _P]),
               "mxGetNumberOfElements": (C.c_size_t, [_P]),
               "mexrt_call": (C.c_int, [C.c_int, C.POINTER(_P), C.c_int, C.POINTER(_P)]),
               "mexrt_last_error": (C.c_char_p, []), "mexrt_last_error_id": (C.c_char_p, []),
               "mexrt_lock_count": (C.c_int, []), "mexrt_run_atexit": (None, [])}
        for k, (r, a) in sig.items():
            fn = getattr(L, k)
            fn.restype, fn.argtypes = r, a
        _lib = L
    return _lib


class MexError(RuntimeError):
    def __init__(self, ident, msg):
        super().__init__("%s: %s" % (ident, msg))
        self.ident = ident


def to_mx(v):
    L = lib()
    if v is None:
        return L.mxCreateDoubleMatrix(0, 0, 0)
    if isinstance(v, str):
        return L.mxCreateString(v.encode())
    if isinstance(v, dict):
        v = [v]
    if isinstance(v, list) and (not v or isinstance(v[0], dict)):
        if not v:
            return L.mxCreateDoubleMatrix(0, 0, 0)
        names = list(v[0].keys())
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        s = L.mxCreateStructMatrix(1, len(v), len(names), arr)
        for i, e in enumerate(v):
            for k, n in enumerate(names):
                L.mxSetFieldByNumber(s, i, k, to_mx(e[n]))
        return s
    a = np.atleast_2d(np.asarray(v, dtype=np.float64))
    m = L.mxCreateDoubleMatrix(a.shape[0], a.shape[1], 0)
    if a.size:
        af = np.asfortranarray(a)          # keep the temporary alive across the copy
        C.memmove(L.mxGetPr(m), af.ctypes.data, a.size * 8)
        del af
    return m


def from_mx(p):
    L = lib()
    if not p:
        return None
    if L.mxIsStruct(p):
        nf, ne = L.mxGetNumberOfFields(p), L.mxGetNumberOfElements(p)
        names = [L.mxGetFieldNameByNumber(p, k).decode() for k in range(nf)]
        out = [{n: from_mx(L.mxGetFieldByNumber(p, i, k)) for k, n in enumerate(names)} for i in range(ne)]
        return out
    if L.mxIsChar(p):
        return C.cast(L.mxArrayToString(p), C.c_char_p).value.decode()
    m, n = L.mxGetM(p), L.mxGetN(p)
    if m * n == 0:
        return None
    buf = np.ctypeslib.as_array(L.mxGetPr(p), shape=(m * n,)).copy()
    return buf.reshape((m, n), order="F")


def call(cmd, *args, nargout=1):
    """ekfslam_mex(cmd, args...) -> tuple of outputs (Python values); raises MexError like Octave's error()."""
    L = lib()
    prhs = (_P * (len(args) + 1))(to_mx(cmd), *[to_mx(a) for a in args])
    plhs = (_P * max(nargout, 1))()
    if L.mexrt_call(nargout, plhs, len(args) + 1, prhs):
        raise MexError(L.mexrt_last_error_id().decode(), L.mexrt_last_error().decode())
    outs = tuple(from_mx(plhs[i]) for i in range(nargout))
    return outs[0] if nargout == 1 else outs
