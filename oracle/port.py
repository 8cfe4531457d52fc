"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/liboracle_port.so (oracle/port.cc, the
plain-C++ restatement of the Super4PCS hot path)."""
import ctypes as C
import numpy as np
from . import _build

_f = np.float32
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build_port())
        L.port_create.restype = C.c_void_p
        L.port_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float]
        L.port_destroy.argtypes = [C.c_void_p]
        L.port_get_normalization.argtypes = [C.c_void_p, C.c_void_p]
        L.port_rigid_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_float,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        L.port_verify_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_float, C.c_int,
                                        C.c_void_p, C.c_void_p]
        L.port_verify_batch.restype = C.c_double
        L.port_verify_bruteforce.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        L.port_try_congruent_set.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_float,
                                             C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.port_extract_pairs.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.port_extract_pairs.restype = C.c_long
        L.port_extract_pairs_ordered.argtypes = L.port_extract_pairs.argtypes
        L.port_extract_pairs_ordered.restype = C.c_long
        L.port_get_pairs.argtypes = [C.c_void_p, C.c_void_p]
        L.port_find_quads.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                      C.c_void_p, C.c_long, C.c_void_p, C.c_long]
        L.port_find_quads.restype = C.c_long
        L.port_get_quads.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt=_f):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


class Port:
    """P, Q are the CENTRED sampled clouds."""

    def __init__(self, P, Q, delta, Qn=None, Qrgb=None):
        self._L = lib()
        P, Q, Qn, Qrgb = _c(P), _c(Q), _c(Qn), _c(Qrgb)
        self.nP, self.nQ, self.delta = len(P), len(Q), float(delta)
        self.h = self._L.port_create(_p(P), len(P), _p(Q), _p(Qn), _p(Qrgb), len(Q), float(delta))

    def close(self):
        if self.h:
            self._L.port_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def normalization(self):
        o = np.empty(5, _f)
        self._L.port_get_normalization(self.h, _p(o))
        return o[:3].copy(), float(o[3])

    def rigid_batch(self, base_ids, quads, max_angle_deg=-1.0):
        b, q = _c(base_ids, np.int32), _c(quads, np.int32).reshape(-1, 4)
        K = len(q)
        T, rms, ok = np.empty((K, 16), _f), np.empty(K, _f), np.empty(K, np.int32)
        self._L.port_rigid_batch(self.h, _p(b), _p(q), K, float(max_angle_deg), _p(T), _p(rms), _p(ok))
        return T, rms, ok.astype(bool)

    def verify_batch(self, T16, best_lcp=0.0, nthreads=1):
        T16 = _c(T16).reshape(-1, 16)
        lcp, good = np.empty(len(T16), _f), np.empty(len(T16), np.uint32)
        secs = self._L.port_verify_batch(self.h, _p(T16), len(T16), float(best_lcp), int(nthreads),
                                         _p(lcp), _p(good))
        return lcp, good, secs

    def verify_bruteforce(self, T16):
        T16 = _c(T16).reshape(-1, 16)
        good = np.empty(len(T16), np.uint32)
        self._L.port_verify_bruteforce(self.h, _p(T16), len(T16), _p(good))
        return good

    def try_congruent_set(self, base_ids, quads, best_lcp_in, max_angle_deg=-1.0):
        b, q = _c(base_ids, np.int32), _c(quads, np.int32).reshape(-1, 4)
        st, T = np.empty(2, _f), np.empty(16, _f)
        bi = C.c_long(-1)
        self._L.port_try_congruent_set(self.h, _p(b), _p(q), len(q), float(max_angle_deg),
                                       float(best_lcp_in), _p(st), C.addressof(bi), _p(T))
        return dict(best_lcp=float(st[0]), n_gate=int(st[1]), best_index=int(bi.value), T=T.copy())


DEFAULT_BASE9 = np.array([0, 0, 0, 0, 0, 0, -1, -1, -1], _f)


def _extract_pairs(self, pair_distance, pair_normals_angle, eps, base_p1=None, base_p2=None,
                   filters=(-1.0, -1.0, -1.0, -1.0)):
    """filters = (max_normal_difference, max_translation_distance, max_angle, max_color_distance)"""
    b1 = DEFAULT_BASE9 if base_p1 is None else _c(base_p1).reshape(9)
    b2 = DEFAULT_BASE9 if base_p2 is None else _c(base_p2).reshape(9)
    f4 = _c(np.array(filters, _f))
    n = self._L.port_extract_pairs(self.h, float(pair_distance), float(pair_normals_angle), float(eps),
                                   _p(b1), _p(b2), _p(f4))
    out = np.empty((n, 2), np.int32)
    if n:
        self._L.port_get_pairs(self.h, _p(out))
    return out


Port.extract_pairs = _extract_pairs


def _extract_pairs_ordered(self, pair_distance, pair_normals_angle, eps, base_p1=None, base_p2=None,
                           filters=(-1.0, -1.0, -1.0, -1.0)):
    """same pair SET in the reference's EMISSION order (octree traversal; stateful like the reference: the order
    depends on every earlier ordered call on this handle)"""
    b1 = DEFAULT_BASE9 if base_p1 is None else _c(base_p1).reshape(9)
    b2 = DEFAULT_BASE9 if base_p2 is None else _c(base_p2).reshape(9)
    f4 = _c(np.array(filters, _f))
    n = self._L.port_extract_pairs_ordered(self.h, float(pair_distance), float(pair_normals_angle), float(eps),
                                           _p(b1), _p(b2), _p(f4))
    out = np.empty((n, 2), np.int32)
    if n:
        self._L.port_get_pairs(self.h, _p(out))
    return out


Port.extract_pairs_ordered = _extract_pairs_ordered


def _find_quads(self, inv1, inv2, thr2, base_xyz, pairs1, pairs2):
    b = _c(base_xyz).reshape(12)
    p1, p2 = _c(pairs1, np.int32).reshape(-1, 2), _c(pairs2, np.int32).reshape(-1, 2)
    n = self._L.port_find_quads(self.h, float(inv1), float(inv2), float(thr2), _p(b), _p(p1), len(p1), _p(p2), len(p2))
    out = np.empty((n, 4), np.int32)
    if n:
        self._L.port_get_quads(self.h, _p(out))
    return out


Port.find_quads = _find_quads


def num_threads():
    return int(lib().port_num_threads())
