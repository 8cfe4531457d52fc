"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/liboracle_ref.so (the unmodified
reference + oracle/ref_harness.cc).  All matrices cross as column-major 16-float arrays
(Eigen's default storage); helpers here expose them as numpy (K,4,4) row-indexed arrays."""
import ctypes as C
import numpy as np
from . import _build

_f = np.float32


class RefOptions(C.Structure):
    _fields_ = [("delta", C.c_float), ("max_normal_difference", C.c_float),
                ("max_translation_distance", C.c_float), ("max_angle", C.c_float),
                ("max_color_distance", C.c_float), ("sample_size", C.c_uint64),
                ("max_time_seconds", C.c_int32), ("random_seed", C.c_uint32),
                ("overlap", C.c_float), ("terminate_threshold", C.c_float)]


def make_options(delta=5.0, max_normal_difference=-1.0, max_translation_distance=-1.0,
                 max_angle=-1.0, max_color_distance=-1.0, sample_size=200,
                 max_time_seconds=60, random_seed=5489, overlap=0.2, terminate_threshold=1.0):
    return RefOptions(delta, max_normal_difference, max_translation_distance, max_angle,
                      max_color_distance, sample_size, max_time_seconds, random_seed,
                      overlap, terminate_threshold)


_libs = {}


def available():
    return _build.build_ref() is not None


def lib(path=None):
    """binding of a library exporting the ref_* harness ABI (oracle/ref_harness.cc).  Default: the
    compiled reference.  tests/test_dropin_gpu.py passes the SAME harness compiled against the
    product's drop-in headers (super4pcs_b200/lib/libb200_harness.so)."""
    if path is None:
        path = _build.build_ref()
        if path is None:
            raise RuntimeError("reference oracle not available (no /root/reference and no prebuilt "
                               "oracle/_ref/liboracle_ref.so)")
    if path not in _libs:
        L = C.CDLL(path)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                 C.POINTER(RefOptions), C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        for n in ("ref_n_sampled_p", "ref_n_sampled_q"):
            getattr(L, n).argtypes = [C.c_void_p]
        for n in ("ref_get_sampled_p", "ref_get_sampled_q", "ref_get_base3d"):
            getattr(L, n).argtypes = [C.c_void_p] * 4
        L.ref_get_init_state.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_set_best_lcp.argtypes = [C.c_void_p, C.c_float]
        L.ref_get_best_lcp.argtypes = [C.c_void_p]
        L.ref_get_best_lcp.restype = C.c_float
        L.ref_set_base3d.argtypes = [C.c_void_p] * 4
        L.ref_select_quadrilateral.argtypes = [C.c_void_p] * 4
        L.ref_extract_pairs.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]
        L.ref_extract_pairs.restype = C.c_long
        L.ref_get_pairs.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_find_quads.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_void_p, C.c_long, C.c_void_p, C.c_long]
        L.ref_find_quads.restype = C.c_long
        L.ref_get_quads.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_rigid_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_verify_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_float, C.c_int, C.c_void_p]
        L.ref_verify_batch.restype = C.c_double
        L.ref_try_congruent_set.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_perform_n_steps.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_compute_transformation.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                 C.POINTER(RefOptions), C.c_void_p]
        L.ref_compute_transformation.restype = C.c_float
        L.ref_compute_transformation_traced.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(RefOptions),
                                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _libs[path] = L
    return _libs[path]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt=_f):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def colmajor_to_mats(T16):
    """(K,16) column-major -> (K,4,4) with [k,r,c]."""
    return np.ascontiguousarray(np.asarray(T16, dtype=_f).reshape(-1, 4, 4).transpose(0, 2, 1))


def mats_to_colmajor(M):
    return np.ascontiguousarray(np.asarray(M, dtype=_f).reshape(-1, 4, 4).transpose(0, 2, 1)).reshape(-1, 16)


class RefMatcher:
    """The reference's MatchSuper4PCS after init(P, Q)."""

    def __init__(self, P, Q, options, Pn=None, Prgb=None, Qn=None, Qrgb=None, identity_sampler=True,
                 libpath=None):
        L = lib(libpath)
        self._L = L
        P, Q = _c(P), _c(Q)
        Pn, Prgb, Qn, Qrgb = _c(Pn), _c(Prgb), _c(Qn), _c(Qrgb)
        self.options = options
        self.h = L.ref_create(_p(P), _p(Pn), _p(Prgb), len(P), _p(Q), _p(Qn), _p(Qrgb), len(Q),
                              C.byref(options), 1 if identity_sampler else 0)
        self.nP = L.ref_n_sampled_p(self.h)
        self.nQ = L.ref_n_sampled_q(self.h)

    def close(self):
        if self.h:
            self._L.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _cloud(self, fn, n):
        xyz, nrm, rgb = (np.empty((n, 3), _f) for _ in range(3))
        fn(self.h, _p(xyz), _p(nrm), _p(rgb))
        return xyz, nrm, rgb

    def sampled_p(self):
        return self._cloud(self._L.ref_get_sampled_p, self.nP)

    def sampled_q(self):
        return self._cloud(self._L.ref_get_sampled_q, self.nQ)

    def init_state(self):
        o = np.empty(10, _f)
        self._L.ref_get_init_state(self.h, _p(o))
        return dict(best_lcp=float(o[0]), diameter=float(o[1]), max_base_diameter=float(o[2]),
                    number_of_trials=int(o[3]), centroid_p=o[4:7].copy(), centroid_q=o[7:10].copy())

    def set_best_lcp(self, v):
        self._L.ref_set_best_lcp(self.h, float(v))

    def best_lcp(self):
        return float(self._L.ref_get_best_lcp(self.h))

    def set_base3d(self, xyz, nrm=None, rgb=None):
        xyz, nrm, rgb = _c(xyz), _c(nrm), _c(rgb)
        self._L.ref_set_base3d(self.h, _p(xyz), _p(nrm), _p(rgb))

    def base3d(self):
        return self._cloud(self._L.ref_get_base3d, 4)

    def select_quadrilateral(self):
        i1, i2 = C.c_float(), C.c_float()
        ids = np.zeros(4, np.int32)
        ok = self._L.ref_select_quadrilateral(self.h, C.addressof(i1), C.addressof(i2), _p(ids))
        return bool(ok), float(i1.value), float(i2.value), ids

    def extract_pairs(self, d, normal_angle, eps, b1, b2, sort=True):
        n = self._L.ref_extract_pairs(self.h, float(d), float(normal_angle), float(eps), int(b1), int(b2))
        out = np.empty((n, 2), np.int32)
        if n:
            self._L.ref_get_pairs(self.h, _p(out))
        if sort and n:
            out = out[np.lexsort((out[:, 1], out[:, 0]))]
        return out

    def find_quads(self, inv1, inv2, thr1, thr2, pairs1, pairs2):
        p1, p2 = _c(pairs1, np.int32), _c(pairs2, np.int32)
        n = self._L.ref_find_quads(self.h, float(inv1), float(inv2), float(thr1), float(thr2),
                                   _p(p1), len(p1), _p(p2), len(p2))
        out = np.empty((n, 4), np.int32)
        if n:
            self._L.ref_get_quads(self.h, _p(out))
        return out

    def rigid_batch(self, base_ids, quads):
        b, q = _c(base_ids, np.int32), _c(quads, np.int32).reshape(-1, 4)
        K = len(q)
        T = np.empty((K, 16), _f)
        rms = np.empty(K, _f)
        ok = np.empty(K, np.int32)
        self._L.ref_rigid_batch(self.h, _p(b), _p(q), K, _p(T), _p(rms), _p(ok))
        return T, rms, ok.astype(bool)

    def verify_batch(self, T16, best_lcp=0.0, nthreads=1):
        T16 = _c(T16).reshape(-1, 16)
        out = np.empty(len(T16), _f)
        secs = self._L.ref_verify_batch(self.h, _p(T16), len(T16), float(best_lcp), int(nthreads), _p(out))
        return out, secs

    def perform_n_steps(self, n):
        """Match4PCSBase::Perform_N_steps on the live matcher (RNG / base state stay probe-able)"""
        st = np.empty(3, _f)
        T = np.empty(16, _f)
        r = self._L.ref_perform_n_steps(self.h, int(n), _p(st), _p(T))
        return dict(ret=bool(r), best_lcp=float(st[0]), n_progress=int(st[1]), n_candidates=int(st[2]), T=T)

    def try_congruent_set(self, base_ids, quads):
        b, q = _c(base_ids, np.int32), _c(quads, np.int32).reshape(-1, 4)
        st = np.empty(3, _f)
        T = np.empty(16, _f)
        ids = np.empty(8, np.int32)
        r = self._L.ref_try_congruent_set(self.h, _p(b), _p(q), len(q), _p(st), _p(T), _p(ids))
        return dict(ret=bool(r), best_lcp=float(st[0]), n_gate=int(st[1]), n_visits=int(st[2]),
                    T=T.copy(), base=ids[:4].copy(), congruent=ids[4:].copy())


def compute_transformation(P, Q, options, Pn=None, Qn=None, Prgb=None, Qrgb=None, libpath=None):
    L = lib(libpath)
    P, Q = _c(P), _c(Q).copy()
    Pn, Qn, Prgb, Qrgb = _c(Pn), _c(Qn), _c(Prgb), _c(Qrgb)
    T = np.empty(16, _f)
    score = L.ref_compute_transformation(_p(P), _p(Pn), _p(Prgb), len(P), _p(Q), _p(Qn), _p(Qrgb), len(Q),
                                         C.byref(options), _p(T))
    return float(score), T, Q


def compute_transformation_traced(P, Q, options, libpath=None, max_trace=4096):
    """whole pipeline through a base-class pointer with a global-transform visitor; returns
    (score, T16, trace[n,18]) where trace rows are (fraction, best_LCP, 4x4 column-major) per RANSAC iteration"""
    L = lib(libpath)
    P, Q = _c(P), _c(Q)
    score = C.c_float(0)
    T = np.empty(16, _f)
    tr = np.zeros((max_trace, 18), _f)
    n = L.ref_compute_transformation_traced(_p(P), len(P), _p(Q), len(Q), C.byref(options), C.addressof(score),
                                            _p(T), _p(tr), max_trace)
    return float(score.value), T, tr[:min(n, max_trace)].copy()


def num_threads():
    return int(lib().ref_num_threads())
