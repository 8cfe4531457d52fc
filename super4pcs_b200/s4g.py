"""ctypes binding of libs4g.so (include/s4g.h).  One `Context` = one GPU + the resident clouds."""
import ctypes as C
import os
import re
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_f = np.float32


class S4GError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "lib", "libs4g.so")


def header_path():
    return os.path.join(_ROOT, "include", "s4g.h")


class TcsResult(C.Structure):
    _fields_ = [("key", C.c_uint64), ("best_count", C.c_uint32), ("best_index", C.c_int32),
                ("n_gate_pass", C.c_uint32), ("n_q", C.c_uint32), ("best_T", C.c_float * 16),
                ("best_rms", C.c_float), ("centroid1", C.c_float * 3), ("centroid2", C.c_float * 3),
                ("best_quad", C.c_int32 * 4)]


class BaseDesc(C.Structure):
    _fields_ = [("pair_distance", C.c_float * 2), ("pair_normals_angle", C.c_float * 2), ("base_p", (C.c_float * 9) * 4),
                ("base_xyz_p", C.c_float * 12), ("invariant1", C.c_float), ("invariant2", C.c_float)]


class BaseResult(C.Structure):
    _fields_ = [("n_pairs", C.c_int64 * 2), ("n_quads", C.c_int64), ("tcs", TcsResult)]


class PairFilters(C.Structure):
    _fields_ = [("max_normal_difference", C.c_float), ("max_translation_distance", C.c_float),
                ("max_angle", C.c_float), ("max_color_distance", C.c_float)]


_lib = None


def declared_symbols():
    """names of every function include/s4g.h declares"""
    txt = open(header_path()).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(s4g_[a-z0-9_]+)\s*\(", txt)))


def exported_symbols(names=None):
    """subset of `names` (default: all declared) that the built library exports"""
    L = C.CDLL(lib_path())
    out = []
    for n in (names or declared_symbols()):
        try:
            getattr(L, n)
            out.append(n)
        except AttributeError:
            pass
    return out


def load_library():
    """Loads libs4g.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise S4GError("CUDA extension %s is missing: run `python -c 'import __graft_entry__ as g; "
                       "g.build()'` (there is no CPU fallback)" % p)
    L = C.CDLL(p)
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    sig = {
        "s4g_abi_version": ([], i32),
        "s4g_create": ([i32, C.POINTER(vp)], i32),
        "s4g_destroy": ([vp], None),
        "s4g_error_string": ([vp], C.c_char_p),
        "s4g_set_stream": ([vp, vp], i32),
        "s4g_synchronize": ([vp], i32),
        "s4g_device_count": ([vp], i32),
        "s4g_set_cloud_p": ([vp, vp, i32, f32], i32),
        "s4g_set_cloud_q": ([vp, vp, vp, vp, i32], i32),
        "s4g_get_q_normalization": ([vp, vp], i32),
        "s4g_get_grid_stats": ([vp, vp], i32),
        "s4g_verify": ([vp, vp, i32, vp], i32),
        "s4g_verify_dev": ([vp, vp, i32, vp], i32),
        "s4g_verify_probe_stats": ([vp, vp, i32, vp], i32),
        "s4g_verify_best_dev": ([vp, vp, i32, vp, vp, vp], i32),
        "s4g_verify_best": ([vp, vp, i32, vp, vp, vp], i32),
        "s4g_comm_unique_id": ([vp], i32),
        "s4g_comm_init_rank": ([vp, vp, i32, i32], i32),
        "s4g_comm_init_all": ([vp, i32], i32),
        "s4g_comm_destroy": ([vp], i32),
        "s4g_comm_info": ([vp, vp], i32),
        "s4g_comm_set_timeout": ([vp, i32], i32),
        "s4g_rigid_batch": ([vp, vp, vp, i64, f32, vp, vp, vp], i32),
        "s4g_try_congruent_set": ([vp, vp, vp, i64, f32, f32, i32, i32, C.POINTER(TcsResult)], i32),
        "s4g_try_congruent_set_dev": ([vp, vp, vp, i64, f32, f32, i32, i32, C.POINTER(TcsResult)], i32),
        "s4g_try_congruent_set_resident": ([vp, vp, f32, f32, i32, i32, C.POINTER(TcsResult)], i32),
        "s4g_extract_pairs": ([vp, f32, f32, f32, vp, vp, C.POINTER(PairFilters), i32, C.POINTER(i64)], i32),
        "s4g_get_pairs": ([vp, i32, vp], i32),
        "s4g_set_pairs": ([vp, i32, vp, i64], i32),
        "s4g_count_pairs": ([vp, f32, f32, C.POINTER(i64)], i32),
        "s4g_count_pairs_rows": ([vp, f32, f32, vp, C.POINTER(i64)], i32),
        "s4g_find_quads": ([vp, f32, f32, f32, vp, C.POINTER(i64)], i32),
        "s4g_try_bases": ([vp, vp, i32, f32, C.POINTER(PairFilters), f32, f32, f32, vp], i32),
        "s4g_get_quads": ([vp, vp], i32),
        "s4g_get_timings": ([vp, vp], i32),
        "s4g_voxel_sample": ([vp, vp, i64, f32, vp, C.POINTER(i64)], i32),
    }
    for name, (args, res) in sig.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            continue  # reported by the symbol-export test, not here
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt=_f):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


COMM_ID_BYTES = 128


def comm_unique_id():
    """ncclGetUniqueId through libs4g (rank 0 calls it and ships the bytes to its peers)"""
    buf = np.zeros(COMM_ID_BYTES, np.uint8)
    rc = load_library().s4g_comm_unique_id(_p(buf))
    if rc != 0:
        raise S4GError("s4g_comm_unique_id failed (rc %d): NCCL is not loadable in this process" % rc)
    return buf.tobytes()


def comm_init_all(contexts):
    """one process, several devices: ncclCommInitAll over the contexts (rank = position in the list)"""
    arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
    rc = load_library().s4g_comm_init_all(arr, len(contexts))
    if rc != 0:
        raise S4GError("s4g_comm_init_all failed (rc %d): %s" % (rc, contexts[0]._err()))


def device_count():
    """CUDA devices visible to libs4g.so (s4g_device_count); 0 when there is none"""
    n = C.c_int(0)
    load_library().s4g_device_count(C.byref(n))
    return int(n.value)


class Context:
    """RAII wrapper of s4g_ctx.  Clouds are the CENTRED sampled clouds."""

    def __init__(self, device=0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.s4g_create(int(device), C.byref(h))
        if rc != 0:
            raise S4GError("s4g_create(device=%d) failed with code %d: no usable CUDA device "
                           "(libs4g has no CPU fallback)" % (device, rc))
        self.h = h
        self.device = device
        self.nP = self.nQ = 0
        self.delta = None

    def close(self):
        if getattr(self, "h", None):
            self._L.s4g_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _err(self):
        return self._L.s4g_error_string(self.h).decode()

    def _chk(self, rc):
        if rc != 0:
            raise S4GError("libs4g error %d: %s" % (rc, self._L.s4g_error_string(self.h).decode()))

    # ---- plumbing
    def set_stream(self, cuda_stream_ptr):
        self._chk(self._L.s4g_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._chk(self._L.s4g_synchronize(self.h))

    def timings(self):
        o = np.zeros(5, np.float64)
        self._chk(self._L.s4g_get_timings(self.h, _p(o)))
        return dict(verify_ms=o[0], rigid_ms=o[1], pairs_ms=o[2], quads_ms=o[3], launches=int(o[4]))

    # ---- clouds
    def set_cloud_p(self, P, delta):
        P = _c(P).reshape(-1, 3)
        self._chk(self._L.s4g_set_cloud_p(self.h, _p(P), len(P), float(delta)))
        self.nP, self.delta = len(P), float(delta)

    def set_cloud_q(self, Q, normals=None, rgb=None):
        Q, normals, rgb = _c(Q).reshape(-1, 3), _c(normals), _c(rgb)
        self._chk(self._L.s4g_set_cloud_q(self.h, _p(Q), _p(normals), _p(rgb), len(Q)))
        self.nQ = len(Q)

    def q_normalization(self):
        o = np.zeros(5, _f)
        self._chk(self._L.s4g_get_q_normalization(self.h, _p(o)))
        return o[:3].copy(), float(o[3])

    def grid_stats(self):
        o = np.zeros(6, np.float64)
        self._chk(self._L.s4g_get_grid_stats(self.h, _p(o)))
        return dict(cell_edge=o[0], bricks=int(o[1]), brick_edge=int(o[2]), cells=int(o[3]),
                    points_per_occupied_cell=o[4], resident_bytes=o[5])

    # ---- a8
    def verify(self, T_colmajor):
        T = _c(T_colmajor).reshape(-1, 16)
        counts = np.zeros(len(T), np.uint32)
        self._chk(self._L.s4g_verify(self.h, _p(T), len(T), _p(counts)))
        return counts

    def verify_dev(self, d_T_ptr, K, d_counts_ptr):
        self._chk(self._L.s4g_verify_dev(self.h, C.c_void_p(d_T_ptr), int(K), C.c_void_p(d_counts_ptr)))

    def verify_best_dev(self, d_T_ptr, K, d_index_ptr, d_counts_ptr, d_key_ptr):
        """Verify + first-maximum key (+ the maximum over the communicator's ranks), stream-ordered, nothing synchronised"""
        self._chk(self._L.s4g_verify_best_dev(self.h, C.c_void_p(d_T_ptr), int(K),
                                              C.c_void_p(d_index_ptr) if d_index_ptr else None,
                                              C.c_void_p(d_counts_ptr), C.c_void_p(d_key_ptr)))

    def verify_best(self, T_colmajor, index=None):
        """host buffers -> (counts, key); key = max (count << 32 | 0xFFFFFFFF - index) over this and the peers' lists"""
        T = _c(T_colmajor).reshape(-1, 16)
        idx = None if index is None else _c(index, np.uint32).reshape(-1)
        counts, key = np.zeros(len(T), np.uint32), np.zeros(1, np.uint64)
        self._chk(self._L.s4g_verify_best(self.h, _p(T), len(T), _p(idx), _p(counts), _p(key)))
        return counts, int(key[0])

    # ---- row e: the reduction of a sharded candidate set inside the library (NCCL, loaded on first use)
    def comm_init_rank(self, unique_id, n_ranks, rank):
        buf = np.frombuffer(bytes(unique_id), np.uint8).copy()
        assert buf.size == COMM_ID_BYTES
        self._chk(self._L.s4g_comm_init_rank(self.h, _p(buf), int(n_ranks), int(rank)))

    def comm_destroy(self):
        self._chk(self._L.s4g_comm_destroy(self.h))

    def comm_info(self):
        o = np.zeros(4, np.int32)
        self._chk(self._L.s4g_comm_info(self.h, _p(o)))
        return dict(ranks=int(o[0]), rank=int(o[1]), nccl_version=int(o[2]), collectives=int(o[3]))

    def comm_set_timeout(self, seconds):
        self._chk(self._L.s4g_comm_set_timeout(self.h, int(seconds)))

    def verify_probe_stats(self, T_colmajor):
        T = _c(T_colmajor).reshape(-1, 16)
        o = np.zeros(5, np.uint64)
        self._chk(self._L.s4g_verify_probe_stats(self.h, _p(T), len(T), _p(o)))
        return dict(points_tested=int(o[0]), ranges_read=int(o[1]), brick_entries_read=int(o[2]),
                    bitmap_words_read=int(o[3]), tile_pairs_culled=int(o[4]))

    # ---- a6 / a7
    def rigid_batch(self, base_xyz, quads, max_angle_deg=-1.0):
        b, q = _c(base_xyz).reshape(12), _c(quads, np.int32).reshape(-1, 4)
        K = len(q)
        T, rms, ok = np.zeros((K, 16), _f), np.zeros(K, _f), np.zeros(K, np.int32)
        self._chk(self._L.s4g_rigid_batch(self.h, _p(b), _p(q), K, float(max_angle_deg), _p(T), _p(rms), _p(ok)))
        return T, rms, ok.astype(bool)

    @staticmethod
    def _tcs_dict(r):
        return dict(key=int(r.key), best_count=int(r.best_count), best_index=int(r.best_index),
                    n_gate_pass=int(r.n_gate_pass), n_q=int(r.n_q), T=np.array(r.best_T, _f),
                    rms=float(r.best_rms), centroid1=np.array(r.centroid1, _f),
                    centroid2=np.array(r.centroid2, _f), quad=np.array(r.best_quad, np.int32))

    def try_congruent_set(self, base_xyz, quads, rms_threshold, max_angle_deg=-1.0, shard_rank=0, shard_world=1):
        b, q = _c(base_xyz).reshape(12), _c(quads, np.int32).reshape(-1, 4)
        r = TcsResult()
        self._chk(self._L.s4g_try_congruent_set(self.h, _p(b), _p(q), len(q), float(max_angle_deg),
                                                float(rms_threshold), int(shard_rank), int(shard_world), C.byref(r)))
        return self._tcs_dict(r)

    def try_congruent_set_dev(self, base_xyz, d_quads_ptr, K, rms_threshold, max_angle_deg=-1.0, shard_rank=0,
                              shard_world=1):
        b = _c(base_xyz).reshape(12)
        r = TcsResult()
        self._chk(self._L.s4g_try_congruent_set_dev(self.h, _p(b), C.c_void_p(d_quads_ptr), int(K),
                                                    float(max_angle_deg), float(rms_threshold), int(shard_rank),
                                                    int(shard_world), C.byref(r)))
        return self._tcs_dict(r)

    def try_congruent_set_resident(self, base_xyz, rms_threshold, max_angle_deg=-1.0, shard_rank=0, shard_world=1):
        b = _c(base_xyz).reshape(12)
        r = TcsResult()
        self._chk(self._L.s4g_try_congruent_set_resident(self.h, _p(b), float(max_angle_deg), float(rms_threshold),
                                                         int(shard_rank), int(shard_world), C.byref(r)))
        return self._tcs_dict(r)

    # ---- f1: several bases per launch chain
    def try_bases(self, bases, eps, thr2, rms_threshold, filters=None, max_angle_deg=-1.0):
        """bases: list of dicts(d1, d2, na1, na2, b9 (4 x 9), bxp (4 x 3), inv1, inv2) -> list of dicts(n_pairs, n_quads, tcs)"""
        B = len(bases)
        arr = (BaseDesc * B)()
        for k, b in enumerate(bases):
            arr[k].pair_distance[0], arr[k].pair_distance[1] = float(b["d1"]), float(b["d2"])
            arr[k].pair_normals_angle[0], arr[k].pair_normals_angle[1] = float(b.get("na1", 0.0)), float(b.get("na2", 0.0))
            b9 = _c(b["b9"]).reshape(4, 9)
            for i in range(4):
                for j in range(9):
                    arr[k].base_p[i][j] = float(b9[i, j])
            bxp = _c(b["bxp"]).reshape(12)
            for j in range(12):
                arr[k].base_xyz_p[j] = float(bxp[j])
            arr[k].invariant1, arr[k].invariant2 = float(b["inv1"]), float(b["inv2"])
        res = (BaseResult * B)()
        f = filters if filters is not None else PairFilters(-1, -1, -1, -1)
        self._chk(self._L.s4g_try_bases(self.h, C.byref(arr), B, float(eps), C.byref(f), float(thr2), float(max_angle_deg),
                                        float(rms_threshold), C.byref(res)))
        return [dict(n_pairs=[int(r.n_pairs[0]), int(r.n_pairs[1])], n_quads=int(r.n_quads), tcs=self._tcs_dict(r.tcs)) for r in res]

    # ---- a2 / a3
    def extract_pairs(self, pair_distance, pair_normals_angle, eps, base_p1=None, base_p2=None, filters=None,
                      slot=0, fetch=True):
        """base_p1/base_p2: 9 floats (pos, normal, rgb) of base_3D_[base_point1/2]."""
        d9 = np.array([0, 0, 0, 0, 0, 0, -1, -1, -1], _f)
        b1 = d9 if base_p1 is None else _c(base_p1).reshape(9)
        b2 = d9 if base_p2 is None else _c(base_p2).reshape(9)
        f = filters if filters is not None else PairFilters(-1, -1, -1, -1)
        n = C.c_int64(0)
        self._chk(self._L.s4g_extract_pairs(self.h, float(pair_distance), float(pair_normals_angle), float(eps),
                                            _p(b1), _p(b2), C.byref(f), int(slot), C.byref(n)))
        if not fetch:
            return int(n.value)
        return self.get_pairs(slot, int(n.value))

    def get_pairs(self, slot, n):
        out = np.zeros((n, 2), np.int32)
        if n:
            self._chk(self._L.s4g_get_pairs(self.h, int(slot), _p(out)))
        return out

    def set_pairs(self, slot, pairs):
        p = _c(pairs, np.int32).reshape(-1, 2)
        self._chk(self._L.s4g_set_pairs(self.h, int(slot), _p(p), len(p)))

    def count_pairs(self, pair_distance, eps):
        n = C.c_int64(0)
        self._chk(self._L.s4g_count_pairs(self.h, float(pair_distance), float(eps), C.byref(n)))
        return int(n.value)

    def count_pairs_rows(self, pair_distance, eps):
        """(total, rows) with rows[a] = number of ordered pairs (a, .) of the shell query"""
        n = C.c_int64(0)
        rows = np.zeros(self.nQ, np.uint32)
        self._chk(self._L.s4g_count_pairs_rows(self.h, float(pair_distance), float(eps), _p(rows), C.byref(n)))
        return int(n.value), rows

    # ---- f2
    def voxel_sample(self, xyz, voxel):
        """indices (ascending) of the first point of every voxel of edge `voxel`"""
        X = _c(xyz).reshape(-1, 3)
        out = np.zeros(len(X), np.int32)
        n = C.c_int64(0)
        self._chk(self._L.s4g_voxel_sample(self.h, _p(X), len(X), float(voxel), _p(out), C.byref(n)))
        return out[:n.value].copy()

    # ---- a4 / a5
    def find_quads(self, inv1, inv2, thr2, base_xyz, fetch=True):
        b = _c(base_xyz).reshape(12)
        n = C.c_int64(0)
        self._chk(self._L.s4g_find_quads(self.h, float(inv1), float(inv2), float(thr2), _p(b), C.byref(n)))
        if not fetch:
            return int(n.value)
        out = np.zeros((int(n.value), 4), np.int32)
        if n.value:
            self._chk(self._L.s4g_get_quads(self.h, _p(out)))
        return out
