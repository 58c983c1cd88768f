"""quick-adc_b200 — B200-native Quick ADC search path.

This package is a thin ctypes binding over the C ABI in include/qadc_b200.h
(libqadc_b200.so, hand-written CUDA for sm_100a).  It is what tests/ and bench.py use; the
drop-in C++ host mirror of the reference's scanner/engine API lives in host/.

There is NO CPU path: everything here fails loudly when the CUDA library is missing or no
sm_100 device is present.  Import with ``importlib.import_module("quick-adc_b200")`` (the
directory name carries the reference's hyphen) or through the ``qadc_b200`` alias module at
the repository root.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libqadc_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "qadc_b200.h")

QADC_OK, QADC_EINVAL, QADC_ECUDA, QADC_ESTATE, QADC_EBOUND, QADC_ENOMEM = 0, -1, -2, -3, -4, -5


class QadcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"qadc error {code}: {msg}")
        self.code = code


class Metrics(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("index_us", "rotate_us", "table_us", "scan_us", "h2d_us", "d2h_us")]


_lib = None


def load_library(build_if_missing=True):
    """Loads libqadc_b200.so (building it with nvcc when absent or stale). Never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    # Build only when the library is absent (never when merely older than the sources: several
    # ranks may import concurrently, and file times do not survive a copy to the GPU box);
    # __graft_entry__.build() / quick-adc_b200/build.py rebuild explicitly.
    if build_if_missing and not os.path.exists(LIB_PATH) and os.path.exists(_build.NVCC):
        try:
            _build.build(force=True)
        except Exception as e:
            raise RuntimeError(f"cannot build libqadc_b200.so: {e}")
    path = os.environ.get("QADC_LIB", LIB_PATH)   # development knob: an alternative build of the same library
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(path)
    vp, i32, u32, f32 = C.c_void_p, C.c_int, C.c_uint32, C.c_float
    L.qadc_abi_version.restype = i32
    L.qadc_create.argtypes = [i32, vp, C.POINTER(vp)]
    L.qadc_destroy.argtypes = [vp]
    L.qadc_destroy.restype = None
    L.qadc_last_error.argtypes = [vp]
    L.qadc_last_error.restype = C.c_char_p
    L.qadc_set_pq.argtypes = [vp, i32, i32, i32, vp, vp]
    L.qadc_set_coarse.argtypes = [vp, i32, vp]
    L.qadc_begin_database.argtypes = [vp, i32, vp, i32]
    L.qadc_upload_codes.argtypes = [vp, i32, u32, u32, vp, vp, i32]
    L.qadc_upload_database.argtypes = [vp, vp, vp, i32]
    L.qadc_upload_partitions.argtypes = [vp, vp, vp]
    L.qadc_set_position_base.argtypes = [vp, i32, u32]
    L.qadc_set_prefix.argtypes = [vp, i32, vp, u32, i32]
    L.qadc_set_prefixes.argtypes = [vp, vp, vp, i32]
    L.qadc_finalize.argtypes = [vp, f32]
    L.qadc_search.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp]
    L.qadc_search_device.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp]
    L.qadc_search_assigned_device.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    L.qadc_coarse_partial_device.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    L.qadc_tables_local_device.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    L.qadc_set_owned_partitions.argtypes = [vp, vp]
    L.qadc_search_bounded_device.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    L.qadc_coarse_merge_device.argtypes = [vp, vp, i32, i32, i32, vp]
    L.qadc_adc_load.argtypes = [vp, i32, vp, vp, vp]
    L.qadc_adc_search.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    L.qadc_synchronize.argtypes = [vp]
    L.qadc_last_launch_count.argtypes = [vp]
    L.qadc_last_scan_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.qadc_scan_ms_history.argtypes = [vp, C.POINTER(C.c_float), i32]
    L.qadc_scan_ms_history.restype = i32
    L.qadc_merge_shards_device.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    L.qadc_build_tables.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.qadc_scan_with_tables.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp]
    L.qadc_dump_distances.argtypes = [vp, i32, vp, vp]
    L.qadc_download_codes.argtypes = [vp, i32, vp]
    L.qadc_set_option.argtypes = [vp, C.c_char_p, C.c_long]
    L.qadc_encode.argtypes = [vp, vp, u32, vp, vp]
    L.qadc_multi_create.argtypes = [vp, i32, C.POINTER(vp)]
    L.qadc_multi_destroy.argtypes = [vp]
    L.qadc_multi_destroy.restype = None
    L.qadc_multi_last_error.argtypes = [vp]
    L.qadc_multi_last_error.restype = C.c_char_p
    L.qadc_multi_device_count.argtypes = [vp]
    L.qadc_multi_uses_nccl.argtypes = [vp]
    L.qadc_multi_context.argtypes = [vp, i32]
    L.qadc_multi_context.restype = vp
    L.qadc_multi_set_pq.argtypes = [vp, i32, i32, i32, vp, vp]
    L.qadc_multi_set_coarse.argtypes = [vp, i32, vp]
    L.qadc_multi_load.argtypes = [vp, i32, vp, vp, vp, f32]
    L.qadc_multi_search.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp]
    for name in ("qadc_multi_create", "qadc_multi_device_count", "qadc_multi_uses_nccl", "qadc_multi_set_pq",
                 "qadc_multi_set_coarse", "qadc_multi_load", "qadc_multi_search"):
        getattr(L, name).restype = i32
    for name in ("qadc_set_prefixes", "qadc_upload_database", "qadc_upload_partitions", "qadc_create", "qadc_set_pq", "qadc_set_coarse", "qadc_begin_database", "qadc_upload_codes",
                 "qadc_set_position_base", "qadc_set_prefix", "qadc_finalize", "qadc_search", "qadc_search_device",
                 "qadc_synchronize", "qadc_last_launch_count", "qadc_last_scan_ms", "qadc_merge_shards_device",
                 "qadc_build_tables", "qadc_scan_with_tables", "qadc_dump_distances", "qadc_download_codes",
                 "qadc_set_option", "qadc_encode", "qadc_search_assigned_device", "qadc_coarse_partial_device",
                 "qadc_coarse_merge_device", "qadc_adc_load", "qadc_adc_search", "qadc_tables_local_device",
                 "qadc_search_bounded_device", "qadc_set_owned_partitions"):
        getattr(L, name).restype = i32
    _lib = L
    return L


def _ptr(a):
    """numpy array / int device pointer / None -> void*."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return a.ctypes.data_as(C.c_void_p)


class Index:
    """One GPU-resident Quick ADC database = scanner_4 + engine state (db_query_4.cpp:73-310,
    query_common.hpp:149-309) behind the C ABI."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.qadc_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc:
            raise QadcError(rc, self.lib.qadc_last_error(None).decode())
        self.h = h
        self.dim = self.m = 0
        self.K = 0

    def _ck(self, rc):
        if rc:
            raise QadcError(rc, self.lib.qadc_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.qadc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- quantisers ---------------------------------------------------------------------
    def set_pq(self, dim, m, codebooks, rotation=None, bits=4):
        cb = np.ascontiguousarray(codebooks, np.float32).reshape(-1)
        rot = None if rotation is None else np.ascontiguousarray(rotation, np.float32).reshape(-1)
        self._ck(self.lib.qadc_set_pq(self.h, dim, m, bits, _ptr(cb), _ptr(rot)))
        self.dim, self.m, self.bits = dim, m, bits

    def set_coarse(self, centroids):
        c = np.ascontiguousarray(centroids, np.float32)
        self._ck(self.lib.qadc_set_coarse(self.h, c.shape[0], _ptr(c)))
        self.K = c.shape[0]

    # ---- database -----------------------------------------------------------------------
    def begin_database(self, sizes, has_labels):
        s = np.ascontiguousarray(sizes, np.uint32)
        self._ck(self.lib.qadc_begin_database(self.h, len(s), _ptr(s), int(has_labels)))
        self.sizes = s

    def upload_codes(self, part_i, first, codes, labels=None):
        codes = np.ascontiguousarray(codes, np.uint8)
        lab = None if labels is None else np.ascontiguousarray(labels, np.uint32)
        self._ck(self.lib.qadc_upload_codes(self.h, part_i, first, codes.shape[0], _ptr(codes), _ptr(lab), 0))

    def upload_database(self, codes, labels=None):
        """All partitions at once: rows in partition order (qadc_upload_database)."""
        codes = np.ascontiguousarray(codes, np.uint8)
        lab = None if labels is None else np.ascontiguousarray(labels, np.uint32)
        self._ck(self.lib.qadc_upload_database(self.h, _ptr(codes), _ptr(lab), 0))

    def upload_database_device(self, d_codes, d_labels=None):
        self._ck(self.lib.qadc_upload_database(self.h, _ptr(int(d_codes)), _ptr(None if d_labels is None else int(d_labels)), 1))

    def upload_codes_device(self, part_i, first, count, d_codes, d_labels=None):
        self._ck(self.lib.qadc_upload_codes(self.h, part_i, first, count, _ptr(int(d_codes)),
                                            _ptr(None if d_labels is None else int(d_labels)), 1))

    def set_position_base(self, part_i, base):
        self._ck(self.lib.qadc_set_position_base(self.h, part_i, base))

    def set_prefix(self, part_i, codes):
        codes = np.ascontiguousarray(codes, np.uint8)
        self._ck(self.lib.qadc_set_prefix(self.h, part_i, _ptr(codes), codes.shape[0], 0))

    def set_prefix_device(self, part_i, d_codes, count):
        self._ck(self.lib.qadc_set_prefix(self.h, part_i, _ptr(int(d_codes)), count, 1))

    def set_prefixes_device(self, d_codes, counts):
        """Explicit keep-prefixes of all partitions at once (device pointer, counts[p] vectors each)."""
        c = np.ascontiguousarray(counts, np.uint32)
        self._ck(self.lib.qadc_set_prefixes(self.h, _ptr(int(d_codes)), _ptr(c), 1))

    def set_prefixes(self, codes, counts):
        codes = np.ascontiguousarray(codes, np.uint8)
        c = np.ascontiguousarray(counts, np.uint32)
        self._ck(self.lib.qadc_set_prefixes(self.h, _ptr(codes), _ptr(c), 0))

    def finalize(self, keep):
        self._ck(self.lib.qadc_finalize(self.h, np.float32(keep)))

    def load_flat(self, codes, keep):
        """flat_db: one partition, no labels (databases.hpp:77-134)."""
        codes = np.ascontiguousarray(codes, np.uint8)
        self.begin_database([codes.shape[0]], False)
        self.upload_codes(0, 0, codes)
        self.finalize(keep)

    def load_ivf(self, codes, labels, offsets, keep):
        """index_db: partition p = rows offsets[p]:offsets[p+1] of codes/labels (databases.hpp:176-250)."""
        offsets = np.asarray(offsets, np.int64)
        sizes = np.diff(offsets).astype(np.uint32)
        self.begin_database(sizes, True)
        lo, hi = int(offsets[0]), int(offsets[-1])
        self.upload_database(codes[lo:hi], labels[lo:hi])
        self.finalize(keep)

    def encode(self, vectors):
        """PQ-encode float vectors (and, for an IVF context, assign them to coarse cells)."""
        v = np.ascontiguousarray(vectors, np.float32)
        codes = np.empty((v.shape[0], self.m * getattr(self, "bits", 4) // 8), np.uint8)
        assign = np.empty(v.shape[0], np.int32) if self.K else None
        self._ck(self.lib.qadc_encode(self.h, _ptr(v), v.shape[0], _ptr(assign), _ptr(codes)))
        return (codes, assign) if self.K else codes

    # ---- search -------------------------------------------------------------------------
    def search(self, queries, ma, r, want_metrics=False):
        q = np.ascontiguousarray(queries, np.float32)
        nq = q.shape[0]
        ids = np.empty((nq, r), np.uint32)
        d = np.empty((nq, r), np.int8)
        cnt = np.empty(nq, np.int32)
        met = Metrics()
        self._ck(self.lib.qadc_search(self.h, _ptr(q), nq, ma, r, _ptr(ids), _ptr(d), _ptr(cnt), C.byref(met)))
        return (ids, d, cnt, met) if want_metrics else (ids, d, cnt)

    def search_host_buffers(self, q_ptr, nq, ma, r, ids_ptr, d_ptr, cnt_ptr):
        """qadc_search on raw host pointers (pinned buffers owned by the caller)."""
        self._ck(self.lib.qadc_search(self.h, _ptr(int(q_ptr)), nq, ma, r, _ptr(int(ids_ptr)), _ptr(int(d_ptr)),
                                      _ptr(int(cnt_ptr)), None))

    def search_device(self, d_queries, nq, ma, r, d_ids, d_dists, d_counts, d_keys=None):
        self._ck(self.lib.qadc_search_device(self.h, _ptr(int(d_queries)), nq, ma, r, _ptr(int(d_ids)),
                                             _ptr(int(d_dists)), _ptr(int(d_counts)),
                                             _ptr(None if d_keys is None else int(d_keys))))

    def search_assigned_device(self, d_queries, d_assign, nq, ma, r, d_ids, d_dists, d_counts, d_keys=None):
        """qadc_search_device with the coarse assignment (nq*ma int32 on the device) supplied."""
        self._ck(self.lib.qadc_search_assigned_device(self.h, _ptr(int(d_queries)), _ptr(int(d_assign)), nq, ma, r,
                                                      _ptr(int(d_ids)), _ptr(int(d_dists)), _ptr(int(d_counts)),
                                                      _ptr(None if d_keys is None else int(d_keys))))

    def set_owned_partitions(self, owned):
        """Owner-computes: boolean mask [partition_count] of the partitions this shard answers for (empty ones included)."""
        o = np.ascontiguousarray(np.asarray(owned).astype(np.uint8))
        self._ck(self.lib.qadc_set_owned_partitions(self.h, _ptr(o)))

    def tables_local_device(self, d_queries, d_assign, nq, ma, r, d_local):
        """Owner-computes step 1: tables of the owned probes; d_local [nq][r+1] floats = (min entry, r smallest prefix distances)."""
        self._ck(self.lib.qadc_tables_local_device(self.h, _ptr(int(d_queries)), _ptr(int(d_assign)), nq, ma, r, _ptr(int(d_local))))

    def search_bounded_device(self, d_gathered, G, nq, ma, r, d_ids, d_dists, d_counts, d_keys=None):
        """Owner-computes step 3: bounds from the gathered [G][nq][r+1] shares, int8 tables, scan of the owned lists."""
        self._ck(self.lib.qadc_search_bounded_device(self.h, _ptr(int(d_gathered)), G, nq, ma, r, _ptr(int(d_ids)),
                                                     _ptr(int(d_dists)), _ptr(int(d_counts)),
                                                     _ptr(None if d_keys is None else int(d_keys))))

    def coarse_partial_device(self, d_queries, nq, ma, c_first, c_count, d_out_keys):
        """This rank's ma best cells among [c_first, c_first + c_count) as uint64 keys [nq][ma]."""
        self._ck(self.lib.qadc_coarse_partial_device(self.h, _ptr(int(d_queries)), nq, ma, c_first, c_count,
                                                     _ptr(int(d_out_keys))))

    def coarse_merge_device(self, d_keys, G, nq, ma, d_assign):
        """ma smallest of the gathered [G][nq][ma] keys per query -> assignment [nq][ma] int32."""
        self._ck(self.lib.qadc_coarse_merge_device(self.h, _ptr(int(d_keys)), G, nq, ma, _ptr(int(d_assign))))

    # ---- plain ADC (db_query) -------------------------------------------------------------
    def adc_load(self, codes, labels=None, offsets=None):
        """Row-major codes [n, m*bits/8]; inverted lists: labels [n] and offsets [K+1]."""
        codes = np.ascontiguousarray(codes, np.uint8)
        off = np.ascontiguousarray([0, codes.shape[0]] if offsets is None else offsets, np.uint64)
        lab = None if labels is None else np.ascontiguousarray(labels, np.uint32)
        self._ck(self.lib.qadc_adc_load(self.h, len(off) - 1, _ptr(off), _ptr(codes), _ptr(lab)))

    def adc_search(self, queries, ma, r):
        q = np.ascontiguousarray(queries, np.float32)
        nq = q.shape[0]
        ids = np.empty((nq, r), np.uint32)
        d = np.empty((nq, r), np.float32)
        cnt = np.empty(nq, np.int32)
        self._ck(self.lib.qadc_adc_search(self.h, _ptr(q), nq, ma, r, _ptr(ids), _ptr(d), _ptr(cnt)))
        return ids, d, cnt

    def synchronize(self):
        self._ck(self.lib.qadc_synchronize(self.h))

    def merge_shards_device(self, d_keys, d_ids, G, nq, r, d_out_ids, d_out_dists, d_out_counts, d_out_keys=None):
        self._ck(self.lib.qadc_merge_shards_device(self.h, _ptr(int(d_keys)), _ptr(None if d_ids is None else int(d_ids)),
                                                   G, nq, r, _ptr(int(d_out_ids)), _ptr(int(d_out_dists)),
                                                   _ptr(int(d_out_counts)),
                                                   _ptr(None if d_out_keys is None else int(d_out_keys))))

    def last_launch_count(self):
        return self.lib.qadc_last_launch_count(self.h)

    def last_scan_ms(self):
        ms = C.c_float()
        self._ck(self.lib.qadc_last_scan_ms(self.h, C.byref(ms)))
        return ms.value

    def scan_ms_history(self, n):
        buf = (C.c_float * n)()
        got = self.lib.qadc_scan_ms_history(self.h, buf, n)
        if got < 0:
            self._ck(got)
        return [buf[i] for i in range(got)]

    def set_option(self, key, value):
        self._ck(self.lib.qadc_set_option(self.h, key.encode(), int(value)))

    # ---- parity entry points ------------------------------------------------------------
    def build_tables(self, queries, ma, r, assign_in=None):
        q = np.ascontiguousarray(queries, np.float32)
        nq = q.shape[0]
        out = dict(assign=np.empty((nq, ma), np.int32), tables=np.empty((nq, ma, self.m, 16), np.float32),
                   qmin=np.empty(nq, np.float32), qmax=np.empty(nq, np.float32),
                   qtables=np.empty((nq, ma, self.m, 16), np.int8))
        ai = None if assign_in is None else np.ascontiguousarray(assign_in, np.int32)
        rc = self.lib.qadc_build_tables(self.h, _ptr(q), nq, ma, r, _ptr(ai), _ptr(out["assign"]), _ptr(out["tables"]),
                                        _ptr(out["qmin"]), _ptr(out["qmax"]), _ptr(out["qtables"]))
        out["rc"] = rc
        if rc and rc != QADC_EBOUND:
            self._ck(rc)
        return out

    def scan_with_tables(self, assign, qtables, r):
        a = np.ascontiguousarray(assign, np.int32)
        nq, ma = a.shape
        t = np.ascontiguousarray(qtables, np.int8)
        ids = np.empty((nq, r), np.uint32)
        d = np.empty((nq, r), np.int8)
        cnt = np.empty(nq, np.int32)
        self._ck(self.lib.qadc_scan_with_tables(self.h, _ptr(a), _ptr(t), nq, ma, r, _ptr(ids), _ptr(d), _ptr(cnt)))
        return ids, d, cnt

    def dump_distances(self, part_i, qtable):
        t = np.ascontiguousarray(qtable, np.int8).reshape(-1)
        out = np.empty(int(self.sizes[part_i]), np.int8)
        self._ck(self.lib.qadc_dump_distances(self.h, part_i, _ptr(t), _ptr(out)))
        return out

    def download_codes(self, part_i):
        out = np.empty((int(self.sizes[part_i]), self.m // 2), np.uint8)
        self._ck(self.lib.qadc_download_codes(self.h, part_i, _ptr(out)))
        return out


class MultiIndex:
    """One database sharded over several GPUs by ONE process (qadc_multi_*): what the db_query_4 CLI uses for
    `-g 0,1,...`.  A device listed more than once = virtual shards on that GPU (tests on a single-GPU box)."""

    def __init__(self, devices):
        self.lib = load_library()
        dv = np.ascontiguousarray(devices, np.int32)
        h = C.c_void_p()
        rc = self.lib.qadc_multi_create(_ptr(dv), len(dv), C.byref(h))
        if rc:
            raise QadcError(rc, self.lib.qadc_multi_last_error(None).decode())
        self.h = h
        self.dim = self.m = self.K = 0

    def _ck(self, rc):
        if rc:
            raise QadcError(rc, self.lib.qadc_multi_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.qadc_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def uses_nccl(self):
        return bool(self.lib.qadc_multi_uses_nccl(self.h))

    def set_option(self, key, value):
        for g in range(self.lib.qadc_multi_device_count(self.h)):
            ctx = C.c_void_p(self.lib.qadc_multi_context(self.h, g))
            if self.lib.qadc_set_option(ctx, key.encode(), int(value)):
                raise QadcError(QADC_EINVAL, self.lib.qadc_last_error(ctx).decode())

    def set_pq(self, dim, m, codebooks, rotation=None, bits=4):
        cb = np.ascontiguousarray(codebooks, np.float32).reshape(-1)
        rot = None if rotation is None else np.ascontiguousarray(rotation, np.float32).reshape(-1)
        self._ck(self.lib.qadc_multi_set_pq(self.h, dim, m, bits, _ptr(cb), _ptr(rot)))
        self.dim, self.m = dim, m

    def set_coarse(self, centroids):
        c = np.ascontiguousarray(centroids, np.float32)
        self._ck(self.lib.qadc_multi_set_coarse(self.h, c.shape[0], _ptr(c)))
        self.K = c.shape[0]

    def _load(self, parts_codes, parts_labels, keep):
        P = len(parts_codes)
        sizes = np.array([c.shape[0] for c in parts_codes], np.uint32)
        pc = (C.c_void_p * P)(*[c.ctypes.data if c.shape[0] else None for c in parts_codes])
        pl = None
        if parts_labels is not None:
            pl = (C.c_void_p * P)(*[l.ctypes.data if l.shape[0] else None for l in parts_labels])
        self._ck(self.lib.qadc_multi_load(self.h, P, _ptr(sizes), pc, pl, np.float32(keep)))

    def load_flat(self, codes, keep):
        self._load([np.ascontiguousarray(codes, np.uint8)], None, keep)

    def load_ivf(self, codes, labels, offsets, keep):
        codes = np.ascontiguousarray(codes, np.uint8)
        labels = np.ascontiguousarray(labels, np.uint32)
        K = len(offsets) - 1
        self._load([codes[offsets[p]:offsets[p + 1]] for p in range(K)], [labels[offsets[p]:offsets[p + 1]] for p in range(K)], keep)

    def search(self, queries, ma, r, want_metrics=False):
        q = np.ascontiguousarray(queries, np.float32)
        nq = q.shape[0]
        ids = np.empty((nq, r), np.uint32)
        d = np.empty((nq, r), np.int8)
        cnt = np.empty(nq, np.int32)
        met = Metrics()
        self._ck(self.lib.qadc_multi_search(self.h, _ptr(q), nq, ma, r, _ptr(ids), _ptr(d), _ptr(cnt), C.byref(met)))
        return (ids, d, cnt, met) if want_metrics else (ids, d, cnt)
