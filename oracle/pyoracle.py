"""ctypes loaders for the CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``Oracle``  -> oracle/libqadc_oracle.so   (plain-C restatement, oracle/qadc_oracle.c)
* ``Ref``     -> oracle/_ref/libqadc_ref.so (UNMODIFIED reference sources behind
                 oracle/ref_harness.cpp; built in the CPU container where /root/reference
                 exists, shipped prebuilt to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  Nothing under quick-adc_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libqadc_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libqadc_ref.so")
REF_ADC_SO = os.path.join(HERE, "_ref", "libqadc_ref_adc.so")   # db_query.cpp (plain ADC), own library

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def _opt(arr):
    """ndarray -> void* (or NULL)."""
    return None if arr is None else arr.ctypes.data_as(C.c_void_p)


def build(ref=True):
    """Compile the checkers (called from __graft_entry__.build())."""
    subprocess.check_call(["make", "-s", "-C", HERE, "libqadc_oracle.so"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class Oracle:
    """Plain-C restatement (oracle/qadc_oracle.c)."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = self.lib = C.CDLL(ORACLE_SO)
        L.qo_interleaved_size_4.restype = C.c_long
        L.qo_interleaved_size_4.argtypes = [C.c_uint, C.c_int]
        L.qo_interleave_partition_4.argtypes = [u8p, u8p, C.c_int, C.c_uint]
        L.qo_distances_rowmajor.argtypes = [u8p, C.c_long, C.c_int, i8p, i8p]
        L.qo_distances_interleaved.argtypes = [u8p, C.c_uint, C.c_int, i8p, i8p]
        L.qo_scan_ref_heap.argtypes = [u8p, C.c_void_p, C.c_uint, C.c_int, i8p, u32p, i8p, C.c_int,
                                       C.POINTER(C.c_int)]
        L.qo_heap_push_i8.argtypes = [u32p, i8p, C.c_int, C.POINTER(C.c_int), C.c_uint32, C.c_int8]
        L.qo_heap_push_f32.argtypes = [u32p, f32p, C.c_int, C.POINTER(C.c_int), C.c_uint32, C.c_float]
        L.qo_tables_direct.argtypes = [f32p, C.c_long, C.c_int, C.c_int, f32p, f32p]
        L.qo_tables_blasform.argtypes = [f32p, C.c_long, C.c_int, C.c_int, f32p, f32p]
        L.qo_rotate.argtypes = [f32p, C.c_long, C.c_int, f32p, f32p]
        L.qo_coarse_assign.argtypes = [f32p, C.c_int, C.c_int, f32p, C.c_int, C.c_int, i32p, C.c_void_p]
        L.qo_residuals.argtypes = [f32p, C.c_int, f32p, i32p, C.c_int, f32p]
        L.qo_start_size.restype = C.c_uint
        L.qo_start_size.argtypes = [C.c_uint, C.c_float]
        L.qo_scan_4_heap.argtypes = [u8p, C.c_void_p, C.c_uint, C.c_int, f32p, u32p, f32p, C.c_int,
                                     C.POINTER(C.c_int)]
        L.qo_adc_float_all.argtypes = [u8p, C.c_long, C.c_int, f32p, f32p]
        L.qo_quantize_tables.argtypes = [f32p, C.c_long, C.c_float, C.c_float, i8p]
        L.qo_query_bounds.restype = C.c_int
        L.qo_query_bounds.argtypes = [f32p, C.c_long, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.qo_scan_with_tables.restype = C.c_int
        L.qo_scan_with_tables.argtypes = [u8p, C.c_void_p, i64p, C.c_int, i32p, C.c_int, i8p, C.c_int,
                                          C.c_uint, u32p, i8p, C.c_void_p]
        L.qo_encode.argtypes = [f32p, C.c_long, C.c_int, C.c_int, f32p, u8p]
        L.qo_search.restype = C.c_int
        L.qo_search.argtypes = [C.c_int, C.c_int, f32p, C.c_void_p, C.c_int, C.c_void_p, u8p, C.c_void_p,
                                i64p, C.c_float, f32p, C.c_int, C.c_int, C.c_int, C.c_void_p, u32p, i8p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

    # -- layout -------------------------------------------------------------------------
    def interleave(self, codes):
        n, cs = codes.shape
        out = np.empty(self.lib.qo_interleaved_size_4(n, cs), np.uint8)
        self.lib.qo_interleave_partition_4(out, np.ascontiguousarray(codes), cs, n)
        return out

    # -- int8 distances -----------------------------------------------------------------
    def distances(self, codes, qtab):
        n, cs = codes.shape
        out = np.empty(n, np.int8)
        self.lib.qo_distances_rowmajor(np.ascontiguousarray(codes), n, cs * 2,
                                       np.ascontiguousarray(qtab.reshape(-1)), out)
        return out

    def distances_interleaved(self, part, n, m, qtab):
        out = np.empty(((n + 15) // 16) * 16, np.int8)
        self.lib.qo_distances_interleaved(part, n, m, np.ascontiguousarray(qtab.reshape(-1)), out)
        return out

    def scan_ref_heap(self, part, labels, size, m, qtab, r, heap=None):
        """Emulates scan_avx_4 incl. its heap. heap = (keys, vals, size) to continue one."""
        if heap is None:
            keys = np.zeros(r, np.uint32)
            vals = np.zeros(r, np.int8)
            hs = C.c_int(0)
            self.lib.qo_heap_push_i8(keys, vals, r, C.byref(hs), 0, 127)  # db_query_4.cpp:276
        else:
            keys, vals, n0 = heap
            hs = C.c_int(n0)
        self.lib.qo_scan_ref_heap(part, _opt(labels), size, m, np.ascontiguousarray(qtab.reshape(-1)),
                                  keys, vals, r, C.byref(hs))
        return keys, vals, hs.value

    # -- float stages -------------------------------------------------------------------
    def tables_direct(self, vecs, m, codebooks):
        vecs = np.ascontiguousarray(vecs, np.float32)
        count, dim = vecs.shape
        out = np.empty((count, m, 16), np.float32)
        self.lib.qo_tables_direct(vecs, count, dim, m, np.ascontiguousarray(codebooks.reshape(-1)), out.reshape(-1))
        return out

    def tables_blasform(self, vecs, m, codebooks):
        vecs = np.ascontiguousarray(vecs, np.float32)
        count, dim = vecs.shape
        out = np.empty((count, m, 16), np.float32)
        self.lib.qo_tables_blasform(vecs, count, dim, m, np.ascontiguousarray(codebooks.reshape(-1)), out.reshape(-1))
        return out

    def rotate(self, vecs, rotation):
        vecs = np.ascontiguousarray(vecs, np.float32)
        out = np.empty_like(vecs)
        self.lib.qo_rotate(vecs, vecs.shape[0], vecs.shape[1], np.ascontiguousarray(rotation), out)
        return out

    def coarse_assign(self, queries, centroids, ma):
        queries = np.ascontiguousarray(queries, np.float32)
        nq, dim = queries.shape
        assign = np.empty((nq, ma), np.int32)
        dists = np.empty((nq, ma), np.float32)
        self.lib.qo_coarse_assign(queries, nq, dim, np.ascontiguousarray(centroids), centroids.shape[0], ma,
                                  assign.reshape(-1), _opt(dists))
        return assign, dists

    def encode(self, vectors, m, codebooks, bits=4):
        v = np.ascontiguousarray(vectors, np.float32)
        codes = np.zeros((v.shape[0], m * bits // 8), np.uint8)
        f = self.lib.qo_encode_n
        f.restype = None
        f(_opt(v), C.c_long(v.shape[0]), C.c_int(v.shape[1]), C.c_int(m), C.c_int(bits),
          _opt(np.ascontiguousarray(codebooks, np.float32).reshape(-1)), _opt(codes))
        return codes

    def start_size(self, size, keep):
        return self.lib.qo_start_size(size, np.float32(keep))

    def adc_float(self, codes, table):
        n, cs = codes.shape
        out = np.empty(n, np.float32)
        self.lib.qo_adc_float_all(np.ascontiguousarray(codes), n, cs * 2,
                                  np.ascontiguousarray(table.reshape(-1), np.float32), out)
        return out

    def prefix_qmax(self, prefixes, tables, r):
        """scanner_4::query_scan_start: prefixes = list of row-major code arrays, tables[a]."""
        keys = np.zeros(r, np.uint32)
        vals = np.zeros(r, np.float32)
        hs = C.c_int(0)
        self.lib.qo_heap_push_f32(keys, vals, r, C.byref(hs), 0, np.finfo(np.float32).max)
        for codes, t in zip(prefixes, tables):
            if codes.shape[0] == 0:
                continue
            self.lib.qo_scan_4_heap(np.ascontiguousarray(codes), None, codes.shape[0], codes.shape[1] * 2,
                                    np.ascontiguousarray(t.reshape(-1), np.float32), keys, vals, r, C.byref(hs))
        return float(vals[0])

    def quantize(self, tables, qmin, qmax):
        t = np.ascontiguousarray(tables, np.float32)
        out = np.empty(t.shape, np.int8)
        self.lib.qo_quantize_tables(t.reshape(-1), t.size, qmin, qmax, out.reshape(-1))
        return out

    def scan_with_tables(self, codes, labels, offsets, assign, qtabs, r, pos_base=0):
        m = codes.shape[1] * 2
        ids = np.empty(r, np.uint32)
        d = np.empty(r, np.int8)
        keys = np.empty(r, np.uint64)
        n = self.lib.qo_scan_with_tables(np.ascontiguousarray(codes), _opt(labels),
                                         np.ascontiguousarray(offsets, np.int64), m,
                                         np.ascontiguousarray(assign, np.int32), len(assign),
                                         np.ascontiguousarray(qtabs.reshape(-1)), r, pos_base, ids, d,
                                         _opt(keys))
        return ids, d, n, keys

    def search(self, db, queries, ma, r, assign_in=None, want_tables=True):
        """Full canonical pipeline. db: dict(dim,m,codebooks,rotation,centroids,codes,labels,offsets,keep)."""
        queries = np.ascontiguousarray(queries, np.float32)
        nq = queries.shape[0]
        m = db["m"]
        K = len(db["offsets"]) - 1
        out = dict(ids=np.empty((nq, r), np.uint32), d=np.empty((nq, r), np.int8),
                   count=np.empty(nq, np.int32), assign=np.empty((nq, ma), np.int32),
                   qmin=np.empty(nq, np.float32), qmax=np.empty(nq, np.float32))
        if want_tables:
            out["tables"] = np.empty((nq, ma, m, 16), np.float32)
            out["qtables"] = np.empty((nq, ma, m, 16), np.int8)
        ai = None if assign_in is None else np.ascontiguousarray(assign_in, np.int32)
        rc = self.lib.qo_search(db["dim"], m, np.ascontiguousarray(db["codebooks"].reshape(-1)),
                                _opt(db.get("rotation")), K, _opt(db.get("centroids")),
                                np.ascontiguousarray(db["codes"]), _opt(db.get("labels")),
                                np.ascontiguousarray(db["offsets"], np.int64), np.float32(db["keep"]),
                                queries, nq, ma, r, _opt(ai), out["ids"].reshape(-1), out["d"].reshape(-1),
                                _opt(out["count"]), _opt(out["assign"]), _opt(out.get("tables")),
                                _opt(out["qmin"]), _opt(out["qmax"]), _opt(out.get("qtables")))
        out["rc"] = rc
        return out

    def adc_search(self, db, queries, ma, r):
        """Plain ADC (db_query): db as for search() plus "bits" (4 or 8); float distances."""
        queries = np.ascontiguousarray(queries, np.float32)
        nq = queries.shape[0]
        K = len(db["offsets"]) - 1 if db.get("centroids") is not None else 0
        ids = np.empty((nq, r), np.uint32)
        d = np.empty((nq, r), np.float32)
        cnt = np.empty(nq, np.int32)
        f = self.lib.qo_adc_search
        f.restype = C.c_int
        rc = f(C.c_int(db["dim"]), C.c_int(db["m"]), C.c_int(db.get("bits", 4)),
               _opt(np.ascontiguousarray(db["codebooks"], np.float32).reshape(-1)), _opt(db.get("rotation")), C.c_int(K),
               _opt(db.get("centroids")), _opt(np.ascontiguousarray(db["codes"])), _opt(db.get("labels")),
               _opt(np.ascontiguousarray(db["offsets"], np.int64)), _opt(queries), C.c_int(nq), C.c_int(ma), C.c_int(r),
               _opt(ids), _opt(d), _opt(cnt))
        if rc:
            raise ValueError("qo_adc_search: unsupported configuration")
        return dict(ids=ids, d=d, count=cnt)


class Ref:
    """The unmodified reference behind oracle/ref_harness.cpp."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self):
        L = self.lib = C.CDLL(REF_SO)
        L.ref_blas_single_thread()
        L.ref_interleaved_size_4.restype = C.c_long
        L.ref_interleaved_size_4.argtypes = [C.c_uint, C.c_int]
        L.ref_interleave_partition_4.argtypes = [u8p, u8p, C.c_int, C.c_uint]
        L.ref_scan_avx_4.restype = C.c_int
        L.ref_scan_avx_4.argtypes = [u8p, C.c_void_p, C.c_uint, C.c_int, i8p, C.c_int, C.c_int, u32p, i8p]
        L.ref_dump_distances.restype = C.c_int
        L.ref_dump_distances.argtypes = [u8p, C.c_uint, C.c_int, i8p, i8p]
        L.ref_quantize_tables.argtypes = [f32p, C.c_int, C.c_float, C.c_float, i8p]
        L.ref_tables.restype = C.c_int
        L.ref_tables.argtypes = [f32p, C.c_int, C.c_int, C.c_int, f32p, C.c_int, f32p]
        L.ref_scan_4_qmax.restype = C.c_float
        L.ref_scan_4_qmax.argtypes = [u8p, C.c_uint, C.c_int, f32p, C.c_int]
        L.ref_find_k_neighbors.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, f32p, f32p, i32p]
        L.ref_residuals.argtypes = [f32p, C.c_int, f32p, i32p, C.c_int, f32p]
        L.ref_flat_create.restype = C.c_void_p
        L.ref_flat_create.argtypes = [C.c_int, C.c_int, f32p, u8p, C.c_uint]
        L.ref_ivf_create.restype = C.c_void_p
        L.ref_ivf_create.argtypes = [C.c_int, C.c_int, f32p, C.c_int, f32p, u8p, u32p, i64p]
        L.ref_prepare.argtypes = [C.c_void_p, C.c_float]
        L.ref_starts_size.restype = C.c_uint
        L.ref_starts_size.argtypes = [C.c_void_p, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_query_scan.restype = C.c_int
        L.ref_query_scan.argtypes = [C.c_void_p, i32p, C.c_int, f32p, C.c_int, u32p, i8p]
        L.ref_query_bounds.argtypes = [C.c_void_p, i32p, C.c_int, f32p, C.c_int, C.POINTER(C.c_float),
                                       C.POINTER(C.c_float)]
        L.ref_search.restype = C.c_int
        L.ref_search.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u32p, i8p,
                                 i32p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_encode.argtypes = [C.c_int, C.c_int, f32p, f32p, C.c_int, u8p]
        L.ref_set_rotation.argtypes = [C.c_void_p, f32p]
        L.ref_save_db.restype = C.c_int
        L.ref_save_db.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_load_db.restype = C.c_void_p
        L.ref_load_db.argtypes = [C.c_char_p]
        L.ref_db_info.argtypes = [C.c_void_p, i32p]

    def interleave(self, codes):
        n, cs = codes.shape
        out = np.empty(self.lib.ref_interleaved_size_4(n, cs), np.uint8)
        self.lib.ref_interleave_partition_4(out, np.ascontiguousarray(codes), cs, n)
        return out

    def scan_avx_4(self, part, labels, size, m, qtab, r, sentinel=True):
        keys = np.zeros(r, np.uint32)
        vals = np.zeros(r, np.int8)
        n = self.lib.ref_scan_avx_4(part, _opt(labels), size, m, np.ascontiguousarray(qtab.reshape(-1)), r,
                                    int(sentinel), keys, vals)
        return keys, vals, n

    def dump_distances(self, part, size, m, qtab):
        out = np.empty(size, np.int8)
        rc = self.lib.ref_dump_distances(part, size, m, np.ascontiguousarray(qtab.reshape(-1)), out)
        assert rc == 0
        return out

    def quantize(self, tables, qmin, qmax):
        t = np.ascontiguousarray(tables, np.float32).reshape(-1, 16)
        out = np.empty(t.shape, np.int8)
        self.lib.ref_quantize_tables(t.reshape(-1), t.shape[0], qmin, qmax, out.reshape(-1))
        return out.reshape(np.shape(tables))

    def tables(self, vecs, m, codebooks, blas_form):
        vecs = np.ascontiguousarray(vecs, np.float32)
        count, dim = vecs.shape
        out = np.empty((count, m, 16), np.float32)
        self.lib.ref_tables(vecs, count, dim, m, np.ascontiguousarray(codebooks.reshape(-1)), int(blas_form),
                            out.reshape(-1))
        return out

    def scan_4_qmax(self, codes, table, r):
        return float(self.lib.ref_scan_4_qmax(np.ascontiguousarray(codes), codes.shape[0], codes.shape[1] * 2,
                                              np.ascontiguousarray(table.reshape(-1), np.float32), r))

    def find_k_neighbors(self, vectors, neighbors, k):
        vectors = np.ascontiguousarray(vectors, np.float32)
        out = np.empty((vectors.shape[0], k), np.int32)
        self.lib.ref_find_k_neighbors(vectors.shape[0], neighbors.shape[0], vectors.shape[1], k, vectors,
                                      np.ascontiguousarray(neighbors), out.reshape(-1))
        return out

    def encode(self, vectors, m, codebooks, bits=4):
        v = np.array(vectors, np.float32, order="C", copy=True)
        codes = np.zeros((v.shape[0], m * bits // 8), np.uint8)
        f = self.lib.ref_encode_bits
        f.restype = None
        f(C.c_int(v.shape[1]), C.c_int(m), C.c_int(bits), _opt(np.ascontiguousarray(codebooks, np.float32).reshape(-1)),
          _opt(v), C.c_int(v.shape[0]), _opt(codes))
        return codes

    def index_add_vectors(self, vectors, m, codebooks, centroids, rotation=None):
        """index_db::add_vectors (databases.hpp:270-298), with an opq when `rotation` is given: (assign, codes) per vector."""
        v = np.array(vectors, np.float32, order="C", copy=True)
        n, dim = v.shape
        assign = np.zeros(n, np.int32)
        codes = np.zeros((n, m // 2), np.uint8)
        f = self.lib.ref_index_add_vectors
        f.restype = None
        f(C.c_int(dim), C.c_int(m), _opt(np.ascontiguousarray(codebooks, np.float32).reshape(-1)),
          _opt(None if rotation is None else np.ascontiguousarray(rotation, np.float32).reshape(-1)),
          C.c_int(centroids.shape[0]), _opt(np.ascontiguousarray(centroids, np.float32)), _opt(v), C.c_uint(n),
          _opt(assign), _opt(codes))
        return assign, codes

    class Handle:
        def __init__(self, ref, ptr, m, dim):
            self.ref, self.ptr, self.m, self.dim = ref, ptr, m, dim

        def prepare(self, keep):
            self.ref.lib.ref_prepare(self.ptr, np.float32(keep))

        def starts_size(self, p):
            return self.ref.lib.ref_starts_size(self.ptr, p)

        def query_scan(self, assign, tables, r):
            keys = np.zeros(r, np.uint32)
            vals = np.zeros(r, np.int8)
            t = np.array(tables, np.float32, order="C", copy=True).reshape(-1)
            a = np.ascontiguousarray(assign, np.int32)
            n = self.ref.lib.ref_query_scan(self.ptr, a, len(a), t, r, keys, vals)
            return keys, vals, n

        def query_bounds(self, assign, tables, r):
            t = np.array(tables, np.float32, order="C", copy=True).reshape(-1)
            a = np.ascontiguousarray(assign, np.int32)
            qmin, qmax = C.c_float(), C.c_float()
            self.ref.lib.ref_query_bounds(self.ptr, a, len(a), t, r, C.byref(qmin), C.byref(qmax))
            return qmin.value, qmax.value

        def search(self, queries, ma, r, nthreads=1, blas_tables=False, want_tables=False):
            queries = np.ascontiguousarray(queries, np.float32)
            nq = queries.shape[0]
            keys = np.zeros((nq, r), np.uint32)
            vals = np.zeros((nq, r), np.int8)
            sizes = np.zeros(nq, np.int32)
            assign = np.zeros((nq, ma), np.int32)
            tables = np.zeros((nq, ma, self.m, 16), np.float32) if want_tables else None
            times = np.zeros(4, np.float64)
            self.ref.lib.ref_search(self.ptr, queries, nq, ma, r, nthreads, int(blas_tables),
                                    keys.reshape(-1), vals.reshape(-1), sizes, _opt(assign), _opt(tables),
                                    _opt(times))
            return dict(keys=keys, vals=vals, sizes=sizes, assign=assign, tables=tables, times_us=times)

        def set_rotation(self, rotation):
            """OPQ (quantizers.hpp:248-301); call before prepare()."""
            self.ref.lib.ref_set_rotation(self.ptr, np.ascontiguousarray(rotation, np.float32).reshape(-1))

        def save(self, path):
            """flatdb_create.cpp:49-53 over oracle/shims/cereal."""
            if self.ref.lib.ref_save_db(self.ptr, str(path).encode()) != 0:
                raise IOError("ref_save_db failed")

        def info(self):
            out = np.zeros(6, np.int32)
            self.ref.lib.ref_db_info(self.ptr, out)
            return dict(zip(("index", "opq", "dim", "m", "bits", "partitions"), (int(v) for v in out)))

        def close(self):
            if self.ptr:
                self.ref.lib.ref_destroy(self.ptr)
                self.ptr = None

    def load(self, path):
        """query_common.hpp:321-328 (load_database) over oracle/shims/cereal."""
        p = self.lib.ref_load_db(str(path).encode())
        if not p:
            raise IOError("ref_load_db failed")
        h = Ref.Handle(self, p, 0, 0)
        i = h.info()
        h.m, h.dim = i["m"], i["dim"]
        return h

    def flat(self, dim, m, codebooks, codes):
        p = self.lib.ref_flat_create(dim, m, np.ascontiguousarray(codebooks.reshape(-1)),
                                     np.ascontiguousarray(codes), codes.shape[0])
        return Ref.Handle(self, p, m, dim)

    def ivf(self, dim, m, codebooks, centroids, codes, labels, offsets):
        p = self.lib.ref_ivf_create(dim, m, np.ascontiguousarray(codebooks.reshape(-1)), centroids.shape[0],
                                    np.ascontiguousarray(centroids), np.ascontiguousarray(codes),
                                    np.ascontiguousarray(labels, np.uint32),
                                    np.ascontiguousarray(offsets, np.int64))
        return Ref.Handle(self, p, m, dim)


class RefAdc:
    """The unmodified db_query.cpp path (scanner_simple) behind oracle/ref_harness_adc.cpp."""

    @staticmethod
    def available():
        return os.path.exists(REF_ADC_SO)

    def __init__(self):
        L = self.lib = C.CDLL(REF_ADC_SO)
        L.refadc_create.restype = C.c_void_p
        L.refadc_destroy.argtypes = [C.c_void_p]
        L.refadc_search.restype = C.c_int

    def search(self, db, queries, ma, r):
        """db: dict(dim, m, bits, codebooks, [rotation], [centroids, labels], codes, offsets) -> ids, dists
        (heap sorted by distance; always r entries, unfilled ones are (0, FLT_MAX - t))."""
        queries = np.ascontiguousarray(queries, np.float32)
        nq = queries.shape[0]
        ivf = db.get("centroids") is not None
        K = len(db["offsets"]) - 1 if ivf else 0
        keep = [np.ascontiguousarray(db["codebooks"], np.float32).reshape(-1), np.ascontiguousarray(db["codes"]),
                np.ascontiguousarray(db["offsets"], np.int64)]
        lab = np.ascontiguousarray(db["labels"], np.uint32) if ivf else None
        h = self.lib.refadc_create(C.c_int(db["dim"]), C.c_int(db["m"]), C.c_int(db.get("bits", 4)), _opt(keep[0]),
                                   _opt(db.get("rotation")), C.c_int(K), _opt(db.get("centroids")), _opt(keep[1]),
                                   _opt(lab), _opt(keep[2]))
        ids = np.zeros((nq, r), np.uint32)
        d = np.zeros((nq, r), np.float32)
        self.lib.refadc_search(C.c_void_p(h), _opt(queries), C.c_int(nq), C.c_int(ma), C.c_int(r), _opt(ids), _opt(d))
        self.lib.refadc_destroy(C.c_void_p(h))
        return ids, d
