"""Writers for the files the host CLI (host/db_query_4) reads: the ".qdb" database container
documented in host/databases.hpp, ".pq.data"/".opq.data" quantiser files in the reference's
own format (quantizers.cpp:27-46, convert-quantizer.py:18) and .fvecs/.ivecs vector files."""
import struct

import numpy as np


def write_qdb(path, dim, m, codebooks, codes, rotation=None, centroids=None, labels=None, offsets=None, bits=4):
    cb = np.ascontiguousarray(codebooks, np.float32).reshape(-1)
    assert cb.size == dim * (1 << bits)
    ivf = centroids is not None
    codes = np.ascontiguousarray(codes, np.uint8)
    with open(path, "wb") as f:
        f.write(b"QADCDB1\0")
        K = len(offsets) - 1 if ivf else 1
        f.write(struct.pack("6i", 1 if ivf else 0, 0 if rotation is None else 1, dim, m, bits, K))
        f.write(cb.tobytes())
        if rotation is not None:
            f.write(np.ascontiguousarray(rotation, np.float32).tobytes())
        if ivf:
            f.write(np.ascontiguousarray(centroids, np.float32).tobytes())
            offsets = np.asarray(offsets, np.int64)
            f.write(np.diff(offsets).astype(np.uint64).tobytes())
            lab = np.ascontiguousarray(labels, np.uint32)
            for p in range(K):
                f.write(codes[offsets[p]:offsets[p + 1]].tobytes())
                f.write(lab[offsets[p]:offsets[p + 1]].tobytes())
        else:
            f.write(np.array([codes.shape[0]], np.uint64).tobytes())
            f.write(codes.tobytes())


def write_pq_data(path, dim, m, codebooks, rotation=None, bits=4):
    """int32 dim, m, bits; float codebooks[dim * 2^bits]; (.opq.data) float rotation[dim*dim]."""
    with open(path, "wb") as f:
        f.write(struct.pack("iii", dim, m, bits))
        f.write(np.ascontiguousarray(codebooks, np.float32).tobytes())
        if rotation is not None:
            f.write(np.ascontiguousarray(rotation, np.float32).tobytes())


def write_vecs(path, array):
    """.fvecs (float32) / .ivecs (int32): int32 dimension before every vector."""
    a = np.ascontiguousarray(array)
    assert a.dtype in (np.float32, np.int32)
    n, d = a.shape
    out = np.empty((n, d + 1), np.int32)
    out[:, 0] = d
    out[:, 1:] = a.view(np.int32)
    out.tofile(path)


# ---- the reference's own database files -----------------------------------------------------
# cereal 1.2.2 BinaryOutputArchive of a std::unique_ptr<base_db> (flatdb_create.cpp:49-53); the
# byte layout is documented in host/databases.hpp ("cereal archive").
_FIRST_USE, _SAME_TYPE = 0x80000000, 0x40000000


def _named(f, type_id, name):
    f.write(struct.pack("<IQ", _FIRST_USE | type_id, len(name)) + name + b"\x01")


def _archive_pq(f, dim, m, bits, cb, rotation):
    if rotation is None:
        f.write(struct.pack("<IB", _SAME_TYPE, 1))
    else:
        _named(f, 2, b"opq")
    f.write(struct.pack("<3i", m, bits, dim))
    f.write(cb.tobytes())
    if rotation is not None:
        f.write(np.ascontiguousarray(rotation, np.float32).tobytes())


def write_archive_db(path, dim, m, codebooks, codes, rotation=None, centroids=None, labels=None, offsets=None, bits=4):
    """Same arguments as write_qdb; writes the layout the reference's tools read and write."""
    cb = np.ascontiguousarray(codebooks, np.float32).reshape(-1)
    assert cb.size == dim * (1 << bits)
    codes = np.ascontiguousarray(codes, np.uint8)
    with open(path, "wb") as f:
        if centroids is None:
            _named(f, 1, b"flat_db")
            _archive_pq(f, dim, m, bits, cb, rotation)
            f.write(struct.pack("<IQ", codes.shape[0], codes.size))
            f.write(codes.tobytes())
            return
        offsets = np.asarray(offsets, np.int64)
        K = len(offsets) - 1
        lab = np.ascontiguousarray(labels, np.uint32)
        _named(f, 1, b"index_db")
        f.write(struct.pack("<i", K))
        _archive_pq(f, dim, m, bits, cb, rotation)
        f.write(np.ascontiguousarray(centroids, np.float32).tobytes())
        for p in range(K):
            part = codes[offsets[p]:offsets[p + 1]]
            f.write(struct.pack("<Q", part.size) + part.tobytes())
        for p in range(K):
            part = lab[offsets[p]:offsets[p + 1]]
            f.write(struct.pack("<Q", part.size) + part.tobytes())


def read_db(path):
    """Reads either format into the keyword arguments of write_qdb / write_archive_db."""
    raw = np.fromfile(path, np.uint8)
    pos = [0]

    def take(dtype, count=1):
        a = raw[pos[0]:pos[0] + np.dtype(dtype).itemsize * count].view(dtype)
        if a.size != count:
            raise ValueError("truncated database file")
        pos[0] += a.nbytes
        return a

    out = {}
    if raw[:8].tobytes() == b"QADCDB1\0":
        pos[0] = 8
        kind, pq_kind, dim, m, bits, K = (int(v) for v in take("<i4", 6))
        out.update(dim=dim, m=m, bits=bits, codebooks=take("<f4", dim << bits).reshape(m, 1 << bits, dim // m).copy())
        if pq_kind:
            out["rotation"] = take("<f4", dim * dim).reshape(dim, dim).copy()
        cs = m * bits // 8
        if kind == 0:
            n = int(take("<u8")[0])
            out["codes"] = take(np.uint8, n * cs).reshape(n, cs).copy()
            return out
        out["centroids"] = take("<f4", K * dim).reshape(K, dim).copy()
        sizes = take("<u8", K).astype(np.int64)
        codes, labels = [], []
        for p in range(K):
            codes.append(take(np.uint8, int(sizes[p]) * cs).reshape(-1, cs))
            labels.append(take("<u4", int(sizes[p])))
    else:
        def named():
            tid = int(take("<u4")[0])
            if not tid & _FIRST_USE:
                raise ValueError("not a database file")
            name = take(np.uint8, int(take("<u8")[0])).tobytes()
            if int(take(np.uint8)[0]) != 1:
                raise ValueError("null pointer in archive")
            return name

        def pq():
            tid = int(take("<u4")[0])
            pos[0] -= 4
            is_opq = bool(tid & _FIRST_USE)
            if is_opq:
                if named() != b"opq":
                    raise ValueError("unknown quantizer type")
            else:
                take("<u4")
                take(np.uint8)
            m, bits, dim = (int(v) for v in take("<i4", 3))
            out.update(dim=dim, m=m, bits=bits, codebooks=take("<f4", dim << bits).reshape(m, 1 << bits, dim // m).copy())
            if is_opq:
                out["rotation"] = take("<f4", dim * dim).reshape(dim, dim).copy()

        kind = named()
        if kind == b"flat_db":
            pq()
            n = int(take("<u4")[0])
            nbytes = int(take("<u8")[0])
            out["codes"] = take(np.uint8, nbytes).reshape(n, -1).copy()
            return out
        if kind != b"index_db":
            raise ValueError("unknown database type %r" % kind)
        K = int(take("<i4")[0])
        pq()
        cs = out["m"] * out["bits"] // 8
        out["centroids"] = take("<f4", K * out["dim"]).reshape(K, out["dim"]).copy()
        codes = [take(np.uint8, int(take("<u8")[0])).reshape(-1, cs) for _ in range(K)]
        labels = [take("<u4", int(take("<u8")[0])) for _ in range(K)]
    out["codes"] = np.concatenate(codes) if codes else np.zeros((0, cs), np.uint8)
    out["labels"] = np.concatenate(labels) if labels else np.zeros(0, np.uint32)
    out["offsets"] = np.concatenate([[0], np.cumsum([len(l) for l in labels])]).astype(np.int64)
    return out
