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
