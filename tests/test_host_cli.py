"""The C++ host mirror (quick-adc_b200/host): the reference's db_query_4 command line on top of
the C ABI.  CPU: it builds, links the library and rejects bad input like the reference does.
GPU: its results equal the oracle's canonical results and its CSV has the reference's columns."""
import os
import subprocess

import numpy as np
import pytest

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "quick-adc_b200", "host")
CLI = os.path.join(HOST, "db_query_4")


@pytest.fixture(scope="module")
def cli(qadc):
    qadc.load_library()
    subprocess.check_call(["make", "-s", "-C", HOST])
    return CLI


def test_cli_builds_and_prints_usage(cli):
    p = subprocess.run([cli], capture_output=True, text=True)
    assert p.returncode == 1 and "Usage: db_query_4 [-r R] [-m MA] [-k KEEP_PERCENT] [-b BATCH_SIZE]" in p.stderr


def test_cli_rejects_non_database(cli, tmp_path):
    f = tmp_path / "x.qdb"
    f.write_bytes(b"not a database at all, just bytes" * 4)
    p = subprocess.run([cli, str(f), str(f), str(f)], capture_output=True, text=True)
    assert p.returncode == 1 and "is not a database file" in p.stderr


def test_cli_rejects_corrupt_qdb_headers(cli, tmp_path):
    """Every header field of a .qdb file is validated before it sizes an allocation (m = 0, absurd sizes ...)."""
    import struct
    good = struct.pack("8s6i", b"QADCDB1\0", 0, 0, 64, 16, 4, 1)
    for bad in (struct.pack("8s6i", b"QADCDB1\0", 0, 0, 64, 0, 4, 1),            # m = 0
                struct.pack("8s6i", b"QADCDB1\0", 0, 0, 64, 16, 31, 1),           # huge bits
                struct.pack("8s6i", b"QADCDB1\0", 1, 0, 64, 16, 4, 1 << 30),      # absurd partition count
                good + b"\0" * (64 * 16 * 4) + struct.pack("Q", 1 << 40)):         # absurd vector count
        f = tmp_path / "bad.qdb"
        f.write_bytes(bad)
        p = subprocess.run([cli, str(f), str(f), str(f)], capture_output=True, text=True)
        assert p.returncode == 1 and "corrupt .qdb" in p.stderr, p.stderr


# ---- database files: the reference's archive layout and the .qdb container --------------------
ARCHIVE_KINDS = [(ivf, opq) for ivf in (False, True) for opq in (False, True)]


def archive_kwargs(g, ivf, opq):
    kw = dict(dim=int(g["dim"]), m=int(g["m"]), codebooks=g["codebooks"], codes=g["ivf_codes"] if ivf else g["codes"],
              rotation=g["rotation"] if opq else None)
    if ivf:
        kw.update(centroids=g["centroids"], labels=g["labels"], offsets=g["offsets"])
    return kw


def archive_key(ivf, opq):
    return "ref_%s_%s" % ("index" if ivf else "flat", "opq" if opq else "pq")


@pytest.mark.parametrize("ivf,opq", ARCHIVE_KINDS)
def test_archive_writer_matches_reference_files(qadc, tmp_path, ivf, opq):
    """dbfile.write_archive_db == the bytes the reference's save() members produced
    (tests/golden/archives.npz), and read_db gets the arrays back."""
    from qadc_b200 import dbfile
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "archives.npz")))
    kw = archive_kwargs(g, ivf, opq)
    dbfile.write_archive_db(tmp_path / "a.db", **kw)
    assert (tmp_path / "a.db").read_bytes() == g[archive_key(ivf, opq)].tobytes()
    back = dbfile.read_db(tmp_path / "a.db")
    for k, v in kw.items():
        if v is None:
            assert k not in back
        else:
            assert np.array_equal(back[k], v), k


@pytest.mark.parametrize("ivf,opq", ARCHIVE_KINDS)
def test_db_convert_reads_and_writes_reference_files(cli, qadc, tmp_path, ivf, opq):
    """C++ host loader/saver (host/databases.hpp): reference archive -> .qdb -> archive, byte exact."""
    from qadc_b200 import dbfile
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "archives.npz")))
    ref_bytes = g[archive_key(ivf, opq)].tobytes()
    (tmp_path / "ref.db").write_bytes(ref_bytes)
    conv = os.path.join(HOST, "db_convert")
    p = subprocess.run([conv, str(tmp_path / "ref.db"), str(tmp_path / "x.qdb")], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert ("Indexed DB (partitions=6)" if ivf else "Flat DB") in p.stderr and ("opq" if opq else "pq") + " (dim=32" in p.stderr
    assert "Vectors: 300" in p.stderr
    dbfile.write_qdb(tmp_path / "y.qdb", **archive_kwargs(g, ivf, opq))
    assert (tmp_path / "x.qdb").read_bytes() == (tmp_path / "y.qdb").read_bytes()
    p = subprocess.run([conv, str(tmp_path / "x.qdb"), str(tmp_path / "back.db")], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert (tmp_path / "back.db").read_bytes() == ref_bytes
    # a truncated archive is rejected, not half-read
    (tmp_path / "cut.db").write_bytes(ref_bytes[:len(ref_bytes) // 2])
    p = subprocess.run([conv, str(tmp_path / "cut.db"), str(tmp_path / "z.qdb")], capture_output=True, text=True)
    assert p.returncode == 1 and "is not a database file" in p.stderr


@pytest.mark.parametrize("ivf,opq", ARCHIVE_KINDS)
def test_reference_loads_our_files(ref, qadc, tmp_path, ivf, opq):
    """The reference's load_database (query_common.hpp:321-328) reads a file written by this repo
    and searches it exactly like the database it was built from in memory."""
    from qadc_b200 import dbfile
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "archives.npz")))
    kw = archive_kwargs(g, ivf, opq)
    dbfile.write_archive_db(tmp_path / "ours.db", **kw)
    loaded = ref.load(tmp_path / "ours.db")
    assert loaded.info() == dict(index=int(ivf), opq=int(opq), dim=32, m=16, bits=4, partitions=6 if ivf else 1)
    if ivf:
        mem = ref.ivf(kw["dim"], kw["m"], kw["codebooks"], kw["centroids"], kw["codes"], kw["labels"], kw["offsets"])
    else:
        mem = ref.flat(kw["dim"], kw["m"], kw["codebooks"], kw["codes"])
    if opq:
        mem.set_rotation(kw["rotation"])
    # the live reference writes the same bytes as the committed fixture
    mem.save(tmp_path / "theirs.db")
    assert (tmp_path / "theirs.db").read_bytes() == g[archive_key(ivf, opq)].tobytes()
    q = synth.make_queries(np.random.default_rng(3), 5, kw["dim"])
    res = []
    for h in (loaded, mem):
        h.prepare(0.2)
        res.append(h.search(q, 3 if ivf else 1, 10, nthreads=1, blas_tables=False))
        h.close()
    for k in ("keys", "vals", "sizes"):
        assert np.array_equal(res[0][k], res[1][k])


def test_pq_data_round_trip(qadc, tmp_path):
    """.pq.data/.opq.data in the reference's format (quantizers.cpp:27-46)."""
    from qadc_b200 import dbfile
    import struct
    rng = np.random.default_rng(0)
    cb = synth.make_pq(rng, 64, 16)
    rot = rng.standard_normal((64, 64)).astype(np.float32)
    path = tmp_path / "q.opq.data"
    dbfile.write_pq_data(path, 64, 16, cb, rot)
    raw = path.read_bytes()
    assert struct.unpack("iii", raw[:12]) == (64, 16, 4)
    assert len(raw) == 12 + 4 * (64 * 16 + 64 * 64)
    assert np.array_equal(np.frombuffer(raw[12:12 + 4 * 64 * 16], np.float32), cb.reshape(-1))


def cli_keep(percent):
    """`-k PERCENT` as db_query_4.cpp:342 computes it: atof(optarg) * ONE_PERCENT (double x float(0.01)) -> float."""
    return np.float32(np.float64(percent) * np.float64(np.float32(0.01)))


def run_cli(cli, tmp_path, db_kwargs, queries, gt, r, ma, keep_percent, batch, archive=False, gpus=None):
    from qadc_b200 import dbfile
    (dbfile.write_archive_db if archive else dbfile.write_qdb)(tmp_path / "db.qdb", **db_kwargs)
    dbfile.write_vecs(tmp_path / "q.fvecs", queries)
    dbfile.write_vecs(tmp_path / "gt.ivecs", gt)
    out = tmp_path / "res.bin"
    p = subprocess.run([cli, "-r", str(r), "-m", str(ma), "-k", str(keep_percent), "-b", str(batch), "-o", str(out)]
                       + (["-g", gpus] if gpus else []) +
                       [str(tmp_path / "db.qdb"), str(tmp_path / "q.fvecs"), str(tmp_path / "gt.ivecs")],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    lines = p.stdout.strip().splitlines()
    assert lines[0] == "r,recall,ma,adc_type,keep,index_us,rotate_us,table_us,scan_us"   # db_query_4.cpp:387
    fields = lines[1].split(",")
    raw = np.fromfile(out, np.uint8).reshape(queries.shape[0], r * 5)
    ids = raw[:, :4 * r].copy().view(np.uint32)
    d = raw[:, 4 * r:].view(np.int8)
    return fields, ids, d


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [1, 7])
def test_cli_flat_matches_oracle(cli, oracle, tmp_path, batch):
    rng = np.random.default_rng(31)
    dim, m, n, nq, r = 128, 16, 30000, 20, 50
    cb = synth.make_pq(rng, dim, m)
    codes = synth.make_codes(rng, n, m)
    q = synth.make_queries(rng, nq, dim)
    exp = oracle.search(dict(dim=dim, m=m, codebooks=cb, codes=codes, keep=cli_keep(2), offsets=np.array([0, n], np.int64)),
                        q, 1, r, want_tables=False)
    gt = exp["ids"][:, :1].astype(np.int32).copy()
    gt[::2] = n + 5   # half of the queries cannot be recalled
    fields, ids, d = run_cli(cli, tmp_path, dict(dim=dim, m=m, codebooks=cb, codes=codes), q, gt, r, 1, 2, batch)
    assert fields[0] == str(r) and fields[2] == "1" and fields[3] == "qadc"
    assert abs(float(fields[1]) - 0.5) < 1e-9
    assert np.array_equal(d, exp["d"])
    # the CLI sorts by distance only (like kv_binheap::sort): compare ids as sets per distance
    for qi in range(nq):
        for v in np.unique(d[qi]):
            assert set(ids[qi][d[qi] == v].tolist()) == set(exp["ids"][qi][exp["d"][qi] == v].tolist())


@pytest.mark.gpu
@pytest.mark.parametrize("archive", [False, True])
def test_cli_ivf_matches_oracle(cli, oracle, tmp_path, archive):
    """archive=True: the database is an OPQ index_db file in the reference's own layout."""
    rng = np.random.default_rng(32)
    dim, m, n, K, ma, nq, r = 96, 32, 20000, 40, 6, 9, 30
    cb = synth.make_pq(rng, dim, m)
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    codes, labels, offsets = synth.make_ivf(rng, n, K, m, empty=(7,))
    q = synth.make_queries(rng, nq, dim)
    rot = np.linalg.qr(rng.standard_normal((dim, dim)))[0].astype(np.float32) if archive else None
    db = dict(dim=dim, m=m, codebooks=cb, centroids=cents, codes=codes, labels=labels, keep=cli_keep(5), offsets=offsets)
    if archive:
        db["rotation"] = rot
    exp = oracle.search(db, q, ma, r, want_tables=False)
    gt = exp["ids"][:, :1].astype(np.int32).copy()
    fields, ids, d = run_cli(cli, tmp_path, dict(dim=dim, m=m, codebooks=cb, codes=codes, centroids=cents, labels=labels,
                                                 offsets=offsets, rotation=rot), q, gt, r, ma, 5, 4, archive=archive)
    assert float(fields[1]) == 1.0 and fields[2] == str(ma)
    assert np.array_equal(d, exp["d"])
    for qi in range(nq):
        for v in np.unique(d[qi]):
            assert set(ids[qi][d[qi] == v].tolist()) == set(exp["ids"][qi][exp["d"][qi] == v].tolist())


@pytest.mark.gpu
@pytest.mark.parametrize("ivf,use_opq", [(False, False), (True, False), (True, True)])
def test_db_build_then_query(cli, oracle, tmp_path, ivf, use_opq):
    """floats -> db_build (GPU encoder, the reference's flatdb_create/db_add chain) -> db_query_4:
    same results as the oracle's encoder + search on the same inputs.  The inverted-list case goes
    through a file in the reference's archive layout, the flat one through a .qdb container."""
    from qadc_b200 import dbfile
    rng = np.random.default_rng(33 + ivf)
    dim, m, n, nq, r, K, ma = 64, 16, 12000, 10, 20, 20, 4
    dbname = str(tmp_path / ("db.index" if ivf else "db.qdb"))
    cb = synth.make_pq(rng, dim, m)
    base = rng.standard_normal((n, dim)).astype(np.float32)
    q = synth.make_queries(rng, nq, dim)
    rot = np.linalg.qr(rng.standard_normal((dim, dim)))[0].astype(np.float32) if use_opq else None
    qfile = tmp_path / ("q.opq.data" if use_opq else "q.pq.data")     # OPQ + inverted lists: the README's flagship index
    dbfile.write_pq_data(qfile, dim, m, cb, rot)
    dbfile.write_vecs(tmp_path / "base.fvecs", base)
    dbfile.write_vecs(tmp_path / "q.fvecs", q)
    cmd = [os.path.join(HOST, "db_build")]
    if ivf:
        cents = base[rng.permutation(n)[:K]].copy()
        dbfile.write_vecs(tmp_path / "cents.fvecs", cents)
        cmd += ["-c", str(tmp_path / "cents.fvecs")]
    cmd += [str(qfile), str(tmp_path / "base.fvecs"), dbname]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert ("Indexed DB (partitions=20)" if ivf else "Flat DB") in p.stderr and ("opq" if use_opq else "pq") + " (dim=64, sq=16x4)" in p.stderr
    # expected database built with the oracle: assignment, residual codes, insertion order per cell
    keep = cli_keep(10)
    if ivf:
        assign, _ = oracle.coarse_assign(base, cents, 1)
        assign = assign[:, 0]
        resid = (base - cents[assign]).astype(np.float32)
        codes_all = oracle.encode(oracle.rotate(resid, rot) if use_opq else resid, m, cb)
        order = np.argsort(assign, kind="stable")
        offsets = np.zeros(K + 1, np.int64)
        offsets[1:] = np.cumsum(np.bincount(assign, minlength=K))
        db = dict(dim=dim, m=m, codebooks=cb, centroids=cents, codes=codes_all[order], labels=order.astype(np.uint32),
                  keep=keep, offsets=offsets)
        if use_opq:
            db["rotation"] = rot
        exp = oracle.search(db, q, ma, r, want_tables=False)
    else:
        db = dict(dim=dim, m=m, codebooks=cb, codes=oracle.encode(base, m, cb), keep=keep, offsets=np.array([0, n], np.int64))
        exp = oracle.search(db, q, 1, r, want_tables=False)
    gt = exp["ids"][:, :1].astype(np.int32).copy()
    dbfile.write_vecs(tmp_path / "gt.ivecs", gt)
    out = tmp_path / "res.bin"
    p = subprocess.run([cli, "-r", str(r), "-m", str(ma if ivf else 1), "-k", "10", "-b", "0", "-o", str(out),
                        dbname, str(tmp_path / "q.fvecs"), str(tmp_path / "gt.ivecs")],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert float(p.stdout.strip().splitlines()[1].split(",")[1]) == 1.0
    raw = np.fromfile(out, np.uint8).reshape(nq, r * 5)
    ids, d = raw[:, :4 * r].copy().view(np.uint32), raw[:, 4 * r:].view(np.int8)
    assert np.array_equal(d, exp["d"])
    for qi in range(nq):
        for v in np.unique(d[qi]):
            assert set(ids[qi][d[qi] == v].tolist()) == set(exp["ids"][qi][exp["d"][qi] == v].tolist())


def _device_list():
    """Every GPU of the box when there are several (NCCL path), else three virtual shards on GPU 0."""
    import torch
    n = torch.cuda.device_count()
    return ",".join(str(i) for i in range(n)) if n > 1 else "0,0,0"


@pytest.mark.gpu
@pytest.mark.parametrize("ivf", [False, True])
def test_cli_sharded_over_gpus_equals_one_gpu(cli, tmp_path, ivf):
    """db_query_4 -g 0,1,...: the database sharded over the GPUs by the one CLI process (qadc_multi_*: NCCL
    all-gather of the per-GPU top-r lists + merge) prints the same CSV fields and dumps byte-identical
    results as the one-GPU run."""
    rng = np.random.default_rng(41 + ivf)
    dim, m, nq, r = 128, 16, 37, 100
    cb = synth.make_pq(rng, dim, m)
    q = synth.make_queries(rng, nq, dim)
    if ivf:
        n, K, ma = 60000, 96, 12
        cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
        codes, labels, offsets = synth.make_ivf(rng, n, K, m, empty=(3, 50))
        kw = dict(dim=dim, m=m, codebooks=cb, codes=codes, centroids=cents, labels=labels, offsets=offsets)
    else:
        n, ma = 150000, 1
        kw = dict(dim=dim, m=m, codebooks=cb, codes=synth.make_codes(rng, n, m))
    gt = rng.integers(0, n, (nq, 1)).astype(np.int32)
    (tmp_path / "one").mkdir(); (tmp_path / "many").mkdir()
    f1, ids1, d1 = run_cli(cli, tmp_path / "one", kw, q, gt, r, ma, 5, 16)
    f2, ids2, d2 = run_cli(cli, tmp_path / "many", kw, q, gt, r, ma, 5, 16, gpus=_device_list())
    assert f1[:5] == f2[:5]                      # r, recall, ma, adc_type, keep
    assert np.array_equal(d1, d2) and np.array_equal(ids1, ids2)


# ---- db_query: the plain ADC tool ("next" row N4) -------------------------------------------------
def test_db_query_builds_and_prints_usage(cli):
    p = subprocess.run([os.path.join(HOST, "db_query")], capture_output=True, text=True)
    assert p.returncode == 1 and "Usage: db_query: [-r R] [-m MA] [-b BATCH_SIZE]" in p.stderr   # db_query.cpp:48-52


@pytest.mark.gpu
@pytest.mark.parametrize("m,bits,ivf", [(8, 8, False), (16, 8, True), (16, 4, True)])
def test_db_query_matches_oracle(cli, oracle, tmp_path, m, bits, ivf):
    """db_query on a database file in the reference's archive layout: same distances as the oracle's
    plain ADC (bitwise), same ids per distance, recall column, CSV header of db_query.cpp:116-119."""
    from qadc_b200 import dbfile
    rng = np.random.default_rng(60 + m + bits)
    dim, n, nq, r, K, ma = 8 * m, 20000, 9, 30, 12, 4
    cb = rng.standard_normal((m, 1 << bits, dim // m)).astype(np.float32)
    codes = rng.integers(0, 256, (n, m * bits // 8), dtype=np.uint8)
    q = synth.make_queries(rng, nq, dim)
    db = dict(dim=dim, m=m, bits=bits, codebooks=cb, codes=codes, offsets=np.array([0, n], np.int64))
    kw = dict(dim=dim, m=m, codebooks=cb, codes=codes, bits=bits)
    if ivf:
        sizes = rng.multinomial(n, np.ones(K) / K)
        sizes[5] = 0
        db.update(offsets=np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64),
                  centroids=(2 * rng.standard_normal((K, dim))).astype(np.float32), labels=rng.permutation(n).astype(np.uint32))
        kw.update(centroids=db["centroids"], labels=db["labels"], offsets=db["offsets"])
    else:
        ma = 1
    exp = oracle.adc_search(db, q, ma, r)
    gt = exp["ids"][:, :1].astype(np.int32).copy()
    gt[1::2] = n + 7                                          # every other query cannot be recalled
    dbfile.write_archive_db(tmp_path / "db.bin", **kw)
    dbfile.write_vecs(tmp_path / "q.fvecs", q)
    dbfile.write_vecs(tmp_path / "gt.ivecs", gt)
    out = tmp_path / "res.bin"
    p = subprocess.run([os.path.join(HOST, "db_query"), "-r", str(r), "-m", str(ma), "-b", "4", "-o", str(out),
                        str(tmp_path / "db.bin"), str(tmp_path / "q.fvecs"), str(tmp_path / "gt.ivecs")],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    lines = p.stdout.strip().splitlines()
    assert lines[0] == "r,recall,ma,adc_type,index_us,rotate_us,table_us,scan_us"
    f = lines[1].split(",")
    assert f[0] == str(r) and f[2] == str(ma) and f[3] == "adc" and abs(float(f[1]) - 5 / 9) < 1e-6
    raw = np.fromfile(out, np.uint8).reshape(nq, r * 8)
    ids = raw[:, :4 * r].copy().view(np.uint32)
    d = raw[:, 4 * r:].copy().view(np.float32)
    assert np.array_equal(d.view(np.uint32), exp["d"].view(np.uint32))
    for qi in range(nq):
        for v in np.unique(d[qi]):
            assert set(ids[qi][d[qi] == v].tolist()) == set(exp["ids"][qi][exp["d"][qi] == v].tolist())


def test_file_formats_random_shapes(cli, qadc, tmp_path):
    """Python writer/reader against the C++ loader/saver over random small databases: flat and
    inverted lists, PQ and OPQ, 4- and 8-bit codes, empty partitions, both formats, both directions."""
    from qadc_b200 import dbfile
    rng = np.random.default_rng(2024)
    conv = os.path.join(HOST, "db_convert")
    for case in range(24):
        bits = int(rng.choice([4, 8]))
        m = int(rng.choice([16, 32] if bits == 4 else [4, 8, 16]))
        dim = m * int(rng.integers(1, 5))
        n = int(rng.integers(0, 400))
        ivf, opq = bool(case & 1), bool(case & 2)
        kw = dict(dim=dim, m=m, bits=bits, codebooks=rng.standard_normal((m, 1 << bits, dim // m)).astype(np.float32),
                  codes=rng.integers(0, 256, (n, m * bits // 8), dtype=np.uint8),
                  rotation=rng.standard_normal((dim, dim)).astype(np.float32) if opq else None)
        if ivf:
            K = int(rng.integers(1, 9))
            sizes = rng.multinomial(n, np.ones(K) / K)
            kw.update(centroids=rng.standard_normal((K, dim)).astype(np.float32), labels=rng.permutation(n).astype(np.uint32),
                      offsets=np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64))
        first, second = ("a.db", "b.qdb") if case & 4 else ("a.qdb", "b.db")
        (dbfile.write_qdb if first.endswith(".qdb") else dbfile.write_archive_db)(tmp_path / first, **kw)
        p = subprocess.run([conv, str(tmp_path / first), str(tmp_path / second)], capture_output=True, text=True)
        assert p.returncode == 0, (case, p.stderr)
        assert "Vectors: %d" % n in p.stderr
        (dbfile.write_qdb if second.endswith(".qdb") else dbfile.write_archive_db)(tmp_path / "expect", **kw)
        assert (tmp_path / second).read_bytes() == (tmp_path / "expect").read_bytes(), case
        back = dbfile.read_db(tmp_path / second)
        for k, v in kw.items():
            if v is None:
                assert k not in back
            elif k in ("dim", "m", "bits"):
                assert back[k] == v
            else:
                assert np.array_equal(back[k], v), (case, k)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,m,bits,n", [(64, 8, 8, 15000), (8, 4, 16, 3000)])
def test_db_build_8bit_then_db_query(cli, oracle, tmp_path, dim, m, bits, n):
    """8- and 16-bit quantisers: floats -> db_build (GPU encoder; flat database in the reference's archive layout)
    -> db_query == the oracle's encoder + plain ADC on the same inputs."""
    from qadc_b200 import dbfile
    rng = np.random.default_rng(71)
    nq, r = 8, 20
    if bits == 16:
        cb = synth.lattice_codebook16(m)
        base = (3.0 * rng.standard_normal((n, dim))).astype(np.float32)
        q = (3.0 * synth.make_queries(rng, nq, dim)).astype(np.float32)
    else:
        cb = rng.standard_normal((m, 256, dim // m)).astype(np.float32)
        base = rng.standard_normal((n, dim)).astype(np.float32)
        q = synth.make_queries(rng, nq, dim)
    dbfile.write_pq_data(tmp_path / "q.pq.data", dim, m, cb, bits=bits)
    dbfile.write_vecs(tmp_path / "base.fvecs", base)
    dbfile.write_vecs(tmp_path / "q.fvecs", q)
    p = subprocess.run([os.path.join(HOST, "db_build"), str(tmp_path / "q.pq.data"), str(tmp_path / "base.fvecs"),
                        str(tmp_path / "db.flat")], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert f"pq (dim={dim}, sq={m}x{bits})" in p.stderr
    codes = oracle.encode(base, m, cb, bits)
    assert np.array_equal(dbfile.read_db(tmp_path / "db.flat")["codes"], codes)
    exp = oracle.adc_search(dict(dim=dim, m=m, bits=bits, codebooks=cb, codes=codes, offsets=np.array([0, n], np.int64)), q, 1, r)
    dbfile.write_vecs(tmp_path / "gt.ivecs", exp["ids"][:, :1].astype(np.int32).copy())
    out = tmp_path / "res.bin"
    p = subprocess.run([os.path.join(HOST, "db_query"), "-r", str(r), "-o", str(out), str(tmp_path / "db.flat"),
                        str(tmp_path / "q.fvecs"), str(tmp_path / "gt.ivecs")], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert float(p.stdout.strip().splitlines()[1].split(",")[1]) == 1.0
    raw = np.fromfile(out, np.uint8).reshape(nq, r * 8)
    d = raw[:, 4 * r:].copy().view(np.float32)
    assert np.array_equal(d.view(np.uint32), exp["d"].view(np.uint32))
