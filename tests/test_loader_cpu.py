"""MatrixMarket ingest: mpg_mm_read_host against the CSR the reference's own LoadMatrix<double>() built
from the same files (tests/golden/mm_expected.json, made by tests/golden/make_mm_goldens.py through oracle/_ref), bit for
bit — explicit zero diagonals, symmetric mirroring, unmerged duplicates in stable order — and the reference's error text."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EXP = json.load(open(os.path.join(HERE, "golden", "mm_expected.json")))


@pytest.mark.parametrize("name", sorted(EXP))
def test_loader_matches_reference_loadmatrix(g, name):
    rm, ind, val = g.read_matrix_market(os.path.join(HERE, "golden", "mm", name))
    e = EXP[name]
    np.testing.assert_array_equal(rm, np.array(e["row_map"], np.int32))
    np.testing.assert_array_equal(ind, np.array(e["inds"], np.int32))
    np.testing.assert_array_equal(val, np.array([float.fromhex(v) for v in e["vals"]]))
    n = len(rm) - 1
    for r in range(n):   # canonical form: diagonal present, ascending columns (duplicates allowed)
        cols = ind[rm[r]:rm[r + 1]]
        assert r in cols and np.all(np.diff(cols) >= 0)


def test_loader_round_trips_generated_matrix(g, orc):
    rm, ind, val = orc.gen("cd27:4")
    a = g.read_matrix_market(os.path.join(HERE, "golden", "mm", "cd27_4_shuffled.mtx"))
    np.testing.assert_array_equal(a[0], rm); np.testing.assert_array_equal(a[1], ind); np.testing.assert_array_equal(a[2], val)


@pytest.mark.parametrize("content,msg", [
    ("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n", "Unsupported matrix type"),        # LoadMatrix.hpp:48-54
    ("%MatrixMarket matrix coordinate real general\n1 1 1\n1 1 1\n", "Banner is missing"),             # :34-35
    ("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1 0\n", "Unsupported matrix type"),
    ("%%MatrixMarket matrix coordinate pattern general\n1 1 1\n1 1\n", "Unsupported matrix type"),
    ("%%MatrixMarket matrix coordinate real general\n", "Malformed matrix size information"),           # :43-46
])
def test_loader_errors_like_the_reference(g, tmp_path, content, msg):
    p = tmp_path / "bad.mtx"
    p.write_text(content)
    with pytest.raises(g.MpgError, match=msg):
        g.read_matrix_market(p)
    with pytest.raises(g.MpgError, match="Could not access file"):                                      # :22-25
        g.read_matrix_market(tmp_path / "missing.mtx")


def test_live_reference_loader_agrees(g):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
    import oracle_ref
    if not oracle_ref.available():
        pytest.skip("oracle/_ref not built")
    for name in sorted(EXP):
        a = g.read_matrix_market(os.path.join(HERE, "golden", "mm", name))
        b = oracle_ref.load_matrix(os.path.join(HERE, "golden", "mm", name))
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)


VEXP = json.load(open(os.path.join(HERE, "golden", "mmvec_expected.json")))


@pytest.mark.parametrize("key", sorted(VEXP))
def test_vector_loader_matches_reference_loadvector(g, key):
    """mpg_mm_read_vector_host vs what the reference's own LoadVector<double>() (LoadMatrix.hpp:156-233, through oracle/_ref) returned
    for the same file and column: array and coordinate forms, bit for bit"""
    name, col = key.split(":")
    v = g.read_matrix_market_vector(os.path.join(HERE, "golden", "mmvec", name), int(col))
    np.testing.assert_array_equal(v, np.array([float.fromhex(x) for x in VEXP[key]]))


def test_vector_loader_errors_like_the_reference(g, tmp_path):
    with pytest.raises(g.MpgError, match="Could not access file"):
        g.read_matrix_market_vector(tmp_path / "missing.mtx")
    with pytest.raises(g.MpgError, match="Column 3 is too large for the 3 vectors"):                    # LoadMatrix.hpp:192-197
        g.read_matrix_market_vector(os.path.join(HERE, "golden", "mmvec", "array_3cols.mtx"), 3)
    p = tmp_path / "bad.mtx"
    p.write_text("%MatrixMarket matrix array real general\n2 1\n1\n2\n")
    with pytest.raises(g.MpgError, match="Banner is missing"):
        g.read_matrix_market_vector(p)
    p.write_text("%%MatrixMarket matrix array real general\n")
    with pytest.raises(g.MpgError, match="Malformed matrix size information"):
        g.read_matrix_market_vector(p)


def test_live_reference_vector_loader_agrees(g):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
    import oracle_ref
    if not oracle_ref.available():
        pytest.skip("oracle/_ref not built")
    for key in sorted(VEXP):
        name, col = key.split(":")
        path = os.path.join(HERE, "golden", "mmvec", name)
        np.testing.assert_array_equal(g.read_matrix_market_vector(path, int(col)), oracle_ref.load_vector(path, int(col)))


def _slabs(n):
    cuts = sorted({0, n, n // 3, (2 * n) // 3, min(1, n), max(n - 1, 0), n // 2})
    return list(zip(cuts[:-1], cuts[1:])) + [(0, 0), (n, n), (0, n), (0, -1)]


@pytest.mark.parametrize("name", sorted(EXP))
def test_slab_loader_equals_rows_of_the_full_read(g, name):
    """mpg_mm_read_slab_host (partition-aware ingest): rows [lo, hi) read alone == the same rows of the canonical CSR
    the reference's LoadMatrix builds (explicit zero diagonals, mirrored symmetric entries, unmerged duplicates in stable order), bit
    for bit; the global row map it can return == the full one; the slabs of a partition tile the matrix"""
    path = os.path.join(HERE, "golden", "mm", name)
    e = EXP[name]
    rm, ind = np.array(e["row_map"], np.int32), np.array(e["inds"], np.int32)
    val = np.array([float.fromhex(v) for v in e["vals"]])
    n = len(rm) - 1
    for lo, hi in _slabs(n):
        n_, nnzg, rl, il, vl, rg = g.read_matrix_market_slab(path, lo, hi, want_global_rowmap=True)
        h = n if hi < 0 else hi
        assert n_ == n and nnzg == len(ind)
        np.testing.assert_array_equal(rg, rm)
        np.testing.assert_array_equal(rl, rm[lo:h + 1] - rm[lo])
        np.testing.assert_array_equal(il, ind[rm[lo]:rm[h]])
        np.testing.assert_array_equal(vl, val[rm[lo]:rm[h]])
    # without the global row map
    n_, nnzg, rl, il, vl = g.read_matrix_market_slab(path, 1 if n > 1 else 0, n)
    np.testing.assert_array_equal(il, ind[rm[1 if n > 1 else 0]:])


def test_slab_loader_streams_in_blocks_and_feeds_the_partition(g, orc, tmp_path):
    """a file larger than one 16 MiB block of the streaming tokenizer (tokens must never be torn at a block boundary), entries shuffled,
    mixed white space: 4 ranks read their nnz-balanced slabs alone; together they are the full canonical CSR, and the split points
    computed from the global row map the slab reader returns are the oracle's"""
    rm, ind, val = orc.gen("cd27:40")      # 64 000 rows, 1.6 M nonzeros -> ~ 40 MB of text
    n = len(rm) - 1
    rows = np.repeat(np.arange(n), np.diff(rm))
    perm = np.random.default_rng(5).permutation(len(ind))
    p = tmp_path / "big.mtx"
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n% a comment line\n%another\n")
        f.write(f"{n} {n} {len(ind)}\n")
        seps = ["\n", " \n", "\t\n", "\n\n"]
        chunks = []
        for k, e in enumerate(perm):
            chunks.append(f"{rows[e] + 1}  {ind[e] + 1}\t{val[e]:.17g}{'' if k % 9 else '   '}{seps[k % 4]}")
        f.write("".join(chunks))
    assert os.path.getsize(p) > (1 << 24) + (1 << 20)
    full = g.read_matrix_market(p)
    np.testing.assert_array_equal(full[0], rm); np.testing.assert_array_equal(full[1], ind); np.testing.assert_array_equal(full[2], val)
    _, nnzg, _, _, _, rg = g.read_matrix_market_slab(p, 0, 0, want_global_rowmap=True)
    np.testing.assert_array_equal(rg, rm)
    assert nnzg == len(ind)
    P = 4
    bnd = g.dist.bounds_nnz(rg, P)
    np.testing.assert_array_equal(bnd, orc.partition_bounds_nnz(rm, P))
    got_ind, got_val, got_rm = [], [], [0]
    for r in range(P):
        _, _, rl, il, vl = g.read_matrix_market_slab(p, int(bnd[r]), int(bnd[r + 1]))
        got_ind.append(il); got_val.append(vl); got_rm.extend((rl[1:] + got_rm[-1]).tolist())
    np.testing.assert_array_equal(np.concatenate(got_ind), ind)
    np.testing.assert_array_equal(np.concatenate(got_val), val)
    np.testing.assert_array_equal(np.array(got_rm, np.int32), rm)


@pytest.mark.parametrize("content,msg", [
    ("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n", "Unsupported matrix type"),
    ("%MatrixMarket matrix coordinate real general\n1 1 1\n1 1 1\n", "Banner is missing"),
    ("%%MatrixMarket matrix coordinate real general\n", "Malformed matrix size information"),
    ("%%MatrixMarket matrix coordinate real general\n% only comments\n", "Malformed matrix size information"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 1\n2 2 1\n", "premature end of entries"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1\n2 x 1\n", "premature end of entries"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1\n", "entry index out of range"),
    ("%%MatrixMarket matrix coordinate real general", "Missing values in banner"),
])
def test_slab_loader_errors_like_the_full_reader(g, tmp_path, content, msg):
    p = tmp_path / "bad.mtx"
    p.write_text(content)
    with pytest.raises(g.MpgError, match=msg):
        g.read_matrix_market_slab(p, 0, -1)
    with pytest.raises(g.MpgError, match=msg):     # the one-block reader says the same
        g.read_matrix_market(p)
    with pytest.raises(g.MpgError, match="Could not access file"):
        g.read_matrix_market_slab(tmp_path / "missing.mtx", 0, -1)
    ok = tmp_path / "ok.mtx"
    ok.write_text("%%MatrixMarket matrix coordinate real general\n2 2 1\n1 2 3.5\n")
    with pytest.raises(g.MpgError, match="row range outside the matrix"):
        g.read_matrix_market_slab(ok, 1, 3)
    n, nnzg, rl, il, vl = g.read_matrix_market_slab(ok, 0, -1)
    assert (n, nnzg, rl.tolist(), il.tolist(), vl.tolist()) == (2, 3, [0, 2, 3], [0, 1, 1], [0.0, 3.5, 0.0])


def _canonical_csr(n, entries, symmetric):
    """LoadMatrix.hpp:62-145 in a few lines: a diagonal slot first in every row (file diagonal overwrites it, last one wins), entries in
    file order with the mirrored ones interleaved as read, stable sort by column"""
    rows = [[(r, 0.0)] for r in range(n)]
    for i, j, v in entries:
        if i == j:
            rows[i][0] = (i, v)
            continue
        rows[i].append((j, v))
        if symmetric:
            rows[j].append((i, v))
    rm, ind, val = [0], [], []
    for r in rows:
        r = sorted(r, key=lambda t: t[0])     # Python's sort is stable
        ind += [c for c, _ in r]; val += [v for _, v in r]; rm.append(len(ind))
    return np.array(rm, np.int32), np.array(ind, np.int32), np.array(val, np.float64)


@pytest.mark.parametrize("threads,block", [(1, 1 << 26), (3, 257), (7, 1000), (16, 64)])
@pytest.mark.parametrize("symmetric", [False, True])
def test_parallel_parser_equals_the_sequential_definition(g, tmp_path, monkeypatch, threads, block, symmetric):
    """the loaders convert the entry stream with several host threads over byte ranges and (slab reader) over blocks of the file: whatever
    the number of threads and the block size, tokens torn by range / block boundaries, entries spread over several lines, duplicates,
    repeated diagonals and trailing text after the last entry give the canonical CSR of the sequential definition"""
    rng = np.random.default_rng(100 * threads + symmetric)
    n, nz = 37, 400
    ent = []
    for _ in range(nz):
        i, j = int(rng.integers(n)), int(rng.integers(n))
        if symmetric and j > i:
            i, j = j, i
        ent.append((i, j, float(np.float32(rng.standard_normal()))))
    seps = [" ", "\t", "  ", "\n", " \n ", "\r\n"]
    text = f"%%MatrixMarket matrix coordinate real {'symmetric' if symmetric else 'general'}\n%c\n{n} {n} {nz}\n"
    for k, (i, j, v) in enumerate(ent):
        text += f"{i + 1}{seps[k % 6]}{j + 1}{seps[(k + 1) % 6]}{v!r}{seps[(k + 3) % 6] if k % 5 else chr(10)}"
    text += "\nthis is not an entry 1 2 3\n"      # ignored: the reader stops after nz entries (LoadMatrix.hpp:68)
    p = tmp_path / "m.mtx"
    p.write_text(text)
    monkeypatch.setenv("MPG_LOADER_THREADS", str(threads))
    monkeypatch.setenv("MPG_LOADER_BLOCK", str(block))
    rm, ind, val = _canonical_csr(n, ent, symmetric)
    a = g.read_matrix_market(p)
    np.testing.assert_array_equal(a[0], rm); np.testing.assert_array_equal(a[1], ind); np.testing.assert_array_equal(a[2], val)
    for lo, hi in [(0, n), (5, 21), (36, 37), (0, 0)]:
        _, nnzg, rl, il, vl, rg = g.read_matrix_market_slab(p, lo, hi, want_global_rowmap=True)
        assert nnzg == len(ind)
        np.testing.assert_array_equal(rg, rm)
        np.testing.assert_array_equal(rl, rm[lo:hi + 1] - rm[lo])
        np.testing.assert_array_equal(il, ind[rm[lo]:rm[hi]]); np.testing.assert_array_equal(vl, val[rm[lo]:rm[hi]])


@pytest.mark.parametrize("threads,block", [(1, 1 << 26), (4, 64)])
@pytest.mark.parametrize("body,msg", [
    ("1 1 1\n9 1 1\n1 1 x\n1 1 1\n", "entry index out of range"),          # the range error is in an earlier entry than the bad token
    ("1 1 1\n1 1 x\n9 1 1\n1 1 1\n", "premature end of entries"),          # the bad token comes first
    ("1 1 1\n9 1 x\n1 1 1\n1 1 1\n", "premature end of entries"),          # same entry: the three fields are read before the indices are checked
    ("1 1 1\n1 0 1\n1 1 1\n1 1 1\n", "entry index out of range"),
    ("1 1 1\n1 1 1\n1 1 1\n1 1\n", "premature end of entries"),            # the file ends inside the last entry
    ("1 1 1\n1 1.5 1\n1 1 1\n1 1 1\n", "premature end of entries"),        # an index that is not an integer
])
def test_parallel_parser_reports_the_first_error_in_file_order(g, tmp_path, monkeypatch, threads, block, body, msg):
    p = tmp_path / "bad.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n2 2 4\n" + body)
    monkeypatch.setenv("MPG_LOADER_THREADS", str(threads))
    monkeypatch.setenv("MPG_LOADER_BLOCK", str(block))
    with pytest.raises(g.MpgError, match=msg):
        g.read_matrix_market(p)
    with pytest.raises(g.MpgError, match=msg):
        g.read_matrix_market_slab(p, 0, -1)
