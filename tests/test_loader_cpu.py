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
