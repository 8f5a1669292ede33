"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly what include/b200mnn.h declares,
validates arguments with the reference's messages, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import batchelor_b200 as bb
from batchelor_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "b200mnn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200mnn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = bb.load()
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200mnn.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes signature table and header disagree"
    exported = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\b(b200mnn_[a-z0-9_]+)\b", exported)))
    assert exported == declared, "library exports symbols the header does not declare (or vice versa)"


def test_version_and_error_string():
    lib = bb.load()
    assert lib.b200mnn_version() >= 100
    assert isinstance(lib.b200mnn_last_error(), bytes)


def test_library_contains_blackwell_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_argument_errors_carry_the_reference_messages():
    # these checks happen before any device work, exactly like the reference's throw sites
    with pytest.raises(bb.B200Error, match="must have length equal to number of rows in 'averaged'"):
        bb.smooth_gaussian_kernel(np.zeros((3, 4)), np.zeros(2, np.int32), np.zeros((3, 10)), 0.1)  # smooth_gaussian_kernel.cpp:18-20
    with pytest.raises(bb.B200Error, match="number of genes do not match up between matrices"):
        bb.adjust_shift_variance(np.zeros((3, 4)), np.zeros((2, 5)), np.zeros((5, 3)), 1.0, [0], [0])  # adjust_shift_variance.cpp:33-36
    with pytest.raises(bb.B200Error, match="number of cells do not match up between matrices"):
        bb.adjust_shift_variance(np.zeros((3, 4)), np.zeros((3, 5)), np.zeros((4, 3)), 1.0, [0], [0])  # :38-41
    with pytest.raises(bb.B200Error, match="subset indices out of range"):
        bb.adjust_shift_variance(np.zeros((3, 4)), np.zeros((3, 5)), np.zeros((5, 3)), 1.0, [4], [0])  # utils.cpp:6-13
    with pytest.raises(bb.B200Error, match="subset indices out of range"):
        bb.adjust_shift_variance(np.zeros((3, 4)), np.zeros((3, 5)), np.zeros((5, 3)), 1.0, [0], [-1])


def _has_gpu():
    return bb.load().b200mnn_device_count() > 0


@pytest.mark.skipif(_has_gpu(), reason="this box has a GPU")
def test_no_cpu_fallback_without_a_device():
    X = np.random.default_rng(0).normal(size=(50, 5))
    for fn in (lambda: bb.queryKNN(X, X, 3), lambda: bb.findMutualNN(X, X, 3), lambda: bb.cosineNorm(X),
               lambda: bb.reducedMNN(X, X + 1), lambda: bb.mnnCorrect(X.T, X.T + 1),
               lambda: bb.find_mutual_nns(np.ones((4, 2), np.int32), np.ones((4, 2), np.int32))):
        with pytest.raises(bb.B200Error, match="no usable CUDA device|no CPU fallback"):
            fn()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "batchelor_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"from\s+\.+\s*oracle", r"liboracle|libbatchelor_ref|oracle/", r"#include\s+.*oracle"):
                    assert not re.search(pat, text, flags=re.M), f"{f} reaches into oracle/ ({pat})"


def test_host_side_tree_and_bookkeeping():
    """Host control flow of the product mirror (no device needed): merge order, k selection, pair re-indexing."""
    from batchelor_b200 import api

    assert api._predefined_tree(4, None) == [[[1, 2], 3], 4]
    assert api._predefined_tree(4, [[1, 2], [3, 4]]) == [[1, 2], [3, 4]]
    assert api._predefined_tree(3, [3, 1, 2]) == [[3, 1], 2]
    with pytest.raises(ValueError, match="invalid leaf nodes"):
        api._predefined_tree(3, [1, 1, 2])
    tree = [["a", "b"], ["c", "d"]]
    assert api._next_merge(tree) == ("c", "d", (1,))  # right subtree first (R/MNN_tree.R:61-69)
    assert api._choose_k(20, None, 10 ** 6) == 20 and api._choose_k(10, 0.05, 1000) == 50 and api._choose_k(10, 0.9, 15) == 14
    assert np.array_equal(api._restore_original_order([3, 1, 2], [2, 3, 1]), [2, 3, 4, 5, 6, 1])
    p = api._reindex_pairings([{"left": np.array([1, 2]), "right": np.array([3, 4])}], np.array([2, 3, 4, 1]))
    assert np.array_equal(p[0]["left"], [4, 1]) and np.array_equal(p[0]["right"], [2, 3])
    parts, reorder, restricted, levels = api._divide_into_batches(np.arange(12.0).reshape(6, 2), [2, 1, 2, 1, 1, 2], True, [1, 2, 6])
    assert [x.shape[0] for x in parts] == [3, 3] and np.array_equal(reorder, [4, 1, 5, 2, 3, 6])
    assert np.array_equal(restricted[0], [1]) and np.array_equal(restricted[1], [1, 3])
    with pytest.raises(ValueError, match="serial BPPARAM"):
        bb.queryKNN(np.zeros((3, 2)), np.zeros((3, 2)), 1, BPPARAM=bb.SerialParam(workers=4))
