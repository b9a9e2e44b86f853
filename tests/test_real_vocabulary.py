"""Realism check with maplab's SHIPPED FREAK vocabulary (tests/golden/, see
fetch_shipped_vocabulary.py): the file parses (T2 layout), the exact fixed-point projection agrees
with a plain fp32 product of the real matrix, and — on the GPU — projection, visited cells and kNN
are bit-identical to the oracle. Random 512-bit descriptors collapse into a few hundred of the
10^6 cells under this vocabulary (SURVEY §8d note), so the inverted lists are thousands of entries
long: this is the multi-trip path of the list scan, which the small synthetic worlds barely touch."""
import os

import numpy as np
import pytest

from maplab_b200 import synthetic
from oracle import pyoracle as po

BLOB = open(os.path.join(os.path.dirname(__file__), "golden", "inverted_multi_index_quantizer_freak.dat"), "rb").read()


def test_shipped_vocabulary_parses_and_projects():
    v = synthetic.parse_vocabulary(BLOB)
    assert v["target_dim"] == 10 and v["P"].shape == (10, 512)
    assert v["W1"].shape == (5, 1000) and v["W2"].shape == (5, 1000)
    assert np.abs(np.linalg.norm(v["P"], axis=1) - 1.0).max() < 1e-3      # unit-norm rows
    ora = po.Engine(BLOB)
    bits = np.random.default_rng(0).integers(0, 256, size=(2000, 64), dtype=np.uint8)
    got = ora.project(bits)
    ref = synthetic.project_float(v["P"], bits)
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 6, 10])
def test_shipped_vocabulary_gpu_parity_long_lists(k):
    from maplab_b200 import capi
    rng = np.random.default_rng(k)
    bits = rng.integers(0, 256, size=(30000, 64), dtype=np.uint8)
    qbits = rng.integers(0, 256, size=(1500, 64), dtype=np.uint8)
    det, ora = capi.Detector(BLOB), po.Engine(BLOB)
    proj, qp = det.project(bits), det.project(qbits)
    assert np.array_equal(proj, ora.project(bits)) and np.array_equal(qp, ora.project(qbits))
    n = len(bits)
    det.insert(0, 0, 0, 0, proj, np.arange(n))
    ora.insert(0, 0, 0, 0, proj, np.arange(n))
    idx, dist = det.knn(qp, k)
    oidx, odist = ora.knn(qp, k)
    assert np.array_equal(idx, oidx) and np.array_equal(dist, odist)
    st = det.last_scan_stats()
    assert st["entries"] / len(qp) > 500, "expected long inverted lists under the shipped vocabulary"
