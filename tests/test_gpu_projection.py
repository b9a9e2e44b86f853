"""Parity of kernel 1 (tcgen05 projection GEMM) with the oracle's exact fixed-point projection:
bit-exact, and within 1e-5 relative of a plain fp32 k-ascending GEMM."""
import numpy as np
import pytest

from maplab_b200 import capi, synthetic
from oracle import pyoracle as po
from helpers import small_world

pytestmark = pytest.mark.gpu


def _check(blob, bits):
    det = capi.Detector(blob)
    got = det.project(bits)
    v = synthetic.parse_vocabulary(blob)
    P = v["P"]
    exp = po.project(np.asfortranarray(P).reshape(-1, order="F"), P.shape[0], P.shape[1],
                     v["target_dim"], bits)
    assert got.shape == exp.shape
    assert np.array_equal(got, exp)
    flt = po.project(np.asfortranarray(P).reshape(-1, order="F"), P.shape[0], P.shape[1],
                     v["target_dim"], bits, float_mode=True)
    scale = np.abs(flt).max()
    assert np.abs(got - flt).max() <= 1e-5 * scale  # north_star tolerance: 1e-5 relative
    return det


@pytest.mark.parametrize("n", [1, 7, 128, 129, 1000, 70001])
def test_projection_sizes(n):
    m, blob, _, _ = small_world()
    rng = np.random.default_rng(n)
    bits = rng.integers(0, 256, size=(n, 64), dtype=np.uint8)
    _check(blob, bits)


def test_projection_empty_block_is_noop():
    m, blob, _, _ = small_world()
    det = capi.Detector(blob)
    out = det.project(np.zeros((0, 64), np.uint8))
    assert out.shape == (0, 10)


def test_projection_extreme_descriptors():
    m, blob, _, _ = small_world()
    bits = np.zeros((300, 64), np.uint8)
    bits[1] = 0xFF
    bits[2, 0] = 1           # only bit 0
    bits[3, 63] = 0x80       # only bit 511
    for i in range(4, 300):  # single bits: pins the LSB-first order of DescriptorToEigenMatrix
        bits[i, (i - 4) // 8 % 64] = 1 << ((i - 4) % 8)
    _check(blob, bits)


def test_projection_brisk_384_and_471_columns():
    rng = np.random.default_rng(5)
    # BRISK: 48-byte descriptors, 384x384 matrix of which the top 10 rows are used
    P = rng.standard_normal((384, 384)).astype(np.float32)
    W = rng.standard_normal((5, 16)).astype(np.float32)
    blob = synthetic.serialize_vocabulary(P, W, W, target_dim=10)
    _check(blob, rng.integers(0, 256, size=(777, 48), dtype=np.uint8))
    # FREAK special case: 471-column matrix uses the first 471 of 512 bits
    P = (rng.standard_normal((10, 471)) * 0.1).astype(np.float32)
    blob = synthetic.serialize_vocabulary(P, W, W, target_dim=10)
    _check(blob, rng.integers(0, 256, size=(515, 64), dtype=np.uint8))


def test_projection_rejects_short_descriptors():
    m, blob, _, _ = small_world()
    det = capi.Detector(blob)
    with pytest.raises(capi.MlcError):
        det.project(np.zeros((4, 48), np.uint8))  # 512-column matrix, 384-bit descriptors
