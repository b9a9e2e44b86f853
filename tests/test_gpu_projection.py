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


def _oracle_project_threads(ora, bits, threads=16):
    import threading
    out = np.empty((len(bits), 10), np.float32)
    chunks = [(s, min(s + 65536, len(bits))) for s in range(0, len(bits), 65536)]

    def work(t):
        for ci in range(t, len(chunks), threads):
            s, e = chunks[ci]
            out[s:e] = ora.project(bits[s:e])

    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return out


def test_projection_many_tiles_per_cta_warm_l2():
    """1 M descriptors = 53 tiles per persistent CTA, launched repeatedly so that the source tiles come from L2:
    the regime in which the raw shared-memory ring was once overwritten under the producers' loads (0.07 % wrong
    rows) while every small-size test stayed green."""
    import torch
    n = 1_000_000
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 256, (n, 64), dtype=np.uint8)
    blob, _ = synthetic.make_vocabulary(bits[:20000], num_words=64, seed=7)
    det = capi.Detector(blob)
    exp = _oracle_project_threads(po.Engine(blob), bits)
    bits_d = torch.from_numpy(bits).cuda()
    out_d = torch.empty((n, 10), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(4):
        out_d.zero_()
        det.project_device(bits_d.data_ptr(), 64, n, out_d.data_ptr(), st)
        torch.cuda.synchronize()
        bad = np.nonzero((out_d.cpu().numpy() != exp).any(1))[0]
        assert len(bad) == 0, f"launch {rep}: {len(bad)} rows differ from the oracle, first {bad[:5]}"


@pytest.mark.parametrize("nbytes", [64, 48])
def test_projection_unaligned_device_buffers(nbytes):
    """Caller buffers that are not 16-byte aligned take the shared-memory-A kernel without bulk copies."""
    import torch
    n = 70_001
    rng = np.random.default_rng(nbytes)
    bits = rng.integers(0, 256, (n, nbytes), dtype=np.uint8)
    if nbytes == 64:
        _, blob, _, _ = small_world()
    else:
        P = rng.standard_normal((384, 384)).astype(np.float32)
        W = rng.standard_normal((5, 16)).astype(np.float32)
        blob = synthetic.serialize_vocabulary(P, W, W, target_dim=10)
    det = capi.Detector(blob)
    exp = det.project(bits)  # aligned path, itself checked against the oracle above
    raw_in = torch.empty(n * nbytes + 64, dtype=torch.uint8, device="cuda")
    raw_out = torch.zeros(n * 10 + 16, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for in_off, out_off in [(4, 0), (0, 1), (8, 3)]:
        src = raw_in[in_off:in_off + n * nbytes]
        src.copy_(torch.from_numpy(bits).reshape(-1))
        dst = raw_out[out_off:out_off + n * 10]
        dst.zero_()
        assert src.data_ptr() % 16 != 0 or dst.data_ptr() % 16 != 0
        det.project_device(src.data_ptr(), nbytes, n, dst.data_ptr(), st)
        torch.cuda.synchronize()
        assert np.array_equal(dst.cpu().numpy().reshape(n, 10), exp)
