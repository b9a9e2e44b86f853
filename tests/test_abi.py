"""CPU-side checks: the C-ABI library loads and exports every symbol include/maplab_lc_b200.h
declares; host-side structures have the header's layout; without a GPU compute calls fail loudly
(there is no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from maplab_b200 import capi, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "maplab_lc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mlc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(capi.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(capi.LIB_PATH)
    declared = _header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(capi.EXPORTS) == declared


def test_struct_layouts_match_header():
    assert ctypes.sizeof(capi.Settings) == 80  # + shard_mode, float_descriptor_dim, hnsw_*
    assert capi.FRAME_DTYPE.itemsize == ctypes.sizeof(capi.Frame) == 32
    assert capi.MATCH_DTYPE.itemsize == 32
    assert capi.CAMERA_DTYPE.itemsize == 8 * 4 + 8 + 8 * 4 + 8 * 9 + 8 * 3
    assert capi.POSE_DTYPE.itemsize == 4 * 10 + 8 + 96
    assert ctypes.sizeof(capi.RansacSettings) == 48


def test_default_settings_are_the_reference_flags():
    s = capi.default_settings()
    assert (s.num_closest_words, s.num_nearest_neighbors, s.scoring, s.engine) == (10, -1, 0, 0)
    assert s.min_image_time_seconds == 10.0 and s.min_verify_matches_num == 10
    assert s.fraction_best_scores == 0.25 and s.knn_epsilon == 2.0 and s.knn_max_radius == 20.0
    r = capi.default_ransac_settings()
    assert (r.min_inlier_count, r.num_ransac_iters, r.seed) == (10, 100, 12345)
    assert r.ransac_pixel_sigma == 2.0 and r.min_inlier_ratio == 0.0
    assert r.max_delta_position_m == -1.0 and r.max_delta_rotation_deg == -1.0   # gates off
    a = capi.AlignmentSettings()
    capi.lib().mlc_default_alignment_settings(ctypes.byref(a))
    assert (a.num_iterations, a.rng_mapping) == (2000, 1)                         # anchor_transform_* flags
    assert a.max_orientation_error_rad == 0.174 and a.max_position_error_m == 2.0
    assert ctypes.sizeof(capi.AlignmentSettings) == 32


def test_vocabulary_roundtrip_and_errors():
    rng = np.random.default_rng(0)
    P = rng.standard_normal((10, 512)).astype(np.float32)
    W1 = rng.standard_normal((5, 20)).astype(np.float32)
    W2 = rng.standard_normal((5, 30)).astype(np.float32)
    blob = synthetic.serialize_vocabulary(P, W1, W2)
    v = synthetic.parse_vocabulary(blob)
    assert v["target_dim"] == 10 and np.array_equal(v["P"], P) and np.array_equal(v["W2"], W2)
    # a truncated quantizer file is rejected before any device work
    with pytest.raises(capi.MlcError, match="vocabulary"):
        capi.Detector(blob[:100])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = synthetic.make_map(300, seed=2)
    blob, _ = synthetic.make_vocabulary(m["bits"], num_words=8)
    with pytest.raises(capi.MlcError):
        capi.Detector(blob)
