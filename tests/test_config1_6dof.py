"""BASELINE.json config 1: self loop-closure on the `vi-map-generator-6dof` map (recipe restated in
maplab_b200/synthetic.make_6dof_map: 20 vertices, 500 landmarks, 48-byte descriptors that are IDENTICAL for
all observations of a landmark, every timestamp 0, one mission) with the shipped BRISK vocabulary (its first
10 projection rows), `--lc_min_image_time_seconds=0` as the `lc` console command needs on this map
(matching-based-engine.cc:192-198). Exercises the 384-bit projection and the distance ties between identical
descriptors (broken by descriptor index) end to end."""
import os

import numpy as np
import pytest

from maplab_b200 import capi, synthetic
from oracle import pyoracle as po
from helpers import fill_oracle, frames_of

BLOB = open(os.path.join(os.path.dirname(__file__), "golden", "brisk_quantizer_top10.dat"), "rb").read()


def _oracle(m, **kw):
    ora = po.Engine(BLOB, po.default_settings(min_image_time_seconds=0.0, **kw))
    proj = ora.project(m["bits"])
    fill_oracle(ora, frames_of(m["frames"]), proj, m["landmarks"])
    c = m["camera"]
    cams = [po.make_camera(c["fu"], c["fv"], c["cu"], c["cv"], c["R_B_C"], c["t_B_C"])]
    return ora, proj, cams


def test_recipe_shape():
    m = synthetic.make_6dof_map()
    assert len(m["frames"]["num_descriptors"]) == 20 and len(m["landmark_xyz"]) == 500
    assert m["bits"].shape[1] == 48 and (m["frames"]["timestamp_ns"] == 0).all()
    assert len(set(m["frames"]["mission_id"].tolist())) == 1
    # all observations of a landmark carry the same descriptor, 30 deterministic bits set
    assert np.array_equal(m["bits"], m["base"][m["landmarks"]])
    for i in range(30):
        assert (m["bits"][:, i] & (1 << (i % 8))).all()
    # landmarks drawn like generateLandmarksCircle: ring of radius 15 +- 1.5 m, height +- 3 m
    r = np.hypot(m["landmark_xyz"][:, 0], m["landmark_xyz"][:, 1])
    assert r.min() >= 13.5 and r.max() <= 16.5 and np.abs(m["landmark_xyz"][:, 2]).max() <= 3.0


def test_oracle_self_loop_closure():
    m = synthetic.make_6dof_map()
    ora, proj, cams = _oracle(m)
    assert ora.num_neighbors() == 1  # < 1e4 descriptors (matching-based-engine.cc:322-324)
    r = po.query_batch(ora, m["frames"], m["bits"], m["keypoints"], m["landmark_xyz"], cams)
    big = m["frames"]["num_descriptors"] >= 80
    assert r["accepted"][big].all()
    ok = r["accepted"].astype(bool)
    # exact keypoints: the true poses come back (a minimal-sample model without refinement, so the odd
    # hypothesis that gathers enough inliers on this thin landmark ring may be off)
    err = np.abs(r["T"][ok] - m["T_G_I"][ok]).reshape(-1, 12).max(1)
    assert np.median(err) < 1e-6 and (err < 1e-3).mean() >= 0.8
    # with the default 10 s window every match is dropped: all timestamps are 0 and there is one mission
    ora10 = po.Engine(BLOB)
    fill_oracle(ora10, frames_of(m["frames"]), proj, m["landmarks"])
    r10 = po.query_batch(ora10, m["frames"], m["bits"], m["keypoints"], m["landmark_xyz"], cams)
    assert r10["accepted"].sum() == 0 and r10["num_matches"].sum() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("k", [-1, 4])
def test_device_equals_oracle(k):
    m = synthetic.make_6dof_map()
    ora, oproj, ocams = _oracle(m, num_nearest_neighbors=k)
    det = capi.Detector(BLOB, capi.default_settings(min_image_time_seconds=0.0, num_nearest_neighbors=k))
    proj = det.project(m["bits"])
    assert np.array_equal(proj, oproj)  # 384-bit descriptors
    frames = frames_of(m["frames"])
    det.insert_batch(frames, proj, m["landmarks"])
    det.set_landmark_positions(m["landmark_xyz"])
    kk = det.num_neighbors()
    assert kk == ora.num_neighbors()
    idx, dist = det.knn(proj, kk)
    oidx, odist = ora.knn(proj, kk)
    assert np.array_equal(idx, oidx) and np.array_equal(dist, odist)
    # (nearly) every descriptor finds an identical one — the epsilon-approximate word search may send a
    # query past its own cell — and ties go to the smallest index
    zero = dist[:, 0] == 0
    assert zero.mean() > 0.9
    assert (idx[zero, 0] <= np.arange(len(idx))[zero]).all()
    out = det.query_batch(frames, m["bits"], m["keypoints"], capi.make_cameras([m["camera"]]))
    exp = po.query_batch(ora, frames, m["bits"], m["keypoints"], m["landmark_xyz"], ocams)
    res = out["results"]
    for f in ("accepted", "num_inliers", "iterations", "ransac_success"):
        assert np.array_equal(res[f], exp[f]), f
    assert np.array_equal(np.diff(out["offsets"]), exp["num_matches"])
    ok = exp["ransac_success"].astype(bool)
    assert np.array_equal(res["T_G_I"].reshape(-1, 3, 4)[ok], exp["T"][ok])
    assert res["accepted"].sum() >= 10
