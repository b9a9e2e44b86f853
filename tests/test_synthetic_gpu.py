"""The counter-based generator of the large configurations (maplab_b200/synthetic_gpu.py): a small world
materialised on the CPU goes through the oracle (loop closures recover the ground-truth poses), and the
shard-aware device build of the same world answers exactly like the oracle."""
import numpy as np
import pytest
import torch

from maplab_b200 import capi, synthetic, synthetic_gpu as sg
from oracle import pyoracle as po
from helpers import fill_oracle

L = 12_000


def _host_world():
    dev = torch.device("cpu")
    lm_per_kf, num_kf = sg.layout(L)
    counts, lm = sg.observations(L, 0, num_kf, dev)
    gidx = torch.arange(lm.shape[0], dtype=torch.int64)
    bits = sg.descriptor_bytes(lm, gidx).numpy()
    frames = sg.frames_for(0, counts.numpy(), 1, num_kf)
    xyz = sg.all_landmark_xyz(L, dev).numpy()
    blob, _ = synthetic.make_vocabulary(sg.vocabulary_sample(L, 20_000), num_words=64, seed=3)
    q = sg.make_queries(L, 0, 12, dev)
    return frames, lm.numpy(), bits, xyz, blob, q


def test_world_shape_and_determinism():
    frames, lm, bits, xyz, blob, q = _host_world()
    assert abs(frames["num_descriptors"].mean() - 500) < 15
    assert abs(np.bincount(lm, minlength=L).mean() - 4.0) < 0.1
    assert np.array_equal(sg.vocabulary_sample(L, 20_000), bits[:20_000])   # any part can be regenerated alone
    q2 = sg.make_queries(L, 4, 8, torch.device("cpu"))
    assert np.array_equal(q2["bits"].numpy(), q["bits"].numpy()[4 * 500:8 * 500])
    assert np.array_equal(q2["keypoints"].numpy(), q["keypoints"].numpy()[4 * 500:8 * 500])
    assert (q["true_landmark"].numpy() >= 0).mean() == pytest.approx(0.8, abs=0.01)


def test_oracle_closes_the_loops_of_the_hash_world():
    frames, lm, bits, xyz, blob, q = _host_world()
    ora = po.Engine(blob, po.default_settings(num_nearest_neighbors=6))
    fill_oracle(ora, frames, ora.project(bits), lm)
    cam = synthetic.camera_dict()
    r = po.query_batch(ora, q["frames"], q["bits"].numpy(), q["keypoints"].numpy(), xyz,
                       [po.make_camera(cam["fu"], cam["fv"], cam["cu"], cam["cv"])])
    acc = r["accepted"].astype(bool)
    assert acc.mean() > 0.9
    assert np.abs(r["T"][acc][:, :, 3] - q["T_G_I"][acc][:, :, 3]).max() < 0.25


@pytest.mark.gpu
def test_device_build_equals_oracle():
    frames, lm, bits, xyz, blob, q = _host_world()
    dev = torch.device("cuda", 0)
    ora = po.Engine(blob, po.default_settings(num_nearest_neighbors=6))
    fill_oracle(ora, frames, ora.project(bits), lm)
    det = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=6))
    info = sg.build_database(det, L, 0, 1, dev, chunk_kf=17)
    assert info["num_descriptors"] == len(bits) == det.num_descriptors()
    det.set_landmark_positions(xyz)
    qd = sg.make_queries(L, 0, 12, dev)
    assert np.array_equal(qd["bits"].cpu().numpy(), q["bits"].numpy())
    assert np.allclose(qd["keypoints"].cpu().numpy(), q["keypoints"].numpy(), atol=1e-9)
    cam = synthetic.camera_dict()
    kp = qd["keypoints"].cpu().numpy()
    out = det.query_batch(qd["frames"], qd["bits"].cpu().numpy(), kp, capi.make_cameras([cam]))
    exp = po.query_batch(ora, qd["frames"], qd["bits"].cpu().numpy(), kp, xyz,
                         [po.make_camera(cam["fu"], cam["fv"], cam["cu"], cam["cv"])])
    res = out["results"]
    for f in ("accepted", "num_inliers", "iterations", "ransac_success"):
        assert np.array_equal(res[f], exp[f]), f
    ok = exp["ransac_success"].astype(bool)
    assert np.array_equal(res["T_G_I"].reshape(-1, 3, 4)[ok], exp["T"][ok])
    assert res["accepted"].mean() > 0.9
