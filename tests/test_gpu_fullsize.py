"""BASELINE.json's headline configuration (1 M landmarks, ~4 M database descriptors, W = 1000 x 1000
cells, 500-descriptor query keyframes, auto-k = 6): the CUDA path diffed against the CPU oracle on
the same world bench.py measures (visited cells, cell assignment, kNN indices and distances, match
counts, verdicts, inlier counts, RANSAC iterations and poses — bit-exact), plus size-independent
properties: determinism, host path == device path, batch-order invariance, sharded merge == single
index, ground-truth pose recovery."""
import numpy as np
import pytest

from maplab_b200 import capi, synthetic
from oracle import pyoracle as po
from helpers import fill_oracle, frames_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    m = synthetic.make_map(1_000_000, seed=1)
    stride = max(len(m["bits"]) // 100_000, 1)
    blob, _ = synthetic.make_vocabulary(m["bits"][::stride][:100_000], num_words=1000, seed=7)
    q = synthetic.make_queries(m, 160, seed=11)
    det = capi.Detector(blob)
    frames = frames_of(m["frames"])
    proj = np.empty((len(m["bits"]), det.dim), np.float32)
    for s in range(0, len(proj), 1 << 20):
        proj[s:s + (1 << 20)] = det.project(m["bits"][s:s + (1 << 20)])
    det.insert_batch(frames, proj, m["landmarks"])
    det.set_landmark_positions(m["landmark_xyz"])
    return m, blob, q, det, frames, proj


@pytest.fixture(scope="module")
def oracle_world(world):
    m, blob, q, det, frames, proj = world
    ora = po.Engine(blob)
    fill_oracle(ora, frames, proj, m["landmarks"])
    return ora


def test_full_size_matches_oracle(world, oracle_world):
    """The headline configuration itself against the oracle (W = 1000 trees, ~4-entry lists in a
    10^6-cell table, auto-k = 6, the persistent scan at full grid): nothing here is property-based."""
    m, blob, q, det, frames, proj = world
    ora = oracle_world
    assert det.num_descriptors() == ora.num_descriptors() == len(proj)
    k = det.num_neighbors()
    assert k == ora.num_neighbors() == 6
    # P1: projection of a database sample and of all query descriptors
    sample = np.arange(0, len(proj), 37)[:100_000]
    assert np.array_equal(proj[sample], ora.project(m["bits"][sample]))
    qp = det.project(q["bits"])
    assert np.array_equal(qp, ora.project(q["bits"]))
    assert len(qp) == 80_000
    # P2 / P3: cell of database descriptors, visited cells of every query descriptor
    v = synthetic.parse_vocabulary(blob)
    imi = po.IMI(po.colmajor(v["W1"]), v["W1"].shape[1], po.colmajor(v["W2"]), v["W2"].shape[1], 5, 10)
    assert np.array_equal(det.coarse_cells(qp, 10), imi.visited_cells(qp, len(qp)))
    db_sample = proj[sample[:20_000]]
    assert np.array_equal(det.coarse_cells(db_sample, 1)[:, 0],
                          np.array([imi.cell_of(d) for d in db_sample]))
    # P4: kNN indices and distances at the auto-k, all 80 000 query descriptors
    idx, dist = det.knn(qp, k)
    oidx, odist = ora.knn(qp, k)
    assert np.array_equal(idx, oidx)
    assert np.array_equal(dist, odist)
    assert (idx >= 0).mean() > 0.5
    # P5-P9: the fused query of all 160 keyframes
    cam = synthetic.camera_dict()
    qframes = frames_of(q["frames"])
    out = det.query_batch(qframes, q["bits"], q["keypoints"], capi.make_cameras([cam]))
    exp = po.query_batch(ora, qframes, q["bits"], q["keypoints"], m["landmark_xyz"],
                         [po.make_camera(cam["fu"], cam["fv"], cam["cu"], cam["cv"])], num_threads=8)
    res = out["results"]
    assert np.array_equal(res["accepted"], exp["accepted"])
    assert np.array_equal(res["num_inliers"], exp["num_inliers"])
    assert np.array_equal(res["iterations"], exp["iterations"])
    assert np.array_equal(res["ransac_success"], exp["ransac_success"])
    assert np.array_equal(np.diff(out["offsets"]), exp["num_matches"])
    ok = exp["ransac_success"].astype(bool)
    assert np.array_equal(res["T_G_I"].reshape(-1, 3, 4)[ok], exp["T"][ok])
    assert res["accepted"].sum() >= 150


def test_full_size_determinism_host_device_and_order(world):
    import torch
    m, blob, q, det, _, _ = world
    cams = capi.make_cameras([synthetic.camera_dict()])
    qframes = frames_of(q["frames"])
    kp = np.ascontiguousarray(q["keypoints"], np.float64)
    r1 = det.query_batch(qframes, q["bits"], kp, cams, want_matches=True)
    r2 = det.query_batch(qframes, q["bits"], kp, cams, want_matches=True)
    assert r1["results"].tobytes() == r2["results"].tobytes()              # idempotent
    assert r1["matches"].tobytes() == r2["matches"].tobytes()
    bits_d, kp_d = torch.from_numpy(q["bits"]).cuda(), torch.from_numpy(kp).cuda()
    r3 = det.query_batch_device(qframes, bits_d.data_ptr(), 64, kp_d.data_ptr(), cams)
    assert r1["results"].tobytes() == r3["results"].tobytes()              # host path == device path
    # reversing the batch reverses the per-vertex verdicts (queries are independent)
    n = 500
    order = np.arange(len(qframes))[::-1]
    rb = det.query_batch(qframes[order].copy(), q["bits"].reshape(-1, n, 64)[order].reshape(-1, 64),
                         kp.reshape(-1, n, 2)[order].reshape(-1, 2), cams)
    assert rb["results"][::-1].tobytes() == r1["results"].tobytes()
    # accepted closures recover the ground-truth pose of the query keyframe
    acc = r1["results"]["accepted"].astype(bool)
    assert acc.mean() > 0.95
    T = r1["results"]["T_G_I"].reshape(-1, 3, 4)[acc]
    assert np.abs(T[:, :, 3] - q["T_G_I"][acc][:, :, 3]).max() < 0.25


def test_full_size_sharded_merge_equals_single_index(world):
    import torch
    m, blob, q, det, frames, proj = world
    k, G = det.num_neighbors(), 2
    qp = det.project(q["bits"])
    ref_i, ref_d = det.knn(qp, k)
    n = len(qp)
    il = torch.empty((G, n, k), dtype=torch.int32, device="cuda")
    dl = torch.empty((G, n, k), dtype=torch.float32, device="cuda")
    for r in range(G):
        sh = capi.Detector(blob, capi.default_settings(shard_rank=r, shard_count=G))
        sh.insert_batch(frames, proj, m["landmarks"])
        i, d = sh.knn(qp, k)
        il[r], dl[r] = torch.from_numpy(i).cuda(), torch.from_numpy(d).cuda()
        del sh
    oi = torch.empty((n, k), dtype=torch.int32, device="cuda")
    od = torch.empty((n, k), dtype=torch.float32, device="cuda")
    det.merge_topk_device(il.data_ptr(), dl.data_ptr(), G, n, k, oi.data_ptr(), od.data_ptr(),
                          torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(oi.cpu().numpy(), ref_i) and np.array_equal(od.cpu().numpy(), ref_d)
    # every neighbour list is sorted by (distance, index) with the missing entries trailing
    dist_rank = np.where(np.isinf(ref_d), np.float64(3.5e38), ref_d.astype(np.float64))
    later = (dist_rank[:, 1:] > dist_rank[:, :-1]) | ((dist_rank[:, 1:] == dist_rank[:, :-1]) &
                                                      ((ref_i[:, 1:] > ref_i[:, :-1]) | (ref_i[:, 1:] < 0)))
    assert later.all()
    assert (np.isinf(ref_d) == (ref_i < 0)).all()
