"""Worker of tests/test_gpu_sharded_nccl.py: one rank of a sharded loop-closure query step through the
C-ABI (mlc_comm_init + mlc_sharded_query_batch), launched by torch.distributed.run with one process per
GPU. Each rank builds only its shard of the database (owned rows), brings a RAGGED slice of the query
keyframes and checks its verdicts against the CPU oracle on the whole database."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from maplab_b200 import capi, synthetic
    from oracle import pyoracle as po
    from helpers import fill_oracle, frames_of, small_world

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m, blob, _, q = small_world(num_queries=8)
    frames = frames_of(m["frames"])
    kw = dict(num_nearest_neighbors=6)
    det = capi.Detector(blob, capi.default_settings(device=local, shard_rank=rank, shard_count=world, **kw))
    ora = po.Engine(blob, po.default_settings(**kw))
    oproj = ora.project(m["bits"])
    fill_oracle(ora, frames, oproj, m["landmarks"])
    # shard-aware build: project and insert the owned descriptors only
    n = len(m["bits"])
    own = np.arange(n) % world == rank
    det.insert_batch_owned(frames, det.project(m["bits"][own]), m["landmarks"])
    det.set_landmark_positions(m["landmark_xyz"])
    det.comm_init_torch()
    assert det.comm_nccl_version() > 0
    # ragged slices of the 8 query keyframes: rank r gets a different number of them
    cuts = np.linspace(0, 8, world + 1).astype(int)
    cuts[1:-1] += 1 if world > 1 else 0
    f0, f1 = int(cuts[rank]), int(cuts[rank + 1])
    qf = frames_of(q["frames"])
    nd = np.concatenate([[0], np.cumsum(qf["num_descriptors"])])
    d0, d1 = int(nd[f0]), int(nd[f1])
    cam = synthetic.camera_dict()
    cams = capi.make_cameras([cam])
    ocams = [po.make_camera(cam["fu"], cam["fv"], cam["cu"], cam["cv"])]
    exp = po.query_batch(ora, qf[f0:f1], q["bits"][d0:d1], q["keypoints"][d0:d1], m["landmark_xyz"], ocams)
    for attempt in range(2):  # the second call reuses every buffer
        out = det.sharded_query_batch(qf[f0:f1].copy(), q["bits"][d0:d1], q["keypoints"][d0:d1], cams,
                                      want_matches=True)
        res = out["results"]
        for f in ("accepted", "num_inliers", "iterations", "ransac_success"):
            assert np.array_equal(res[f], exp[f]), (rank, f)
        assert np.array_equal(np.diff(out["offsets"]), exp["num_matches"])
        ok = exp["ransac_success"].astype(bool)
        assert np.array_equal(res["T_G_I"].reshape(-1, 3, 4)[ok], exp["T"][ok])
    st = det.last_scan_stats()
    assert st["entries"] > 0 and st["scan_ms"] > 0
    # device-input variant + kNN lists against the oracle
    bits_d = torch.from_numpy(q["bits"][d0:d1]).cuda()
    kp_d = torch.from_numpy(np.ascontiguousarray(q["keypoints"][d0:d1], np.float64)).cuda()
    out2 = det.sharded_query_batch_device(qf[f0:f1].copy(), bits_d.data_ptr(), 64, kp_d.data_ptr(), cams)
    assert out2["results"].tobytes() == out["results"].tobytes()
    qp = ora.project(q["bits"][d0:d1])
    qp_d = torch.from_numpy(qp).cuda()
    idx_d = torch.empty((len(qp), 6), dtype=torch.int32, device="cuda")
    dist_d = torch.empty((len(qp), 6), dtype=torch.float32, device="cuda")
    det.sharded_knn_device(qp_d.data_ptr(), len(qp), 6, idx_d.data_ptr(), dist_d.data_ptr())
    oi, od = ora.knn(qp, 6)
    assert np.array_equal(idx_d.cpu().numpy(), oi) and np.array_equal(dist_d.cpu().numpy(), od)
    # a step in which one rank has nothing to ask
    if rank == world - 1:
        out3 = det.sharded_query_batch(qf[:0].copy(), q["bits"][:0], q["keypoints"][:0], cams)
        assert len(out3["results"]) == 0
    else:
        out3 = det.sharded_query_batch(qf[f0:f1].copy(), q["bits"][d0:d1], q["keypoints"][d0:d1], cams)
        assert out3["results"].tobytes() == out["results"].tobytes()
    accepted = torch.tensor([int(res["accepted"].sum())], device="cuda")
    dist.all_reduce(accepted)
    det.comm_destroy()
    dist.barrier()
    if rank == 0:
        print(f"SHARDED_OK world={world} accepted={int(accepted.item())}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
