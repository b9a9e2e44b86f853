"""World-size-2 run of the sharded query step (maplab_b200/sharded.py) over gloo on CPU: the slicing,
both exchanges (all-gather of queries / visit lists, all-to-all of the per-shard top-k lists) and the
merge order are the product's code; the per-rank compute is the CPU oracle (test infrastructure),
each rank indexing only the descriptors i % G == rank. The merged lists of every rank's slice must be
bit-identical to the single-index result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

K, NW = 6, 10


class OracleOps:
    """Stand-in for sharded.DetectorOps on CPU tensors."""

    def __init__(self, ora, voc, shard_desc, rank, world):
        from oracle import pyoracle as po
        self.ora, self.rank, self.world = ora, rank, world
        h, w = voc["W1"].shape
        self.imi = po.IMI(po.colmajor(voc["W1"].tolist()), w, po.colmajor(voc["W2"].tolist()),
                          voc["W2"].shape[1], h, NW)
        self.imi.add(shard_desc, len(shard_desc))

    def project(self, bits, out):
        out.copy_(torch.from_numpy(self.ora.project(bits.numpy())))

    def coarse(self, proj, cells):
        cells.copy_(torch.from_numpy(self.imi.visited_cells(proj.numpy(), proj.shape[0])))

    def scan(self, proj, cells, idx, dst):
        assert np.array_equal(cells.numpy(), self.imi.visited_cells(proj.numpy(), proj.shape[0]))
        i, d = self.imi.knn(np.ascontiguousarray(proj.numpy()), proj.shape[0], idx.shape[1])
        i = np.where(i >= 0, i * self.world + self.rank, -1).astype(np.int32)  # local -> global index
        idx.copy_(torch.from_numpy(i))
        dst.copy_(torch.from_numpy(d))

    def merge(self, idx_lists, dist_lists, idx, dst):
        il, dl = idx_lists.numpy(), dist_lists.numpy()
        G, n, k = il.shape
        oi = np.full((n, k), -1, np.int32)
        od = np.full((n, k), np.inf, np.float32)
        for q in range(n):
            cand = sorted((dl[g, q, j].view(np.uint32).item(), int(il[g, q, j]))
                          for g in range(G) for j in range(k) if il[g, q, j] >= 0)[:k]
            for j, (db, ix) in enumerate(cand):
                oi[q, j] = ix
                od[q, j] = np.uint32(db).view(np.float32)
        idx.copy_(torch.from_numpy(oi))
        dst.copy_(torch.from_numpy(od))

    def verify(self, frames, idx, dst, keypoints):
        return idx.numpy().copy(), dst.numpy().copy()


def _worker(rank, world, port, q_out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import frames_of, small_world
        from maplab_b200 import sharded
        from oracle import pyoracle as po
        m, blob, voc, q = small_world(num_queries=8)
        ora = po.Engine(blob)
        proj = ora.project(m["bits"])
        ops = OracleOps(ora, voc, np.ascontiguousarray(proj[rank::world]), rank, world)
        qframes = frames_of(q["frames"])
        step = sharded.ShardedQueryStep(ops, qframes, rank, world, dim=proj.shape[1], nw=NW, k=K,
                                        desc_bytes=q["bits"].shape[1], device="cpu")
        bits = torch.from_numpy(q["bits"])
        kp = torch.from_numpy(np.ascontiguousarray(q["keypoints"], np.float64))
        idx, dst = step.run(step.slice_of(bits), step.slice_of(kp))
        # single-index reference for this rank's slice
        full = OracleOps(ora, voc, proj, 0, 1)
        qp = ora.project(q["bits"])[step.d0:step.d0 + step.n_s]
        ri, rd = full.imi.knn(np.ascontiguousarray(qp), len(qp), K)
        ok = bool(np.array_equal(idx, ri) and np.array_equal(dst, rd) and (idx >= 0).any())
        owners = set((idx[idx >= 0] % world).tolist())
        q_out.put((rank, ok, len(owners), step.f0, step.f1))
    except Exception as e:  # surface the failure instead of leaving the parent waiting
        q_out.put((rank, False, repr(e), -1, -1))
        raise
    finally:
        dist.destroy_process_group()


def test_sharded_step_world2_gloo():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q_out)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q_out.get(timeout=240) for _ in range(world))
    assert all(r[3] >= 0 for r in res), res
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] for r in res), "merged per-shard lists differ from the single-index result"
    assert all(r[2] == world for r in res), "every shard should contribute neighbours"
    assert (res[0][3], res[0][4], res[1][3], res[1][4]) == (0, 4, 4, 8)


def test_query_slice():
    from maplab_b200 import sharded
    assert [sharded.query_slice(r, 4, 1000) for r in range(4)] == [(0, 250), (250, 500), (500, 750), (750, 1000)]
    with pytest.raises(ValueError):
        sharded.query_slice(0, 3, 1000)
