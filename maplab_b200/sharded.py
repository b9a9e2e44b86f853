"""Host-side mirror of the sharded loop-closure query step over torch.distributed.

The PRODUCT path of this step is inside the library: mlc_comm_* + mlc_sharded_query_batch (csrc/sharded.cu, NCCL
bound at run time; capi.Detector.sharded_query_batch) — bench.py and the multi-GPU tests go through that. This
module keeps the same schedule in Python so that the exchange logic can be exercised on CPU at world size 2 over
gloo (tests/test_sharded_gloo.py, with the oracle as per-rank compute), and as a reference for hosts that drive
the per-stage entry points themselves (mlc_coarse_device / mlc_scan_device / mlc_merge_topk_device).

One process per GPU, the inverted lists of the database sharded over the ranks (descriptor i lives on rank
i % G), everything else replicated.

Per step every rank
  1. projects and coarse-searches ITS slice of the query keyframes (kernels 1, 2a),
  2. all-gathers (projected query, visit list)            — exchange 1: every shard needs all queries,
  3. scans its own lists for ALL queries (kernel 2b, one launch),
  4. all-to-alls the per-shard top-k lists                — exchange 2: rank r receives, from every
     shard, the lists of r's slice,
  5. merges them by (distance, index) (merge_topk kernel) — equal to the single-index result because
     the visited cells depend only on query + vocabulary and global descriptor indices keep the
     tie-breaks global,
  6. runs voting / clustering / RANSAC (kernels 3, 4) on its slice.
Nothing else crosses GPUs. The exchanges go through torch.distributed (NCCL on the GPU box; the
same code runs over gloo with CPU tensors in tests/test_sharded_gloo.py, where the per-rank
compute `ops` is the CPU oracle instead of the CUDA library).

`ops` (duck-typed) works on torch tensors of the step's device:
    project(bits[n,B] u8, out[n,dim] f32); coarse(proj[n,dim], cells[n,nw] i32);
    scan(proj[n,dim], cells[n,nw], idx[n,k] i32, dist[n,k] f32);
    merge(idx_lists[G,n,k], dist_lists[G,n,k], idx[n,k], dist[n,k]);
    verify(frames, idx[n,k], dist[n,k], keypoints[n,2] f64) -> result of the slice
"""
import numpy as np
import torch
import torch.distributed as dist


def query_slice(rank, world, num_frames):
    """Contiguous, equal slices of the query keyframes: [f0, f1) of rank `rank`."""
    if num_frames % world:
        raise ValueError("the query batch must divide evenly over the ranks")
    per = num_frames // world
    return rank * per, (rank + 1) * per


class DetectorOps:
    """`ops` backed by the CUDA library (capi.Detector built with shard_rank / shard_count)."""

    def __init__(self, det, cams, stream=None):
        self.det, self.cams = det, cams
        self.stream = stream if stream is not None else torch.cuda.current_stream().cuda_stream

    def project(self, bits, out):
        self.det.project_device(bits.data_ptr(), bits.shape[1], bits.shape[0], out.data_ptr(), self.stream)

    def coarse(self, proj, cells):
        self.det.coarse_device(proj.data_ptr(), proj.shape[0], cells.shape[1], cells.data_ptr(), self.stream)

    def scan(self, proj, cells, idx, dst):
        self.det.scan_device(proj.data_ptr(), cells.data_ptr(), proj.shape[0], idx.shape[1],
                             idx.data_ptr(), dst.data_ptr(), self.stream)

    def merge(self, idx_lists, dist_lists, idx, dst):
        self.det.merge_topk_device(idx_lists.data_ptr(), dist_lists.data_ptr(), idx_lists.shape[0],
                                   idx.shape[0], idx.shape[1], idx.data_ptr(), dst.data_ptr(), self.stream)

    def verify(self, frames, idx, dst, keypoints):
        torch.cuda.current_stream().synchronize()  # kernels 3/4 run on the detector's own stream
        return self.det.query_from_knn_device(frames, idx.data_ptr(), dst.data_ptr(), idx.shape[1],
                                              keypoints.data_ptr(), self.cams)


class ShardedQueryStep:
    """Buffers + schedule of one rank. All query frames must hold the same number of descriptors
    (equal slices keep the collectives regular)."""

    def __init__(self, ops, frames, rank, world, dim, nw, k, desc_bytes, device):
        nd = np.asarray(frames["num_descriptors"])
        if len(set(nd.tolist())) != 1:
            raise ValueError("sharded step needs the same number of descriptors in every query frame")
        self.ops, self.rank, self.world = ops, rank, world
        self.f0, self.f1 = query_slice(rank, world, len(frames))
        self.frames = frames[self.f0:self.f1].copy()
        per = int(nd[0])
        self.n_s = (self.f1 - self.f0) * per       # query descriptors of this rank's slice
        self.d0 = self.f0 * per                    # first query descriptor of the slice
        n_s, G = self.n_s, world
        f32, i32 = torch.float32, torch.int32
        self.proj_s = torch.empty((n_s, dim), dtype=f32, device=device)
        self.cells_s = torch.empty((n_s, nw), dtype=i32, device=device)
        self.proj_all = torch.empty((G, n_s, dim), dtype=f32, device=device)
        self.cells_all = torch.empty((G, n_s, nw), dtype=i32, device=device)
        self.pidx = torch.empty((G, n_s, k), dtype=i32, device=device)     # my shard's lists, by owner
        self.pdist = torch.empty((G, n_s, k), dtype=f32, device=device)
        self.ridx = torch.empty_like(self.pidx)                            # my slice's lists, by shard
        self.rdist = torch.empty_like(self.pdist)
        self.midx = torch.empty((n_s, k), dtype=i32, device=device)
        self.mdist = torch.empty((n_s, k), dtype=f32, device=device)
        self.desc_bytes = desc_bytes

    def slice_of(self, per_descriptor_tensor):
        return per_descriptor_tensor[self.d0:self.d0 + self.n_s]

    def knn(self, bits_slice):
        """Steps 1-5: merged kNN lists (midx, mdist) of this rank's slice."""
        o, G, n_s = self.ops, self.world, self.n_s
        o.project(bits_slice, self.proj_s)
        o.coarse(self.proj_s, self.cells_s)
        if G > 1:
            # flat [G * n_s, .] views: the layout both NCCL and gloo accept for the fused all-gather
            dist.all_gather_into_tensor(self.proj_all.view(G * n_s, -1), self.proj_s)
            dist.all_gather_into_tensor(self.cells_all.view(G * n_s, -1), self.cells_s)
        else:
            self.proj_all[0].copy_(self.proj_s)
            self.cells_all[0].copy_(self.cells_s)
        # [G, n_s, .] is the whole batch in rank order: one scan launch over all G * n_s queries
        o.scan(self.proj_all.view(G * n_s, -1), self.cells_all.view(G * n_s, -1),
               self.pidx.view(G * n_s, -1), self.pdist.view(G * n_s, -1))
        if G > 1:
            dist.all_to_all_single(self.ridx.view(G * n_s, -1), self.pidx.view(G * n_s, -1))
            dist.all_to_all_single(self.rdist.view(G * n_s, -1), self.pdist.view(G * n_s, -1))
        else:
            self.ridx.copy_(self.pidx)
            self.rdist.copy_(self.pdist)
        o.merge(self.ridx, self.rdist, self.midx, self.mdist)
        return self.midx, self.mdist

    def run(self, bits_slice, keypoints_slice):
        """The whole step for this rank's slice; returns ops.verify's result."""
        idx, dst = self.knn(bits_slice)
        return self.ops.verify(self.frames, idx, dst, keypoints_slice)
