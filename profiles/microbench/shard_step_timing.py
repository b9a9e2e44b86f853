"""Per-phase CUDA-event times of the sharded query step (maplab_b200/sharded.py) under torchrun, N ranks:
    python -m torch.distributed.run --nproc-per-node N profiles/microbench/shard_step_timing.py
N = 2 on B200 (2 M-landmark map, 500 keyframes per rank), ms: project 0.03, coarse 0.59, 2 x all-gather 0.09,
scan (all 500 k queries) 0.38, 2 x all-to-all 0.08, merge 0.02, voting + RANSAC of the slice ~1.7 — the
exchanges are 6 % of the step; the per-rank latency floors of voting / RANSAC bound the scaling."""
import os, sys, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
import bench
from maplab_b200 import capi, sharded, synthetic
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev=torch.device("cuda",local)
dist.init_process_group("nccl", device_id=dev)
m,blob,q=bench.build_world(1_000_000*world,1000,1000)
det=capi.Detector(blob,capi.default_settings(device=local,shard_rank=rank,shard_count=world))
bench.load_database(det,m)
k=det.num_neighbors(); cams=capi.make_cameras([synthetic.camera_dict()])
qframes=bench.frames_array(q["frames"])
ops=sharded.DetectorOps(det,cams)
step=sharded.ShardedQueryStep(ops,qframes,rank,world,det.dim,10,k,64,dev)
bits=torch.from_numpy(q["bits"]).to(dev); kp=torch.from_numpy(np.ascontiguousarray(q["keypoints"],np.float64)).to(dev)
sb,sk=step.slice_of(bits),step.slice_of(kp)
for _ in range(3): step.run(sb,sk)
names=["project","coarse","allgather","scan","alltoall","merge","verify"]
acc=np.zeros(len(names))
G,n_s=world,step.n_s
for it in range(10):
    dist.barrier(); torch.cuda.synchronize()
    ev=[torch.cuda.Event(enable_timing=True) for _ in range(len(names)+1)]
    o=step.ops
    ev[0].record(); o.project(sb,step.proj_s)
    ev[1].record(); o.coarse(step.proj_s,step.cells_s)
    ev[2].record()
    dist.all_gather_into_tensor(step.proj_all.view(G*n_s,-1),step.proj_s); dist.all_gather_into_tensor(step.cells_all.view(G*n_s,-1),step.cells_s)
    ev[3].record(); o.scan(step.proj_all.view(G*n_s,-1),step.cells_all.view(G*n_s,-1),step.pidx.view(G*n_s,-1),step.pdist.view(G*n_s,-1))
    ev[4].record()
    dist.all_to_all_single(step.ridx.view(G*n_s,-1),step.pidx.view(G*n_s,-1)); dist.all_to_all_single(step.rdist.view(G*n_s,-1),step.pdist.view(G*n_s,-1))
    ev[5].record(); o.merge(step.ridx,step.rdist,step.midx,step.mdist)
    ev[6].record()
    t0=time.perf_counter(); o.verify(step.frames,step.midx,step.mdist,sk); torch.cuda.synchronize(); tv=(time.perf_counter()-t0)*1e3
    for i in range(6): acc[i]+=ev[i].elapsed_time(ev[i+1])
    acc[6]+=tv
if rank==0: print(json.dumps(dict(zip(names,(acc/10).round(4).tolist()))))
dist.destroy_process_group()
