import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import torch
from maplab_b200 import capi, synthetic
from oracle import pyoracle as po
n = 2_000_000
rng = np.random.default_rng(0)
bits = rng.integers(0, 256, (n, 64), dtype=np.uint8)
blob, _ = synthetic.make_vocabulary(bits[:20000], num_words=64, seed=7)
det = capi.Detector(blob)
ora = po.Engine(blob)
bits_d = torch.from_numpy(bits).cuda()
out_d = torch.empty((n, 10), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
ref = None
for rep in range(4):
    out_d.zero_()
    det.project_device(bits_d.data_ptr(), 64, n, out_d.data_ptr(), st)
    torch.cuda.synchronize()
    o = out_d.cpu().numpy()
    if ref is None:
        import threading
        ref = np.empty_like(o)
        chunks = [(s, min(s + 65536, n)) for s in range(0, n, 65536)]
        def work(t):
            for ci in range(t, len(chunks), 16):
                s, e = chunks[ci]; ref[s:e] = ora.project(bits[s:e])
        ts = [threading.Thread(target=work, args=(t,)) for t in range(16)]
        [t.start() for t in ts]; [t.join() for t in ts]
    bad = np.nonzero((o != ref).any(1))[0]
    print("rep", rep, "bad rows", len(bad))
    if len(bad):
        tiles = bad // 128
        print(" rows in tile", np.bincount(bad % 128, minlength=128).nonzero()[0][:40])
        print(" tiles", np.unique(tiles)[:20], "cta", np.unique(tiles % 148)[:20], "iter", np.unique(tiles // 148)[:20])
        r = bad[0]; print(" row", r, "got", o[r], "exp", ref[r], "dims bad", np.nonzero(o[r] != ref[r])[0])
        # rows bad per tile
        print(" bad per tile", np.bincount(tiles)[np.unique(tiles)][:20])
