#!/usr/bin/env python
"""Experiment: does the GPU have room for a second query step next to the first one? T host threads, each
with its OWN detector (full copy of the 1 M-landmark index) and 1000 / T query keyframes, run their steps
concurrently on one B200 (ctypes releases the GIL inside the library). Prints keyframes/s for T = 1, 2, 3, 4.
    python profiles/microbench/concurrent_detectors.py"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from maplab_b200 import capi, synthetic
    m, blob, q = bench.build_world(1_000_000, 1000, 1000)
    cams = capi.make_cameras([synthetic.camera_dict()])
    dets = []
    proj = None
    for t in range(4):
        det = capi.Detector(blob, capi.default_settings(device=0))
        if proj is None:
            frames, proj, _ = bench.load_database(det, m)
        else:
            det.insert_batch(frames, proj, m["landmarks"])
            det.initialize()
            det.set_landmark_positions(m["landmark_xyz"])
        dets.append(det)
    qframes = bench.frames_array(q["frames"])
    bits_d = torch.from_numpy(q["bits"]).cuda()
    kp_d = torch.from_numpy(np.ascontiguousarray(q["keypoints"], np.float64)).cuda()
    out = {}
    for T in (1, 2, 3, 4):
        per = 1000 // T
        slices = [(t * per, (t + 1) * per if t < T - 1 else 1000) for t in range(T)]

        def work(t, reps):
            f0, f1 = slices[t]
            for _ in range(reps):
                dets[t].query_batch_device(qframes[f0:f1].copy(), bits_d[f0 * 500:].data_ptr(), 64,
                                           kp_d[f0 * 500:].data_ptr(), cams)

        for reps, timed in ((3, False), (20, True)):
            th = [threading.Thread(target=work, args=(t, reps)) for t in range(T)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            [x.start() for x in th]
            [x.join() for x in th]
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if timed:
                out[f"threads_{T}"] = {"keyframes_per_s": 1000 * reps / dt, "ms_per_1000_keyframes": 1e3 * dt / reps}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
