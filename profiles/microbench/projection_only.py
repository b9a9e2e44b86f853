#!/usr/bin/env python
"""Kernel 1 alone on resident descriptors (for ncu captures and quick A/B timing).
    python profiles/microbench/projection_only.py [n] [bytes_per_descriptor]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    from maplab_b200 import capi, synthetic
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    nbytes = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 256, (n, nbytes), dtype=np.uint8)
    blob, _ = synthetic.make_vocabulary(bits[:20000], num_words=64, seed=7, desc_bits=8 * nbytes)
    det = capi.Detector(blob)
    bits_d = torch.from_numpy(bits).cuda()
    out_d = torch.empty((n, det.dim), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(6):
        e0.record()
        det.project_device(bits_d.data_ptr(), nbytes, n, out_d.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            best = min(best, e0.elapsed_time(e1))
    print(f"{n} descriptors x {nbytes} B: {best:.4f} ms, {n / best / 1e6:.2f} G descriptors/s, "
          f"{n * (nbytes + 4 * det.dim) / best / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
