"""Shared fixtures for the parity tests: seeded synthetic maps fed identically to the CUDA path
(through the C-ABI) and to the CPU oracle."""
import functools

import numpy as np

from maplab_b200 import capi, synthetic
from oracle import pyoracle as po


@functools.lru_cache(maxsize=8)
def small_world(num_landmarks=3000, num_words=64, seed=1, num_queries=6, flip_log2=6, num_missions=1):
    m = synthetic.make_map(num_landmarks, seed=seed, flip_log2=flip_log2, num_missions=num_missions)
    blob, voc = synthetic.make_vocabulary(m["bits"][:20000], num_words=num_words, seed=seed + 100)
    q = synthetic.make_queries(m, num_queries, seed=seed + 200, flip_log2=flip_log2)
    return m, blob, voc, q


def frames_of(fr):
    return capi.make_frames(fr["timestamp_ns"], fr["vertex_id"], fr["mission_id"], fr["frame_index"],
                            fr["num_descriptors"])


def fill_oracle(ora, frames, proj, landmarks):
    at = 0
    for i in range(len(frames)):
        n = int(frames["num_descriptors"][i])
        ora.insert(int(frames["timestamp_ns"][i]), int(frames["vertex_id"][i]),
                   int(frames["frame_index"][i]), int(frames["mission_id"][i]), proj[at:at + n],
                   landmarks[at:at + n])
        at += n


def oracle_settings(**kw):
    return po.default_settings(**kw)
