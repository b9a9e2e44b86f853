"""Memory-safety / agreement fuzz of the two wire decoders (host only): random mutations of valid messages must
either decode or be rejected with an error — never crash — and whenever the protobuf runtime rejects the bytes,
so does the library (the converse need not hold: the library also applies the reference's CHECKs)."""
import numpy as np
import pytest
from google.protobuf.message import DecodeError
from hypothesis import HealthCheck, given, settings, strategies as st

import summary_map_proto as smp
from maplab_b200 import capi, vi_map_io
from test_summary_map import random_map
from test_vi_map_io import _toy_map

SUMMARY = smp.encode(**random_map(3, L=4, O=3, N=9, D=5))
SUMMARY_PACKED = smp.encode(packed=True, **random_map(4, L=4, O=3, N=9, D=5))
VERTICES = _toy_map(np.random.default_rng(5), vertices=2, frames=2, bytes_per_desc=16).SerializeToString()


def mutate(blob, edits):
    b = bytearray(blob)
    for kind, pos, val in edits:
        if not b:
            break
        pos %= len(b)
        if kind == 0:
            b[pos] = val
        elif kind == 1:
            del b[pos]
        elif kind == 2:
            b.insert(pos, val)
        else:
            del b[pos:]
    return bytes(b)


EDITS = st.lists(st.tuples(st.integers(0, 3), st.integers(0, 1 << 20), st.integers(0, 255)), min_size=1, max_size=6)
SETTINGS = dict(max_examples=400, deadline=None, suppress_health_check=[HealthCheck.too_slow])


def runtime_accepts(cls, blob):
    try:
        cls().ParseFromString(blob)
        return True
    except DecodeError:
        return False


@settings(**SETTINGS)
@given(edits=EDITS, packed=st.booleans())
def test_summary_map_decoder_survives_mutations(edits, packed):
    blob = mutate(SUMMARY_PACKED if packed else SUMMARY, edits)
    try:
        got = capi.summary_map_parse(blob)
        ok = True
    except capi.MlcError:
        ok = False
    if ok:
        assert runtime_accepts(smp.messages(), blob)
        exp = smp.decode(blob)
        assert np.array_equal(got["observer_indices"], exp["observer_indices"])
        assert np.array_equal(got["descriptors"].view(np.uint32), exp["descriptors"].view(np.uint32))


@settings(**SETTINGS)
@given(edits=EDITS)
def test_vi_map_decoder_survives_mutations(edits):
    blob = mutate(VERTICES, edits)
    try:
        got = capi.vi_map_read_vertices(blob)
        ok = True
    except capi.MlcError:
        ok = False
    if ok:
        cls = vi_map_io._vi_map_class()
        assert runtime_accepts(cls, blob)
        msg = cls()
        msg.ParseFromString(blob)
        assert len(msg.vertices) == len(got["vertex_num_frames"])
        frames = [f for v in msg.vertices for f in v.n_visual_frame.frames]
        assert got["frame_timestamp_ns"].tolist() == [f.timestamp for f in frames]


@pytest.mark.parametrize("size", [0, 1, 7, 64, 4096])
def test_random_bytes(size):
    rng = np.random.default_rng(size)
    for _ in range(200):
        blob = rng.integers(0, 256, size, dtype=np.uint8).tobytes()
        for fn in (capi.summary_map_parse, capi.vi_map_read_vertices):
            try:
                fn(blob)
            except capi.MlcError:
                pass
