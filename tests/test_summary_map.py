"""Localization summary map file format: the library's own proto2 wire decoder /
encoder (mlc_summary_map_parse / _serialize; no libprotobuf in the build) against the REAL protobuf
runtime on the reference's two messages (tests/summary_map_proto.py), incl. packed / unpacked
repeated fields, unknown fields, merged duplicate sub-messages, truncation; the reference's own
round-trip test (map-structure/localization-summary-map/test/
test_localization_summary_map_protobuf_test.cc: serialize -> deserialize -> equal) on random maps.
Host-only: runs without a GPU."""
import numpy as np
import pytest
from google.protobuf.message import DecodeError

import summary_map_proto as smp
from maplab_b200 import capi

KEYS = ("G_landmark_position", "G_observer_position", "descriptors", "observer_indices",
        "observation_to_landmark_index")


def random_map(seed, L=9, O=4, N=37, D=10, big_ids=False):
    rng = np.random.default_rng(seed)
    return dict(
        G_landmark_position=rng.normal(size=(3, L)).astype(np.float32) * 50,
        G_observer_position=rng.normal(size=(3, O)).astype(np.float32),
        descriptors=rng.normal(size=(D, N)).astype(np.float32),
        observer_indices=rng.integers(0, max(O, 1), N).astype(np.uint32),
        observation_to_landmark_index=(rng.integers(0, 2 ** 32, N) if big_ids
                                       else rng.integers(0, max(L, 1), N)).astype(np.uint32))


def same(a, b):
    return all(np.array_equal(a[k], b[k]) and a[k].shape == b[k].shape for k in KEYS)


@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("shape", [(9, 4, 37, 10), (1, 1, 1, 1), (300, 20, 2000, 10), (5, 2, 0, 10)])
def test_decoder_matches_libprotobuf(packed, shape):
    L, O, N, D = shape
    m = random_map(L + N, L, O, N, D, big_ids=True)  # ids up to 2^32-1: 5-byte varints
    blob = smp.encode(packed=packed, **m)
    got = capi.summary_map_parse(blob)
    assert same(got, m)
    assert same(got, smp.decode(blob))


@pytest.mark.parametrize("seed", range(3))
def test_encoder_is_byte_identical_to_libprotobuf_and_round_trips(seed):
    m = random_map(seed, L=50 + seed, O=3 + seed, N=400)
    blob = capi.summary_map_serialize(**m)
    assert blob == smp.encode(**m)  # proto2, no [packed=true]: one key per element
    assert same(capi.summary_map_parse(blob), m)  # the reference's SerializeAndDeserialize test


def test_special_float_values_keep_their_bits():
    m = random_map(5)
    bits = np.array([0x7fc00001, 0xff800000, 0x80000000, 0x00000001], np.uint32)
    m["descriptors"][0, :4] = bits.view(np.float32)
    got = capi.summary_map_parse(capi.summary_map_serialize(**m))
    assert np.array_equal(got["descriptors"].view(np.uint32), m["descriptors"].view(np.uint32))


def test_unknown_fields_and_groups_are_skipped_and_duplicates_merge():
    a, b = random_map(1), random_map(2, N=11)
    blob_a, blob_b = smp.encode(**a), smp.encode(packed=True, **b)
    unknown = (bytes([(15 << 3) | 0, 0xAC, 0x02])            # varint field 15
               + bytes([(9 << 3) | 1]) + bytes(8)             # fixed64 field 9
               + bytes([(7 << 3) | 2, 3, 1, 2, 3])            # bytes field 7
               + bytes([(6 << 3) | 3, (1 << 3) | 0, 5, (6 << 3) | 4])  # group field 6 { 1: 5 }
               + bytes([(2 << 3) | 5]) + bytes(4))            # field 2 with a foreign wire type
    assert same(capi.summary_map_parse(blob_a + unknown), a)
    # concatenation == MergeFrom: repeated fields append, optional rows / cols: the last one wins ...
    both = blob_a + unknown + blob_b
    assert smp.decode(smp.encode(**random_map(4, N=0)) + blob_b)["descriptors"].shape == b["descriptors"].shape
    empty = random_map(4, N=0)
    merged = smp.encode(**empty) + unknown + blob_b
    got, exp = capi.summary_map_parse(merged), smp.decode(merged)
    assert same(got, exp)
    assert got["G_landmark_position"].shape[1] == 18 and got["descriptors"].shape == (10, 11)
    # ... so with data on both sides rows * cols != data_size: the CHECK_EQ of eigen_proto::deserialize
    with pytest.raises(capi.MlcError, match="rows \\* cols"):
        capi.summary_map_parse(both)


def test_every_truncation_fails_or_parses_like_libprotobuf():
    m = random_map(3, L=3, O=2, N=5, D=4)
    blob = smp.encode(**m)
    cls = smp.messages()
    for cut in range(len(blob)):
        msg = cls()
        try:
            msg.ParseFromString(blob[:cut])
            ok = True
        except DecodeError:
            ok = False
        if not ok:
            with pytest.raises(capi.MlcError, match="malformed"):
                capi.summary_map_parse(blob[:cut])
        elif not msg.HasField("uncompressed_map"):
            # deserialize() reads G_landmark_position (CHECK % 3) before it looks for the sub-message
            why = "Unsupported localization summary map format" if len(msg.G_landmark_position) % 3 == 0 \
                else "not 3 x L"
            with pytest.raises(capi.MlcError, match=why):
                capi.summary_map_parse(blob[:cut])
        else:
            u = msg.uncompressed_map
            consistent = (len(msg.G_landmark_position) % 3 == 0 and len(u.G_observer_position) % 3 == 0
                          and u.descriptors.rows * u.descriptors.cols == len(u.descriptors.data))
            if consistent:
                got = capi.summary_map_parse(blob[:cut])
                assert np.array_equal(got["descriptors"].T.ravel(), np.array(u.descriptors.data, np.float32))
            else:
                with pytest.raises(capi.MlcError):
                    capi.summary_map_parse(blob[:cut])


def test_malformed_inputs_are_rejected():
    for bad in (bytes([0x00, 0x01]),                    # field number 0
                bytes([0x0A, 0x05, 1, 2, 3]),           # length runs past the end
                bytes([0x0A, 0x03, 1, 2, 3]),           # packed floats: length not a multiple of 4
                bytes([0x80] * 11),                     # varint longer than 10 bytes
                bytes([(1 << 3) | 4])):                 # stray end-group
        with pytest.raises(capi.MlcError, match="malformed"):
            capi.summary_map_parse(bad)
    with pytest.raises(capi.MlcError, match="Unsupported"):
        capi.summary_map_parse(b"")
    # a megabyte of nested start-group keys must be refused (libprotobuf's recursion limit is 100), not recursed into
    deep = bytes([(6 << 3) | 3]) * (1 << 20)
    with pytest.raises(capi.MlcError, match="malformed"):
        capi.summary_map_parse(deep)
    from maplab_b200 import capi as _c
    with pytest.raises(_c.MlcError):
        _c.vi_map_read_vertices(deep)
    nested_ok = bytes([(6 << 3) | 3]) * 50 + bytes([(6 << 3) | 4]) * 50  # 50 levels: skipped as an unknown field
    with pytest.raises(capi.MlcError, match="Unsupported"):
        capi.summary_map_parse(nested_ok)


def test_file_as_written_with_proto_use_compression(tmp_path):
    # serializeProtoToFile wraps the message into a GzipOutputStream by default: inflate, then parse
    import gzip
    from maplab_b200 import vi_map_io
    m = random_map(11)
    path = tmp_path / "localization_summary_map"
    path.write_bytes(gzip.compress(smp.encode(**m)))
    assert same(capi.summary_map_parse(vi_map_io.read_proto_bytes(str(path))), m)
    path.write_bytes(smp.encode(**m))  # --proto_use_compression=false
    assert same(capi.summary_map_parse(vi_map_io.read_proto_bytes(str(path))), m)
    with pytest.raises(capi.MlcError):
        capi.summary_map_parse(gzip.compress(smp.encode(**m)))  # the gzip stream itself is not a message
