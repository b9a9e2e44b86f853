#!/usr/bin/env python
"""Builds tests/golden/real_map_brisk.npz + brisk_quantizer_top10.dat from maplab's own test map
(tools/maplab-test-data/test_maps/common_test_map: 127 vertices of a 5-camera Alphasense rig, real
BRISK descriptors, triangulated landmarks) and its shipped BRISK vocabulary. Run in the build
container, where the reference checkout is mounted; the GPU box only sees the committed outputs.

    python tests/golden/extract_real_map.py

What is kept is what `lc` feeds the loop detector (LoopDetectorNode::addVertexToDatabase /
convertFrameToProjectedImage, LCH/src/loop-detector-node.cc:119-200, :273-339): per visual frame the
keypoints with a valid landmark id whose landmark is well constrained (quality kGood — all of this
map's landmarks are), their raw descriptors, measurements and landmark ids; landmark positions in the
global frame (T_G_M * T_M_I(storing vertex) * p_B, vi_map::VIMap::getLandmark_G_p); vertex poses;
the cameras of sensors.yaml. Only cameras CAMS are kept to bound the fixture size.
The map files are gzip'd proto2 messages (vi-map/proto/vi-map/vi_map.proto,
aslam-serialization/visual-frame.proto); maplab_b200/vi_map_io.py reads them with the protobuf runtime
(no protoc in the image).
The quantizer file holds the full 384 x 384 BRISK projection; the path uses its first
lc_target_dimensionality = 10 rows (descriptor-projection.cc:45-49), so only those are stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
MAP = "/root/reference/tools/maplab-test-data/test_maps/common_test_map/vi_map/"
QUANTIZER = ("/root/reference/algorithms/loopclosure/matching-based-loopclosure/share/"
             "inverted_multi_index_quantizer_brisk.dat")
CAMS = (0, 2)
def main():
    from maplab_b200 import vi_map_io
    vi_map = vi_map_io.load_vi_map(MAP)
    x = vi_map_io.loop_closure_inputs(vi_map, CAMS)
    used = np.unique(x["landmarks"])  # renumber to the landmarks the kept cameras observe
    renumber = np.full(len(x["landmark_xyz"]), -1, np.int64)
    renumber[used] = np.arange(len(used))
    cams = [np.concatenate([[c["fu"], c["fv"], c["cu"], c["cv"]], c["dist"],
                            np.vstack([np.hstack([c["R_B_C"], c["t_B_C"][:, None]]), [0, 0, 0, 1]]).ravel()])
            for c in vi_map_io.cameras_of(vi_map["sensors"], CAMS)]
    assert all(c["distortion"] == 3 for c in vi_map_io.cameras_of(vi_map["sensors"], CAMS))  # equidistant
    out = os.path.join(HERE, "real_map_brisk.npz")
    np.savez_compressed(
        out, frames=x["frames"], bits=x["bits"], keypoints=x["keypoints"],
        landmarks=renumber[x["landmarks"]].astype(np.int32), landmark_xyz=x["landmark_xyz"][used],
        T_G_I=x["T_G_I"], cameras=np.stack(cams))
    print(out, os.path.getsize(out), "bytes;", len(x["frames"]), "frames,", len(x["landmarks"]), "descriptors,",
          len(used), "landmarks")

    from maplab_b200 import synthetic
    v = synthetic.parse_vocabulary(open(QUANTIZER, "rb").read())
    print("BRISK quantizer: version", v["version"], "target_dim", v["target_dim"], "P", v["P"].shape, "W1",
          v["W1"].shape, "W2", v["W2"].shape)
    top = synthetic.serialize_vocabulary(np.ascontiguousarray(v["P"][:v["target_dim"]]), v["W1"], v["W2"],
                                         v["target_dim"], v["version"])
    dst = os.path.join(HERE, "brisk_quantizer_top10.dat")
    open(dst, "wb").write(top)
    print(dst, len(top), "bytes")


if __name__ == "__main__":
    main()
