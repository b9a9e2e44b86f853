"""The C++ host shim (include/maplab_lc_b200_shim.h: LoopDetector with the reference's method names)
compiled with plain g++ against the C-ABI library. CPU: it compiles, links and fails loudly without
a device. GPU: a C++ program drives Insert / Initialize / QueryBatch and must report exactly what the
Python binding reports for the same world."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from maplab_b200 import capi, synthetic
from helpers import frames_of, small_world

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


def _build(tmp_path):
    exe = tmp_path / "shim_program"
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "shim_program.cc"), "-o", str(exe),
                           "-L", libdir, "-lmaplab_lc_b200", f"-Wl,-rpath,{libdir}",
                           "-L", "/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    return exe


def _write_world(path, m, blob, q):
    dbf, qf = frames_of(m["frames"]), frames_of(q["frames"])
    head = np.array([len(dbf), len(m["bits"]), len(qf), len(q["bits"]), 10, m["bits"].shape[1], len(blob),
                     len(m["landmark_xyz"])], np.int64)
    with open(path, "wb") as f:
        for part in (head, np.frombuffer(bytes(blob), np.uint8), dbf, m["bits"], m["landmarks"].astype(np.int64),
                     np.ascontiguousarray(m["landmark_xyz"], np.float64), qf, q["bits"],
                     np.ascontiguousarray(q["keypoints"], np.float64)):
            f.write(np.ascontiguousarray(part).tobytes())


def test_shim_compiles_links_and_fails_loudly_without_a_device(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    m, blob, _, q = small_world(num_queries=4)
    _write_world(tmp_path / "world.bin", m, blob, q)
    r = subprocess.run([str(exe), str(tmp_path / "world.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "shim:" in r.stderr          # no CPU fallback


@pytest.mark.gpu
def test_shim_program_matches_python_binding(tmp_path):
    exe = _build(tmp_path)
    m, blob, _, q = small_world(num_queries=12)
    _write_world(tmp_path / "world.bin", m, blob, q)
    out = subprocess.check_output([str(exe), str(tmp_path / "world.bin")], text=True).split("\n")
    det = capi.Detector(blob, capi.default_settings(num_nearest_neighbors=6))
    det.insert_batch(frames_of(m["frames"]), det.project(m["bits"]), m["landmarks"])
    det.set_landmark_positions(m["landmark_xyz"])
    ref = det.query_batch(frames_of(q["frames"]), q["bits"], q["keypoints"],
                          capi.make_cameras([synthetic.camera_dict()]), want_matches=True, want_flags=True)
    res = ref["results"]
    nv, acc, ninl, ndesc = (int(x) for x in out[0].split())
    assert nv == len(res) and acc == int(res["accepted"].sum()) and ndesc == len(m["bits"])
    off = ref["offsets"]
    exp_inl = sum(int((ref["inlier_flags"][off[v]:off[v + 1]] == 3).sum()) for v in range(nv) if res["accepted"][v])
    assert ninl == exp_inl and acc > 0
    for v in range(nv):
        a, n_in, it, tx, ty, tz = out[1 + v].split()
        assert (int(a), int(n_in), int(it)) == (int(res["accepted"][v]), int(res["num_inliers"][v]),
                                                int(res["iterations"][v]))
        T = res["T_G_I"][v].reshape(3, 4)
        assert (float(tx), float(ty), float(tz)) == (T[0, 3], T[1, 3], T[2, 3])


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2])
def test_shim_program_multi_gpu_equals_single_gpu(tmp_path, world):
    """A C++ host per GPU — no Python, no torch — builds its shard, joins the NCCL communicator through the
    C-ABI (mlc_comm_*) and answers its slice with ShardedQueryBatch: the slices together must be the
    single-GPU program's output."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    exe = _build(tmp_path)
    m, blob, _, q = small_world(num_queries=12)
    _write_world(tmp_path / "world.bin", m, blob, q)
    single = subprocess.check_output([str(exe), str(tmp_path / "world.bin")], text=True).split("\n")
    procs = [subprocess.Popen([str(exe), str(tmp_path / "world.bin"), str(r), str(world), str(tmp_path / "comm.id")],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=300)
        assert p.returncode == 0, e[-2000:]
        outs.append([ln for ln in o.split("\n") if not ln.startswith("NCCL version")])  # NCCL's banner is on stdout
    lines = [ln for o in outs for ln in o[1:] if ln]
    assert lines == [ln for ln in single[1:] if ln]
    assert sum(int(o[0].split()[1]) for o in outs) == int(single[0].split()[1])      # accepted
    assert all(int(o[0].split()[3]) == int(single[0].split()[3]) for o in outs)      # NumDescriptors: whole database
