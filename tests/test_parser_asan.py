"""The wire decoders under AddressSanitizer + UBSan: a corpus of mutated / truncated / random blobs is parsed by
a C++ driver (tests/cpp/parser_sanitizer_driver.cc) compiled together with summary_map.cc and vi_map_reader.cc;
any out-of-bounds access aborts the driver. Host only."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

import summary_map_proto as smp
from test_parser_fuzz import SUMMARY, SUMMARY_PACKED, VERTICES, mutate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "maplab_b200", "csrc")


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_decoders_are_clean_under_asan_and_ubsan(tmp_path):
    exe = tmp_path / "driver"
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
           "-I", CSRC, os.path.join(ROOT, "tests", "cpp", "parser_sanitizer_driver.cc"),
           os.path.join(CSRC, "summary_map.cc"), os.path.join(CSRC, "vi_map_reader.cc"), "-o", str(exe)]
    built = subprocess.run(cmd, capture_output=True, text=True)
    if built.returncode != 0 and "sanitize" in built.stderr:
        pytest.skip("this g++ has no sanitizer runtime")
    assert built.returncode == 0, built.stderr
    rng = np.random.default_rng(0)
    blobs = [SUMMARY, SUMMARY_PACKED, VERTICES, b""]
    for base in (SUMMARY, SUMMARY_PACKED, VERTICES):
        blobs += [base[:cut] for cut in range(0, len(base), 3)]
        for _ in range(1500):
            edits = [(int(rng.integers(0, 4)), int(rng.integers(0, 1 << 20)), int(rng.integers(0, 256)))
                     for _ in range(int(rng.integers(1, 6)))]
            blobs.append(mutate(base, edits))
    blobs += [rng.integers(0, 256, int(rng.integers(0, 300)), dtype=np.uint8).tobytes() for _ in range(1000)]
    corpus = tmp_path / "corpus.bin"
    with open(corpus, "wb") as f:
        for b in blobs:
            f.write(struct.pack("<I", len(b)) + b)
    run = subprocess.run([str(exe), str(corpus)], capture_output=True, text=True,
                         env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0"))
    assert run.returncode == 0, run.stderr[-3000:]
    assert f"{len(blobs)} blobs" in run.stdout
