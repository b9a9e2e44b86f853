#!/usr/bin/env python
"""Copies maplab's shipped FREAK quantizer (a trained vocabulary: projection matrix 10x512 + two
5x1000 word sets, written by InvertedMultiIndexVocabulary::Save) into tests/golden/ as a data
fixture for the realism checks of tests/test_real_vocabulary.py. Run in the build container, where
the reference checkout is mounted; the GPU box only sees the committed copy.

    python tests/golden/fetch_shipped_vocabulary.py
"""
import os
import shutil

SRC = ("/root/reference/algorithms/loopclosure/matching-based-loopclosure/share/"
       "inverted_multi_index_quantizer_freak.dat")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "inverted_multi_index_quantizer_freak.dat")

if __name__ == "__main__":
    shutil.copyfile(SRC, DST)
    print(f"{DST}: {os.path.getsize(DST)} bytes")
