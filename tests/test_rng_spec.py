"""The NumPy restatement of Philox4x32-10 that tests/test_rng_gpu.py checks the device stream against is itself pinned
to the published known-answer vector of Random123 (kat_vectors: philox4x32 10 rounds, zero counter, zero key)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_rng_gpu import philox4x32_10  # noqa: E402


def test_philox_restatement_matches_the_published_known_answer():
    w = philox4x32_10(np.array([0], dtype=np.uint64), np.array([0], dtype=np.uint64), 0)
    assert [int(x) for x in w[0]] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
