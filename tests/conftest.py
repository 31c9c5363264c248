import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: CPU test that takes more than a few seconds")


def golden(name):
    with open(os.path.join(GOLDEN, name + ".json"), encoding="utf-8") as f:
        return json.load(f)


def rows_to_grid(rows):
    w = max(len(r) for r in rows)
    g = np.zeros((len(rows), w), np.uint8)
    for y, r in enumerate(rows):
        for x, c in enumerate(r):
            g[y, x] = c == "X"
    return g


def splitmix64(z):
    """the standard 3-step finaliser on z += 0x9E3779B97F4A7C15 (vectorised, uint64)."""
    z = (np.asarray(z, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synth_terrain(w, h, seed=1, t=0, density_q24=11744051):
    """SURVEY.md §8(d) C4/C5 generator: ceiling iff (splitmix64(seed*GOLDEN + (t<<20) + y*w + x) >> 40) < floor(0.7*2^24)."""
    with np.errstate(over="ignore"):
        idx = np.arange(w * h, dtype=np.uint64) + (np.uint64(t) << np.uint64(20))
        base = np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        r = splitmix64(base + idx) >> np.uint64(40)
    return (r < np.uint64(density_q24)).astype(np.uint8).reshape(h, w)


@pytest.fixture(scope="session")
def fixtures():
    return {k: rows_to_grid(v["grid"]) for k, v in golden("fixtures").items()}


@pytest.fixture(scope="session")
def readme():
    d = golden("readme_layouts")
    return rows_to_grid(d["terrain"]), d["layouts"]
