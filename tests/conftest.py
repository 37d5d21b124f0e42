import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the shared library is git-ignored: build it in-tree if this checkout does not have it yet (nvcc cross-compiles
    # without a GPU).  There is no fallback implementation to fall back to.
    lib = os.path.join(ROOT, "clownresampler_b200", "lib", "libclownresampler_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle.cro import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference (oracle/_ref); skipped where it was not prebuilt and cannot be built."""
    from oracle.cro import Reference
    try:
        return Reference()
    except (FileNotFoundError, OSError) as e:  # pragma: no cover
        pytest.skip(f"oracle/_ref not available: {e}")


@pytest.fixture(scope="session")
def flac_pcm():
    """tests/test.flac decoded to s16 (192000 x 2), committed as a fixture by oracle/make_golden.py."""
    raw = gzip.open(os.path.join(GOLD, "test_flac_s16le.bin.gz"), "rb").read()
    return np.frombuffer(raw, dtype="<i2").reshape(-1, 2).copy()


@pytest.fixture(scope="session")
def tripwires():
    return json.load(open(os.path.join(GOLD, "tripwires.json")))


@pytest.fixture(scope="session")
def ref_vectors():
    meta = json.load(open(os.path.join(GOLD, "ref_vectors.json")))
    data = np.load(os.path.join(GOLD, "ref_vectors.npz"))
    return meta, data


def pad(data, radius):
    """Zero padding of `radius` frames each side, as tests/test-low-level.c:145-152 does."""
    data = np.asarray(data, dtype=np.int16)
    out = np.zeros((data.shape[0] + 2 * radius, data.shape[1]), dtype=np.int16)
    out[radius:radius + data.shape[0]] = data
    return out


CTEST_CASES = [(8000, 44100, 44100), (8000, 44100, 8000), (44100, 8000, 44100), (44100, 8000, 8000)]  # tests/CMakeLists.txt:25-47
