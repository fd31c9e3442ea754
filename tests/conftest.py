import gzip
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        from magot_b200 import _lib
        return _lib.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ref_data(tmp_path_factory):
    """The reference's data fixtures (gunzipped copies of /root/reference/test_data) + surrogate C14."""
    d = tmp_path_factory.mktemp("ref_test_data")
    src = os.path.join(GOLDEN, "ref_test_data")
    for f in os.listdir(src):
        if f.endswith(".gz"):
            with gzip.open(os.path.join(src, f), "rb") as fh, open(os.path.join(str(d), f[:-3]), "wb") as out:
                out.write(fh.read())
    import surrogate_c14
    with open(os.path.join(str(d), "C14.fasta"), "w", encoding="latin-1", newline="\n") as fh:
        fh.write(surrogate_c14.build(os.path.join(str(d), "StandardGTF.gtf"), os.path.join(str(d), "CDSannotations.cds")))
    return str(d)


@pytest.fixture(scope="session")
def manifest():
    import json
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def kat():
    import json
    with open(os.path.join(GOLDEN, "kat.json")) as fh:
        return json.load(fh)
