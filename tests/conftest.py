import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import get_oracle
    return get_oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference CPU codec, where oracle/_ref/libndzip_ref.so exists or can be built."""
    from oracle import get_reference
    ref = get_reference()
    if ref is None:
        pytest.skip("oracle/_ref/libndzip_ref.so not present and /root/reference not available")
    return ref


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)["rows"]
