import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_TESTDATA = "/root/reference/pkg/suggest/testdata"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cars_lines():
    with open(os.path.join(GOLDEN, "cars.dict"), "rb") as f:
        return f.read().split(b"\n")[:-1]  # bufio.Scanner lines, pkg/dictionary/helpers.go:38-45


CARS_DESCRIPTION = dict(ngram_size=3, wrap=("$", "$"), pad="$", alphabet=("russian", "english", "numbers", "$"))
TEST_DESCRIPTION = dict(ngram_size=3, wrap=("$", "$"), pad="$", alphabet=("english", "russian", "numbers", "$"))

COLLECTION = ["Nissan March", "Nissan Juke", "Nissan Maxima", "Nissan Murano", "Nissan Note", "Toyota Mark II",
              "Toyota Corolla", "Toyota Corona"]
