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


def golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as fh:
        return json.load(fh)


def oracle_system(name):
    import oracle
    if name == "synthetic10k":
        return oracle.synthetic_grid()
    if name.startswith("synthetic"):
        return oracle.synthetic_grid(side=int(name[len("synthetic"):]))
    if name.startswith("case_ACTIVSg"):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        return oracle.system_from_arrays({k: z[k] for k in z.files})
    return oracle.system_from_arrays(golden(name)["system"])


def product_system(name):
    import jgb200
    if name == "synthetic10k":
        return jgb200.synthetic_grid()
    if name.startswith("synthetic"):
        return jgb200.synthetic_grid(side=int(name[len("synthetic"):]))
    if name.startswith("case_ACTIVSg"):
        return jgb200.power_system(os.path.join(GOLDEN, name + ".npz"))
    return jgb200.power_system(os.path.join(GOLDEN, name + ".json"))


@pytest.fixture(scope="session")
def ctx():
    import jgb200
    c = jgb200.Context(0)
    yield c
    c.close()
