"""The tuning switches of the solver select other code paths (gather in rounds instead of the cp.async.bulk ring, one
stream instead of the per-class DAG schedule, scenario-tile LU instead of the dense LU kernel for a single case, chains of
fronts walked by one CTA). They are read once per process, so every setting re-runs the Newton-Raphson and batch parity
tests in a child process: the A/B numbers in profiles/ must come from paths that are all correct."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [
    {"JGB_STAGED_EA": "0", "JGB_BULK_RING": "0", "JGB_LANES": "1"},
    {"JGB_DENSE_LU_MIN": "0"},
    {"JGB_SEQ": "4"},
]


@pytest.mark.gpu
@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_parity_under_tuning_switches(env):
    if os.environ.get("JGB_VARIANT_CHILD"):
        pytest.skip("child run")
    child = dict(os.environ, JGB_VARIANT_CHILD="1", **env)
    out = subprocess.run([sys.executable, "-m", "pytest", "tests/test_nr_gpu.py", "tests/test_batch_gpu.py", "-x", "-q", "-m", "gpu",
                          "-p", "no:cacheprovider"], cwd=ROOT, env=child, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
