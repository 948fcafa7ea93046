"""The tcgen05 / DMMA Schur-update A/B of DESIGN.md section 7 (juliagrid.jl_b200/csrc/experiments/schur_tcgen05.cu) runs and
both variants reproduce a long-double host reference: the measured refutation stays reproducible."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "juliagrid.jl_b200", "csrc", "experiments", "schur_tcgen05")


@pytest.mark.gpu
def test_schur_update_on_tcgen05_and_dmma_matches_reference():
    if not os.path.exists(EXE):
        pytest.skip("experiment binary not built (make -C juliagrid.jl_b200/csrc experiments)")
    out = subprocess.run([EXE, "148", "2"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("K=")]
    assert len(lines) == 3 and out.stdout.strip().endswith("OK")
    for ln in lines:      # "... err 8.51e-16) ... err 2.97e-14) ..."
        errs = [float(tok.rstrip(")")) for prev, tok in zip(ln.split(), ln.split()[1:]) if prev == "err"]
        assert len(errs) == 2 and errs[0] < 1e-14 and errs[1] < 1e-13
