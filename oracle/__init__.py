"""CPU oracle for the jgb200 hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement, in NumPy / plain C, of the algorithms the reference (mcosovic/JuliaGrid.jl v0.6.2)
runs on the path named by BASELINE.json: `newtonRaphson()/mismatch!()/solve!()` and
`gaussNewton()/increment!()/solve!()`. Each function cites the reference file:line it follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this package.  The product package (`juliagrid.jl_b200/`, imported as `jgb200`) never does; it
fails loudly when its CUDA library is missing.

Parity pinning: the oracle is checked against every golden vector the reference's own tests hold for
this path (`test/data/results.h5` -> `tests/golden/*.json`, made by `tests/golden/make_golden.py`):
IEEE-14 NR (7 iterations), IEEE-30 NR (4 iterations), the derived power vectors, the WLS recovery
property (1e-10) and the bad-data known answers (objective 3227.3, chi2 threshold 109.7).

Third-party arithmetic: the reference's sparse factorisations live in SuiteSparse (UMFPACK/KLU/CHOLMOD via
Julia's SparseArrays stdlib and KLU.jl 0.6 — unpinned, no Manifest.toml, sources not under
/root/reference).  The oracle uses SciPy's SuperLU (`scipy.sparse.linalg.splu`) for those solves; the
reference's tests pin results only through converged voltages / iteration counts, which are independent
of the factoriser (the same goldens are asserted for LU, KLU and QR).
"""

from .system import System, load_matpower, load_hdf5, synthetic_grid, system_from_arrays  # noqa: F401
from .model import ac_model, AcModel  # noqa: F401
from . import nr, wls, post  # noqa: F401
