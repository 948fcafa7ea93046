"""Oracle-side PowerSystem container and loaders (TEST INFRASTRUCTURE).

Follows the reference's loaders: MATPOWER `.m` -> `src/powerSystem/load.jl:292-619`
(MW -> p.u. by 1/baseMVA, degrees -> radians, `ratio == 0 -> 1.0`, bus labels mapped to 1..n in file
order, per-bus supply = sum of in-service generators), HDF5 -> `src/powerSystem/load.jl:141-289`
with `readHDF5` `:1360-1368` (scalar dataset = constant vector).

All index arrays are 0-based here; `oracle.nr.export_one_based` converts to the reference's
1-based Int64 layout for index-set parity checks.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import copy
import re
import numpy as np


@dataclass
class System:
    # buses
    n: int
    bus_type: np.ndarray        # int8: 1 PQ, 2 PV, 3 slack   (bus.layout.type)
    slack: int                  # 0-based                      (bus.layout.slack)
    pd: np.ndarray              # bus.demand.active  [p.u.]
    qd: np.ndarray
    gs: np.ndarray              # bus.shunt.conductance
    bs: np.ndarray
    vm: np.ndarray              # bus.voltage.magnitude
    va: np.ndarray              # bus.voltage.angle [rad]
    supply_p: np.ndarray        # bus.supply.active  (sum over in-service generators)
    supply_q: np.ndarray
    bus_gens: list              # bus.supply.generator: list of generator indices per bus (in-service only)
    # branches
    nbr: int
    frm: np.ndarray             # 0-based (branch.layout.from)
    to: np.ndarray
    r: np.ndarray
    x: np.ndarray
    g: np.ndarray               # branch.parameter.conductance (total line charging conductance)
    b: np.ndarray               # branch.parameter.susceptance
    tap: np.ndarray             # turnsRatio
    shift: np.ndarray           # shiftAngle [rad]
    status: np.ndarray          # int8
    # generators
    ngen: int
    gen_bus: np.ndarray
    gen_p: np.ndarray
    gen_q: np.ndarray
    gen_vm: np.ndarray
    gen_status: np.ndarray
    gen_qmin: np.ndarray = None
    gen_qmax: np.ndarray = None
    base_mva: float = 100.0
    labels: list = field(default_factory=list)

    def copy(self) -> "System":
        return copy.deepcopy(self)

    def rebuild_supply(self):
        """bus.supply.* and bus.supply.generator from in-service generators (load.jl:603-610)."""
        self.supply_p = np.zeros(self.n)
        self.supply_q = np.zeros(self.n)
        self.bus_gens = [[] for _ in range(self.n)]
        for k in range(self.ngen):
            if self.gen_status[k] == 1:
                i = int(self.gen_bus[k])
                self.bus_gens[i].append(k)
                self.supply_p[i] += self.gen_p[k]
                self.supply_q[i] += self.gen_q[k]


def _matrix_block(text: str, name: str):
    m = re.search(r"mpc\." + name + r"\s*=\s*\[(.*?)\]", text, re.S)
    if m is None:
        return []
    rows = []
    for line in m.group(1).splitlines():
        line = line.split("%")[0].replace(";", " ").strip()
        if line:
            rows.append([float(tok) for tok in line.split()])
    return rows


def load_matpower(path: str) -> System:
    """MATPOWER case reader — restates matpowerRead/Bus/Branch/Generator (load.jl:292-619)."""
    text = open(path).read()
    mva = re.search(r"mpc\.baseMVA\s*=\s*([^;]+);", text)
    base = float(mva.group(1)) if mva else 100.0
    inv = 1.0 / base
    deg2rad = np.pi / 180

    bus = _matrix_block(text, "bus")
    # `mpc.gen` must not match `mpc.gencost` (load.jl:308)
    gen = _matrix_block(text, r"gen(?!cost)")
    br = _matrix_block(text, "branch")
    if not bus or not br or not gen:
        raise ValueError("bus / branch / generator data missing")

    n = len(bus)
    labels = [int(row[0]) for row in bus]
    index = {lab: k for k, lab in enumerate(labels)}
    bus_type = np.array([int(row[1]) for row in bus], dtype=np.int8)
    slack_pos = np.flatnonzero(bus_type == 3)
    slack = int(slack_pos[-1]) if len(slack_pos) else 0  # last type-3 bus wins (load.jl:414-416)

    nbr = len(br)
    tap = np.array([row[8] for row in br])
    tap[tap == 0.0] = 1.0

    ngen = len(gen)
    sys = System(
        n=n, bus_type=bus_type, slack=slack,
        pd=np.array([row[2] for row in bus]) * inv, qd=np.array([row[3] for row in bus]) * inv,
        gs=np.array([row[4] for row in bus]) * inv, bs=np.array([row[5] for row in bus]) * inv,
        vm=np.array([row[7] for row in bus]), va=np.array([row[8] for row in bus]) * deg2rad,
        supply_p=np.zeros(n), supply_q=np.zeros(n), bus_gens=[[] for _ in range(n)],
        nbr=nbr,
        frm=np.array([index[int(row[0])] for row in br], dtype=np.int64),
        to=np.array([index[int(row[1])] for row in br], dtype=np.int64),
        r=np.array([row[2] for row in br]), x=np.array([row[3] for row in br]),
        g=np.zeros(nbr), b=np.array([row[4] for row in br]),
        tap=tap, shift=np.array([row[9] for row in br]) * deg2rad,
        status=np.array([int(row[10]) for row in br], dtype=np.int8),
        ngen=ngen,
        gen_bus=np.array([index[int(row[0])] for row in gen], dtype=np.int64),
        gen_p=np.array([row[1] for row in gen]) * inv, gen_q=np.array([row[2] for row in gen]) * inv,
        gen_vm=np.array([row[5] for row in gen]),
        gen_status=np.array([int(row[7]) for row in gen], dtype=np.int8),
        gen_qmax=np.array([row[3] for row in gen]) * inv, gen_qmin=np.array([row[4] for row in gen]) * inv,
        base_mva=base, labels=labels,
    )
    sys.rebuild_supply()
    return sys


def load_hdf5(path: str) -> System:
    """Reference HDF5 case reader (load.jl:141-289); positional 1-based indices -> 0-based."""
    from .hdf5mini import H5File

    f = H5File(path)
    at = f.attrs("/")
    n = int(at["number of buses"])
    nbr = int(at["number of branches"])
    ngen = int(at["number of generators"])

    def vec(p, count, dtype=float):
        v = f[p]
        v = np.asarray(v)
        if v.ndim == 0 or v.size == 1 and count != 1:
            return np.full(count, v.reshape(-1)[0], dtype=dtype)
        return v.astype(dtype)

    bus_type = vec("/bus/layout/type", n, np.int8)
    slack = int(np.flatnonzero(bus_type == 3)[-1])
    sys = System(
        n=n, bus_type=bus_type, slack=slack,
        pd=vec("/bus/demand/active", n), qd=vec("/bus/demand/reactive", n),
        gs=vec("/bus/shunt/conductance", n), bs=vec("/bus/shunt/susceptance", n),
        vm=vec("/bus/voltage/magnitude", n), va=vec("/bus/voltage/angle", n),
        supply_p=np.zeros(n), supply_q=np.zeros(n), bus_gens=[[] for _ in range(n)],
        nbr=nbr,
        frm=vec("/branch/layout/from", nbr, np.int64) - 1, to=vec("/branch/layout/to", nbr, np.int64) - 1,
        r=vec("/branch/parameter/resistance", nbr), x=vec("/branch/parameter/reactance", nbr),
        g=vec("/branch/parameter/conductance", nbr), b=vec("/branch/parameter/susceptance", nbr),
        tap=vec("/branch/parameter/turnsRatio", nbr), shift=vec("/branch/parameter/shiftAngle", nbr),
        status=vec("/branch/layout/status", nbr, np.int8),
        ngen=ngen,
        gen_bus=vec("/generator/layout/bus", ngen, np.int64) - 1,
        gen_p=vec("/generator/output/active", ngen), gen_q=vec("/generator/output/reactive", ngen),
        gen_vm=vec("/generator/voltage/magnitude", ngen),
        gen_status=vec("/generator/layout/status", ngen, np.int8),
        gen_qmin=vec("/generator/capability/minReactive", ngen),
        gen_qmax=vec("/generator/capability/maxReactive", ngen),
        base_mva=float(np.asarray(f["/base/power"]).reshape(-1)[0]) / 1e6,
        labels=list(range(1, n + 1)),
    )
    sys.rebuild_supply()
    return sys


def synthetic_grid(side: int = 100, seed: int = 20261017) -> System:
    """Deterministic synthetic meshed grid of SURVEY.md Appendix D (NOT from the reference).

    side=100 gives the 10 000-bus benchmark grid: 12 699 branches, nnz(Y)=35 398, dim J=18 498.
    The product has an independent copy of this recipe (`jgb200.cases.synthetic_grid`); tests assert
    both produce identical arrays.
    """
    rng = np.random.default_rng(seed)
    n = side * side
    idx = np.arange(n).reshape(side, side)
    f1, t1 = idx[:-1, :].ravel(), idx[1:, :].ravel()
    f2, t2 = idx[0, :-1].ravel(), idx[0, 1:].ravel()
    f3, t3 = idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    k = int(0.27 * n)
    pick = rng.choice(len(f3), k, replace=False)
    pick.sort()
    frm = np.concatenate([f1, f2, f3[pick]]).astype(np.int64)
    to = np.concatenate([t1, t2, t3[pick]]).astype(np.int64)
    m = len(frm)
    x = rng.uniform(0.01, 0.08, m)
    r = x * rng.uniform(0.1, 0.3, m)
    b = rng.uniform(0.0, 0.04, m)
    tap = np.ones(m)
    istr = rng.choice(m, m // 20, replace=False)
    tap[istr] = rng.uniform(0.95, 1.05, len(istr))
    bus_type = np.ones(n, dtype=np.int8)
    pv = rng.choice(np.arange(1, n), int(0.15 * n), replace=False)
    bus_type[pv] = 2
    bus_type[0] = 3
    pd = rng.uniform(0.0, 0.02, n)
    qd = pd * rng.uniform(0.1, 0.4, n)
    gen_bus = np.concatenate([[0], np.sort(pv)]).astype(np.int64)
    ngen = len(gen_bus)
    gen_p = np.full(ngen, pd.sum() / ngen)
    gen_p[0] = 0.0
    sys = System(
        n=n, bus_type=bus_type, slack=0, pd=pd, qd=qd, gs=np.zeros(n), bs=np.zeros(n),
        vm=np.ones(n), va=np.zeros(n), supply_p=np.zeros(n), supply_q=np.zeros(n),
        bus_gens=[[] for _ in range(n)], nbr=m, frm=frm, to=to, r=r, x=x, g=np.zeros(m), b=b, tap=tap,
        shift=np.zeros(m), status=np.ones(m, dtype=np.int8), ngen=ngen, gen_bus=gen_bus, gen_p=gen_p,
        gen_q=np.zeros(ngen), gen_vm=np.full(ngen, 1.02), gen_status=np.ones(ngen, dtype=np.int8),
        gen_qmin=np.full(ngen, -np.inf), gen_qmax=np.full(ngen, np.inf), base_mva=100.0,
        labels=list(range(1, n + 1)),
    )
    sys.rebuild_supply()
    return sys


_ARRAY_FIELDS = ["bus_type", "pd", "qd", "gs", "bs", "vm", "va", "frm", "to", "r", "x", "g", "b", "tap", "shift",
                 "status", "gen_bus", "gen_p", "gen_q", "gen_vm", "gen_status", "gen_qmin", "gen_qmax"]
_INT8 = {"bus_type", "status", "gen_status"}
_INT64 = {"frm", "to", "gen_bus"}


def system_from_arrays(d) -> System:
    """Rebuild a System from the committed fixtures (tests/golden/*.json 'system' dicts or .npz files)."""
    def arr(k):
        v = d[k]
        if k in _INT8:
            return np.asarray(v, dtype=np.int8)
        if k in _INT64:
            return np.asarray(v, dtype=np.int64)
        if isinstance(v, list):
            v = [np.inf if x is None else x for x in v]
        return np.asarray(v, dtype=float)

    n, nbr, ngen = int(d["n"]), int(d["nbr"]), int(d["ngen"])
    kw = {k: arr(k) for k in _ARRAY_FIELDS}
    if isinstance(d.get("gen_qmin"), list):
        kw["gen_qmin"] = np.asarray([-np.inf if x is None else x for x in d["gen_qmin"]], dtype=float)
    sys = System(n=n, slack=int(d["slack"]), nbr=nbr, ngen=ngen, supply_p=np.zeros(n), supply_q=np.zeros(n),
                 bus_gens=[[] for _ in range(n)], base_mva=float(d["base_mva"]),
                 labels=list(d["labels"]) if "labels" in d else list(range(1, n + 1)), **kw)
    sys.rebuild_supply()
    return sys
