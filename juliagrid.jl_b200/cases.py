"""Host-side PowerSystem container and case loaders of the product (data plumbing, no numerics).

Mirrors the fields of the reference's `PowerSystem` that the hot path reads (`src/definition/system.jl:51-233`),
already in per-unit / radians like the reference keeps them internally. Index arrays are 0-based here and are
converted to the reference's 1-based Int64 layout at the C-ABI boundary.
"""
from __future__ import annotations

import copy
import json
import re
from dataclasses import dataclass, field

import numpy as np


@dataclass
class PowerSystem:
    n: int
    bus_type: np.ndarray      # bus.layout.type (Int8): 1 PQ, 2 PV, 3 slack
    slack: int                # bus.layout.slack (0-based here)
    pd: np.ndarray            # bus.demand.active
    qd: np.ndarray            # bus.demand.reactive
    gs: np.ndarray            # bus.shunt.conductance
    bs: np.ndarray            # bus.shunt.susceptance
    vm: np.ndarray            # bus.voltage.magnitude
    va: np.ndarray            # bus.voltage.angle
    nbr: int
    frm: np.ndarray           # branch.layout.from
    to: np.ndarray            # branch.layout.to
    r: np.ndarray
    x: np.ndarray
    g: np.ndarray             # branch.parameter.conductance
    b: np.ndarray             # branch.parameter.susceptance
    tap: np.ndarray           # branch.parameter.turnsRatio
    shift: np.ndarray         # branch.parameter.shiftAngle
    status: np.ndarray        # branch.layout.status
    ngen: int
    gen_bus: np.ndarray
    gen_p: np.ndarray
    gen_q: np.ndarray
    gen_vm: np.ndarray
    gen_status: np.ndarray
    base_mva: float = 100.0
    labels: list = field(default_factory=list)
    model: object = None      # system.model.ac once ac_model() ran
    gen_qmin: np.ndarray = None   # generator.capability.minReactive (default -Inf)
    gen_qmax: np.ndarray = None   # generator.capability.maxReactive (default +Inf)
    supply_p: np.ndarray = None   # bus.supply.active / reactive once reactive_limit() rewrote them; None = derived from
    supply_q: np.ndarray = None   # the in-service generators

    def __post_init__(self):
        if self.gen_qmin is None:
            self.gen_qmin = np.full(self.ngen, -np.inf)
        if self.gen_qmax is None:
            self.gen_qmax = np.full(self.ngen, np.inf)

    def copy(self) -> "PowerSystem":
        return copy.deepcopy(self)

    @property
    def supply(self):
        """bus.supply.active / reactive and bus.supply.generator (first in-service generator per bus, -1 if none)."""
        on = self.gen_status == 1
        sp = np.zeros(self.n)
        sq = np.zeros(self.n)
        np.add.at(sp, self.gen_bus[on], self.gen_p[on])
        np.add.at(sq, self.gen_bus[on], self.gen_q[on])
        first = np.full(self.n, -1, dtype=np.int64)
        idx = np.flatnonzero(on)
        first[self.gen_bus[idx][::-1]] = idx[::-1]
        if self.supply_p is not None:
            return self.supply_p.copy(), self.supply_q.copy(), first
        return sp, sq, first


_FIELDS = ["bus_type", "pd", "qd", "gs", "bs", "vm", "va", "frm", "to", "r", "x", "g", "b", "tap", "shift", "status",
           "gen_bus", "gen_p", "gen_q", "gen_vm", "gen_status"]
_I8 = {"bus_type", "status", "gen_status"}
_I64 = {"frm", "to", "gen_bus"}


def _from_mapping(d) -> PowerSystem:
    kw = {}
    for k in _FIELDS:
        dt = np.int8 if k in _I8 else np.int64 if k in _I64 else np.float64
        kw[k] = np.asarray(d[k], dtype=dt)
    for k in ("gen_qmin", "gen_qmax"):
        if k in d:          # JSON fixtures store +-Inf as null
            sign = -1.0 if k == "gen_qmin" else 1.0
            kw[k] = np.array([sign * np.inf if v is None else float(v) for v in np.asarray(d[k], dtype=object)],
                             dtype=np.float64)
    return PowerSystem(n=int(d["n"]), slack=int(d["slack"]), nbr=int(d["nbr"]), ngen=int(d["ngen"]),
                       base_mva=float(d["base_mva"]),
                       labels=list(d["labels"]) if "labels" in d else list(range(1, int(d["n"]) + 1)), **kw)


def _matrix(text: str, name: str):
    m = re.search(r"mpc\." + name + r"\s*=\s*\[(.*?)\]", text, re.S)
    if not m:
        raise ValueError(f"mpc.{name} is missing")
    rows = [ln.split("%")[0].replace(";", " ").split() for ln in m.group(1).splitlines()]
    return [[float(t) for t in r] for r in rows if r]


def power_system(path: str) -> PowerSystem:
    """powerSystem(file): MATPOWER `.m`, or the repo's `.json` / `.npz` fixtures."""
    if path.endswith(".json"):
        with open(path) as fh:
            d = json.load(fh)
        return _from_mapping(d["system"] if "system" in d else d)
    if path.endswith(".npz"):
        z = np.load(path)
        return _from_mapping({k: z[k] for k in z.files})
    if path.endswith(".h5"):
        return _from_hdf5(path)
    text = open(path).read()
    base = float(re.search(r"mpc\.baseMVA\s*=\s*([^;]+);", text).group(1))
    inv = 1.0 / base
    bus = np.array(_matrix(text, "bus"))
    gen = np.array([r[:8] for r in _matrix(text, r"gen(?!cost)")])
    br = np.array([r[:11] for r in _matrix(text, "branch")])
    labels = bus[:, 0].astype(int).tolist()
    pos = {lab: k for k, lab in enumerate(labels)}
    look = np.vectorize(lambda v: pos[int(v)])
    tap = br[:, 8].copy()
    tap[tap == 0.0] = 1.0
    bus_type = bus[:, 1].astype(np.int8)
    sl = np.flatnonzero(bus_type == 3)
    return PowerSystem(
        n=len(bus), bus_type=bus_type, slack=int(sl[-1]) if len(sl) else 0,
        pd=bus[:, 2] * inv, qd=bus[:, 3] * inv, gs=bus[:, 4] * inv, bs=bus[:, 5] * inv,
        vm=bus[:, 7].copy(), va=bus[:, 8] * (np.pi / 180),
        nbr=len(br), frm=look(br[:, 0]).astype(np.int64), to=look(br[:, 1]).astype(np.int64),
        r=br[:, 2].copy(), x=br[:, 3].copy(), g=np.zeros(len(br)), b=br[:, 4].copy(), tap=tap,
        shift=br[:, 9] * (np.pi / 180), status=br[:, 10].astype(np.int8),
        ngen=len(gen), gen_bus=look(gen[:, 0]).astype(np.int64), gen_p=gen[:, 1] * inv, gen_q=gen[:, 2] * inv,
        gen_vm=gen[:, 5].copy(), gen_status=gen[:, 7].astype(np.int8), base_mva=base, labels=labels,
        gen_qmin=gen[:, 4] * inv, gen_qmax=gen[:, 3] * inv)


def _from_hdf5(path: str) -> PowerSystem:
    """JuliaGrid's HDF5 case layout (load.jl:141-289): positional 1-based indices, per-unit / radian values, a scalar
    dataset stands for a constant vector; the base power attribute is in VA."""
    from .h5 import H5File
    f = H5File(path)
    at = f.attrs("/")
    n, nbr, ngen = (int(at[k]) for k in ("number of buses", "number of branches", "number of generators"))

    def vec(p, count, dtype=np.float64):
        v = np.asarray(f[p])
        if v.ndim == 0 or (v.size == 1 and count != 1):
            return np.full(count, v.reshape(-1)[0], dtype=dtype)
        return v.astype(dtype)

    bus_type = vec("/bus/layout/type", n, np.int8)
    sl = np.flatnonzero(bus_type == 3)
    return PowerSystem(
        n=n, bus_type=bus_type, slack=int(sl[-1]) if len(sl) else 0,
        pd=vec("/bus/demand/active", n), qd=vec("/bus/demand/reactive", n),
        gs=vec("/bus/shunt/conductance", n), bs=vec("/bus/shunt/susceptance", n),
        vm=vec("/bus/voltage/magnitude", n), va=vec("/bus/voltage/angle", n),
        nbr=nbr, frm=vec("/branch/layout/from", nbr, np.int64) - 1, to=vec("/branch/layout/to", nbr, np.int64) - 1,
        r=vec("/branch/parameter/resistance", nbr), x=vec("/branch/parameter/reactance", nbr),
        g=vec("/branch/parameter/conductance", nbr), b=vec("/branch/parameter/susceptance", nbr),
        tap=vec("/branch/parameter/turnsRatio", nbr), shift=vec("/branch/parameter/shiftAngle", nbr),
        status=vec("/branch/layout/status", nbr, np.int8),
        ngen=ngen, gen_bus=vec("/generator/layout/bus", ngen, np.int64) - 1,
        gen_p=vec("/generator/output/active", ngen), gen_q=vec("/generator/output/reactive", ngen),
        gen_vm=vec("/generator/voltage/magnitude", ngen), gen_status=vec("/generator/layout/status", ngen, np.int8),
        gen_qmin=vec("/generator/capability/minReactive", ngen), gen_qmax=vec("/generator/capability/maxReactive", ngen),
        base_mva=float(np.asarray(f["/base/power"]).reshape(-1)[0]) / 1e6, labels=list(range(1, n + 1)))


def synthetic_grid(side: int = 100, seed: int = 20261017) -> PowerSystem:
    """Deterministic synthetic meshed grid (benchmark input). side=100 -> 10 000 buses,
    12 699 branches, 1 500 PV buses, dim J = 18 498. Draw order is part of the recipe."""
    rng = np.random.default_rng(seed)
    n = side * side
    idx = np.arange(n).reshape(side, side)
    cand_f, cand_t = idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    pick = np.sort(rng.choice(len(cand_f), int(0.27 * n), replace=False))
    frm = np.concatenate([idx[:-1, :].ravel(), idx[0, :-1].ravel(), cand_f[pick]]).astype(np.int64)
    to = np.concatenate([idx[1:, :].ravel(), idx[0, 1:].ravel(), cand_t[pick]]).astype(np.int64)
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
    return PowerSystem(n=n, bus_type=bus_type, slack=0, pd=pd, qd=qd, gs=np.zeros(n), bs=np.zeros(n), vm=np.ones(n),
                       va=np.zeros(n), nbr=m, frm=frm, to=to, r=r, x=x, g=np.zeros(m), b=b, tap=tap,
                       shift=np.zeros(m), status=np.ones(m, dtype=np.int8), ngen=ngen, gen_bus=gen_bus, gen_p=gen_p,
                       gen_q=np.zeros(ngen), gen_vm=np.full(ngen, 1.02), gen_status=np.ones(ngen, dtype=np.int8))
