"""Generate the committed golden fixtures from the reference tree (run in the dev container only).

    python tests/golden/make_golden.py [/root/reference]

Reads the reference's own test inputs and MATPOWER-generated golden vectors
(`test/data/case14test.m`, `case30test.m`, `test/data/results.h5`, asserted by
`test/powerFlow/analysis.jl:5-67`) plus the shipped 10k-bus case
(`docs/src/examples/cases/hdf5/case_ACTIVSg10k.h5`) and writes:

  tests/golden/case14test.json, case30test.json   parsed inputs (per-unit, 0-based) + golden NR outputs
  tests/golden/case_ACTIVSg10k.npz                parsed 10k-bus case (inputs only)
  tests/golden/known_answers.json                 WLS known answers of test/stateEstimation/badData.jl:24-41
  tests/golden/monitoring14.json                  the reference's own measurement file src/data/monitoring.h5 (datasets
                                                  in the saveMeasurement layout) + its power system src/data/case14.h5

/root/reference does not exist on the GPU box, so tests only ever read these fixtures.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle.hdf5mini import H5File  # noqa: E402
from oracle.system import load_matpower, load_hdf5  # noqa: E402

SYS_FIELDS = ["bus_type", "pd", "qd", "gs", "bs", "vm", "va", "frm", "to", "r", "x", "g", "b", "tap", "shift",
              "status", "gen_bus", "gen_p", "gen_q", "gen_vm", "gen_status", "gen_qmin", "gen_qmax"]


def system_dict(s):
    d = {"n": s.n, "nbr": s.nbr, "ngen": s.ngen, "slack": s.slack, "base_mva": s.base_mva, "labels": s.labels}
    for k in SYS_FIELDS:
        v = getattr(s, k)
        d[k] = [None if (isinstance(x, float) and not np.isfinite(x)) else x for x in v.tolist()]
    return d


def main(ref="/root/reference"):
    res = H5File(os.path.join(ref, "test/data/results.h5"))
    for case in ("case14test", "case30test"):
        s = load_matpower(os.path.join(ref, f"test/data/{case}.m"))
        out = {"source": f"test/data/{case}.m + test/data/results.h5:/{case}/newtonRaphson",
               "system": system_dict(s), "newtonRaphson": {}}
        grp = f"/{case}/newtonRaphson"
        for k in res.keys(grp):
            v = res[grp + "/" + k]
            out["newtonRaphson"][k] = np.asarray(v).reshape(-1).tolist()
        out["dcPowerFlow"] = {k: np.asarray(res[f"/{case}/dcPowerFlow/{k}"]).reshape(-1).tolist()
                              for k in res.keys(f"/{case}/dcPowerFlow")}
        out["source"] += f" + /{case}/dcPowerFlow"
        rl = f"/{case}/reactiveLimit/newtonRaphson"
        out["reactiveLimit"] = {k: np.asarray(res[rl + "/" + k]).reshape(-1).tolist() for k in res.keys(rl)}
        out["source"] += f" + {rl}"
        for name in ("fastNewtonRaphsonBX", "fastNewtonRaphsonXB"):
            out[name] = {k: np.asarray(res[f"/{case}/{name}/{k}"]).reshape(-1).tolist() for k in res.keys(f"/{case}/{name}")}
        out["source"] += f" + /{case}/fastNewtonRaphsonBX|XB"
        with open(os.path.join(HERE, f"{case}.json"), "w") as fh:
            json.dump(out, fh)
        print(case, "iteration", out["newtonRaphson"]["iteration"])

    s = load_hdf5(os.path.join(ref, "docs/src/examples/cases/hdf5/case_ACTIVSg10k.h5"))
    arrays = {k: getattr(s, k) for k in SYS_FIELDS}
    # float32-exact fields stay float64: the solver's parity is judged on these exact inputs
    np.savez_compressed(os.path.join(HERE, "case_ACTIVSg10k.npz"), n=s.n, nbr=s.nbr, ngen=s.ngen, slack=s.slack,
                        base_mva=s.base_mva, **arrays)
    print("ACTIVSg10k", s.n, s.nbr, s.ngen, os.path.getsize(os.path.join(HERE, "case_ACTIVSg10k.npz")))

    # the largest single-interconnect case the reference ships (70 000 buses): scale test of the single-case path
    s = load_hdf5(os.path.join(ref, "docs/src/examples/cases/hdf5/case_ACTIVSg70k.h5"))
    np.savez_compressed(os.path.join(HERE, "case_ACTIVSg70k.npz"), n=s.n, nbr=s.nbr, ngen=s.ngen, slack=s.slack,
                        base_mva=s.base_mva, **{k: getattr(s, k) for k in SYS_FIELDS})
    print("ACTIVSg70k", s.n, s.nbr, s.ngen, os.path.getsize(os.path.join(HERE, "case_ACTIVSg70k.npz")))

    # a measurement file written by the reference's saveMeasurement (src/data/monitoring.h5, the companion of
    # src/data/case14.h5): the dataset dictionary in the file's own layout + the power system it belongs to
    mon = H5File(os.path.join(ref, "src/data/monitoring.h5"))
    datasets = {}
    for dev in mon.keys("/"):
        for sub in mon.keys("/" + dev):
            if not mon.is_group(f"/{dev}/{sub}"):
                continue                                   # labels are not kept (meters are addressed by position)
            for k in mon.keys(f"/{dev}/{sub}"):
                if k == "label":
                    continue
                a = np.asarray(mon[f"/{dev}/{sub}/{k}"])
                datasets[f"{dev}/{sub}/{k}"] = a.reshape(-1).tolist() if a.ndim else a.item()
    s14 = load_hdf5(os.path.join(ref, "src/data/case14.h5"))
    out = {"source": "src/data/monitoring.h5 (saveMeasurement layout, measurement/save.jl:40-118) + src/data/case14.h5",
           "system": system_dict(s14), "attrs": {k: int(v) for k, v in mon.attrs("/").items()}, "datasets": datasets}
    with open(os.path.join(HERE, "monitoring14.json"), "w") as fh:
        json.dump(out, fh)
    print("monitoring14", out["attrs"])

    known = {
        "badData_one_outlier": {
            "source": "test/stateEstimation/badData.jl:5-41",
            "setup": "case14test.m; bus 1 -> PV, bus 3 -> slack with angle -0.17; NR truth; voltmeters + wattmeters + "
                     "varmeters everywhere, all variances 1e-2; 'Varmeter 4' (reactive injection at 4th bus) = 10.25",
            "objective": 3227.3, "threshold": 109.7, "atol": 0.1,
        },
        "wls_recovery_atol": 1e-10,
    }
    with open(os.path.join(HERE, "known_answers.json"), "w") as fh:
        json.dump(known, fh, indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:])
