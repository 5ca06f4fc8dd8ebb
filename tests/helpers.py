"""Shared test helpers: build the g3py_b200 objects from the oracle's neutral model spec, so that the
identical description drives both sides (the product never imports the oracle)."""
import numpy as np

import g3py_b200 as g3

LEAVES = {"SE": g3.SE, "OU": g3.OU, "MAT32": g3.MAT32, "MAT52": g3.MAT52, "RQ": g3.RQ, "SIN": g3.SIN, "COS": g3.COS, "SINC": g3.SINC, "SM": g3.SM, "WN": g3.WN,
          "Noise": g3.KernelNoise, "KernelDot": g3.KernelDot, "LIN": g3.LIN, "POL": g3.POL, "BW": g3.BW, "VAR": g3.VAR,
          "NN": g3.NN, "NIL": g3.NIL, "KernelEquals": g3.KernelEquals, "KernelEquals2": g3.KernelEquals2}
MAPS = {"Identity": g3.Identity, "LinearMapping": g3.LinearMapping, "LogShifted": g3.LogShifted,
        "BoxCoxShifted": g3.BoxCoxShifted, "BoxCoxLinear": g3.BoxCoxLinear, "ArcsinhLinear": g3.ArcsinhLinear,
        "SinhArcsinh": g3.SinhArcsinh, "Logistic": g3.Logistic, "WarpingTanh": g3.WarpingTanh,
        "WarpingBoxCox": g3.WarpingBoxCox, "BoxCoxLinear2": g3.BoxCoxLinear2}
MEANS = {"Zero": g3.Zero, "Bias": g3.Bias, "Linear": g3.Linear}


def _x_arg(X, dims):
    if dims is None:
        return X
    return (X, slice(int(dims[0]), int(dims[1])))


def _pot(obj, spec):
    if "potential" in spec:
        obj.set_potential(*spec["potential"])
    return obj


def build_kernel(spec, X):
    return _pot(_build_kernel(spec, X), spec)


def _build_kernel(spec, X):
    t = spec["type"]
    if t == "sum":
        return build_kernel(spec["k1"], X) + build_kernel(spec["k2"], X)
    if t == "prod":
        return build_kernel(spec["k1"], X) * build_kernel(spec["k2"], X)
    if t == "max":
        return g3.KernelMax(build_kernel(spec["k1"], X), build_kernel(spec["k2"], X))
    if t == "scale":
        return spec["c"] * build_kernel(spec["k"], X)
    if t == "shift":
        return spec["c"] + build_kernel(spec["k"], X)
    kw = {}
    if "name" in spec:
        kw["name"] = spec["name"]
    if spec.get("var") is not None:
        kw["var"] = spec["var"]
    if t == "POL" and "p" in spec:
        kw["p"] = spec["p"]
    for key in ("eq", "eq1", "eq2"):
        if key in spec:
            kw[key] = spec[key]
    return LEAVES[t](_x_arg(X, spec.get("dims")), **kw)


def build_transport(spec, X):
    parts = []
    for t in spec["chain"]:
        if t["t"] == "ID":
            parts.append(g3.ID())
        elif t["t"] == "TMapping":
            mp = t["mapping"]
            parts.append(g3.TMapping(MAPS[mp["type"]](**({"name": mp["name"]} if "name" in mp else {}))))
        elif t["t"] == "TScale":
            sc = t["scale"]
            parts.append(g3.TScale(MEANS[sc["type"]](_x_arg(X, sc.get("dims")), **({"name": sc["name"]} if "name" in sc else {}))))
        elif t["t"] == "TLocation":
            loc = t["location"]
            parts.append(g3.TLocation(MEANS[loc["type"]](_x_arg(X, loc.get("dims")),
                                                        **({"name": loc["name"]} if "name" in loc else {}))))
        else:
            parts.append(g3.TKernel(build_kernel(t["kernel"], X), noisy=t.get("noisy", False)))
    tr = parts[0]
    for t in parts[1:]:
        tr = tr @ t
    return tr


def build_mapping(mp):
    if mp["type"] == "composed":
        return build_mapping(mp["m1"]) @ build_mapping(mp["m2"])
    if mp["type"] == "invsum":
        return g3.MappingInvSum(build_mapping(mp["m1"]), build_mapping(mp["m2"]))
    mkw = {"name": mp["name"]} if "name" in mp else {}
    if "n" in mp:
        mkw["n"] = mp["n"]
    return _pot(MAPS[mp["type"]](**mkw), mp)


def build_process(spec, X, strict=True):
    kind = spec.get("kind", "gauss")
    if kind == "transport":
        return g3.TGP(X, build_transport(spec, X), strict_constants=strict,
                      **({"name": spec["name"]} if "name" in spec else {}))
    warped = spec.get("warped", spec.get("mapping", {"type": "Identity"})["type"] != "Identity")
    cls = {("gauss", False): g3.GP, ("gauss", True): g3.WGP, ("student", False): g3.TP, ("student", True): g3.WTP}[(kind, warped)]
    loc = spec.get("location", {"type": "Zero"})
    lkw = {"name": loc["name"]} if "name" in loc else {}
    if loc["type"] == "Power":
        location = _pot(g3.Power(_x_arg(X, loc.get("dims")), n=loc.get("n", 2), **lkw), loc)
    elif loc["type"] == "BlackBox":
        location = g3.BlackBox(loc["element"], _x_arg(X, loc.get("dims")), **lkw)
    else:
        location = _pot(MEANS[loc["type"]](_x_arg(X, loc.get("dims")), **lkw), loc)
    mapping = build_mapping(spec.get("mapping", {"type": "Identity"}))
    kw = {}
    if "name" in spec:
        kw["name"] = spec["name"]
    return cls(X, location, build_kernel(spec["kernel"], X), mapping, noisy=spec.get("noisy", True),
               strict_constants=strict, **kw)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def scaled_err(a, b):
    """max |a-b| / max|b| — for vectors whose small entries are sums of large cancelling terms."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) if a.size else 0.0
