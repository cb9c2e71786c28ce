"""Shared helpers for the test-suite (fixtures loading, tolerances)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# north_star tolerance: 1e-4 relative, fp32.  "Relative" is taken against the magnitude of the
# reference tensor: |got - ref| <= RTOL * max(|ref|, max|ref| of the tensor) -- elementwise rtol
# with the tensor's own scale as the floor (outputs pass through ReLU, so exact zeros are common).
RTOL = 1e-4


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = max(float(np.abs(ref).max()) if ref.size else 0.0, 1e-30)
    return float((np.abs(got - ref) / np.maximum(np.abs(ref), scale)).max()) if ref.size else 0.0


def assert_close(got, ref, rtol=RTOL, what=""):
    assert np.asarray(got).shape == np.asarray(ref).shape, (what, np.asarray(got).shape, np.asarray(ref).shape)
    e = rel_err(got, ref)
    assert e <= rtol, f"{what}: relative error {e:.3e} > {rtol:g}"
    return e


def load_unit_cases():
    z = np.load(os.path.join(GOLDEN, "mp_conv_v2_cases.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    cases = []
    for c in meta:
        n = c["name"]
        sd = {k[len(n) + 4:]: z[k] for k in z.files if k.startswith(n + "/sd/")}
        cases.append(dict(meta=c, x=z[f"{n}/x"], idx=z[f"{n}/idx"], etype=z[f"{n}/etype"],
                          out=z[f"{n}/out"], sd=sd))
    return cases


def load_npz(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


def sub_sd(d, prefix):
    p = prefix + "/"
    return {k[len(p):]: v for k, v in d.items() if k.startswith(p)}


def bn_of(sd):
    if "bn.running_mean" not in sd:
        return None
    return {"weight": sd["bn.weight"], "bias": sd["bn.bias"], "running_mean": sd["bn.running_mean"],
            "running_var": sd["bn.running_var"]}
