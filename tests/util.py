"""Shared helpers for the test-suite (fixtures loading, tolerances)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# north_star tolerance: 1e-4 relative, fp32.  Stated ELEMENTWISE with an explicit absolute floor:
#     |got - ref| <= rtol * |ref| + ATOL_FRAC * rtol * scale,      scale = max |ref| over the tensor,
# i.e. for rtol = 1e-4 an element may be off by 1e-4 of its own magnitude plus 2e-5 of the tensor's scale (outputs
# pass through ReLU, so exact zeros are common and need the floor; the kernels' measured error is 4-7e-6 of scale).
# `rel_err` reports the worst element in units of that bound times rtol, so `rel_err <= rtol` is the criterion.
RTOL = 1e-4
ATOL_FRAC = 0.2


def rel_err(got, ref, rtol=RTOL):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if not ref.size:
        return 0.0
    scale = max(float(np.abs(ref).max()), 1e-30)
    bound = np.abs(ref) + ATOL_FRAC * scale                    # the admissible error per element, divided by rtol
    return float((np.abs(got - ref) / bound).max())


def error_profile(got, ref):
    """Elementwise error distribution (for logs): max / 99.9th percentile of |d| / scale and of |d| / |ref| over the
    elements with |ref| > 1e-3 * scale."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = max(float(np.abs(ref).max()), 1e-30)
    d = np.abs(got - ref)
    big = np.abs(ref) > 1e-3 * scale
    rel = d[big] / np.abs(ref[big]) if big.any() else np.zeros(1)
    return dict(max_abs_over_scale=float(d.max() / scale), p999_abs_over_scale=float(np.quantile(d, 0.999) / scale),
                max_rel=float(rel.max()), p999_rel=float(np.quantile(rel, 0.999)))


def assert_close(got, ref, rtol=RTOL, what=""):
    assert np.asarray(got).shape == np.asarray(ref).shape, (what, np.asarray(got).shape, np.asarray(ref).shape)
    e = rel_err(got, ref, rtol)
    assert e <= rtol, f"{what}: elementwise error {e:.3e} x (|ref| + {ATOL_FRAC} scale) exceeds rtol {rtol:g}; {error_profile(got, ref)}"
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
