"""Import the *real* reference modules (container only; /root/reference is absent on the GPU box).

TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Used by tests/golden/make_golden.py to generate golden vectors and by the optional
`-m "not gpu"` cross-checks that run only when /root/reference exists.

The recipe skips lib/__init__.py (it imports lib.data, which needs the un-built MNC pybind
module and the `ad3` package) by registering a stub `lib` package whose __path__ points at the
reference tree, so `lib.model.mpnn` imports cleanly (reference: lib/__init__.py:1,
lib/data/__init__.py:1-2).
"""
import ast
import contextlib
import io
import os
import sys
import types

REF_ROOT = os.environ.get("FGNN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "model", "mpnn"))


def load():
    """Returns the reference `lib.model.mpnn` module (imports it on first call)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import numpy as np
    import torch.nn.functional as F
    if not hasattr(np, "int"):
        np.int = int  # lib/data/ldpc_dataset.py:51 uses the removed alias
    # PyTorch-1.0 behaviour for InstanceNorm on a single spatial element (LDPC "global" factor):
    F._verify_spatial_size = lambda size: None
    if "lib" not in sys.modules or not hasattr(sys.modules["lib"], "__fgnn_stub__"):
        pkg = types.ModuleType("lib")
        pkg.__path__ = [os.path.join(REF_ROOT, "lib")]
        pkg.__fgnn_stub__ = True
        sys.modules["lib"] = pkg
        d = types.ModuleType("lib.data")
        d.__path__ = [os.path.join(REF_ROOT, "lib", "data")]
        sys.modules["lib.data"] = d
        m = types.ModuleType("lib.data.MNC")
        for n in ("s2t", "t2y", "y2b", "zb2x", "init_seed"):
            setattr(m, n, lambda *a, **k: None)
        sys.modules["lib.data.MNC"] = m
    import importlib
    with contextlib.redirect_stdout(io.StringIO()):
        mod = importlib.import_module("lib.model.mpnn")
    return mod


def quiet(fn, *a, **k):
    """Call fn with stdout suppressed (mp_conv_v2.__init__ prints the aggregator, mp_nn.py:69)."""
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def script_functions(script: str, names):
    """ast-extract top-level function defs from a train_*.py script (they import tensorboardX at
    module level, which is not installed) and exec them with numpy/torch in scope."""
    import numpy as np
    import torch
    src = open(os.path.join(REF_ROOT, script)).read()
    tree = ast.parse(src)
    ns = {"np": np, "torch": torch}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), script, "exec")
            exec(code, ns)
    return {n: ns[n] for n in names}


def ldpc_structure():
    """Returns the reference's ldpc_graph_structure_generator class (lib/data/ldpc_dataset.py)."""
    load()
    import importlib
    mod = importlib.import_module("lib.data.ldpc_dataset")
    return mod.ldpc_graph_structure_generator
