"""fgnn_b200 -- B200-native FGNN message passing (the Variable->Factor / Factor->Variable layer).

The directory is named after the reference repo (`factor-graph-neural-network_b200`, not a valid
Python identifier); import it as `fgnn_b200` (the shim package at the repo root) or load this
file by path.  Public surface = the reference's `lib.model.mpnn` names for this path:

    from fgnn_b200 import mp_conv_v2, mp_conv_type, mp_conv_residual, FactorNN, FVModule, mp_sequential

`install()` makes the reference's own scripts use the native core without edits.
"""
from . import _lib
from ._lib import FgnnError
from .build import build
from .mp_nn import (SourcePlan, base_mp_nn, check_async_errors, clear_table_cache, emodel_forward, invalidate_caches, mp_conv_residual, mp_conv_type,
                    mp_conv_v2, mp_forward)
from .factor_nn import (FactorNN, FVModule, factor_mpnn, iid_mapping, iid_mapping_bn, iid_mapping_in,
                        mp_sequential)

__version__ = "0.1.0"


def library_path():
    return _lib.LIB_PATH


def launch_count():
    """Kernels launched by libfgnn_b200.so since it was loaded."""
    return int(_lib.lib().fgnn_launch_count())


def set_programmatic_launch(enabled=True):
    """Programmatic dependent launch between consecutive tensor-core launches (default on): launch
    i+1 overlaps its set-up with the tail of launch i and orders itself behind it before touching
    x / out, so results do not change.  Returns the previous setting."""
    return bool(_lib.lib().fgnn_set_programmatic_launch(1 if enabled else 0))


def install(reference_mpnn=None):
    """Drop the native core into the reference package: after this call
    `lib.model.mpnn.mp_conv_v2` (and the copies `mp_nn_residual`, `factor_mpnn_sp`, `factor_mpnn`
    bound at import time) are fgnn_b200.mp_conv_v2, so train_syn_*.py / train_ldpc.py build their
    models on the sm_100a kernel unchanged.  `reference_mpnn` is the imported `lib.model.mpnn`
    module (default: import it).  The reference's `base_mp_nn` is kept as an extra base so its
    isinstance dispatch (sequential.py:28, factor_mpnn_sp.py:116) still sees our module."""
    import importlib
    import sys
    if reference_mpnn is None:
        reference_mpnn = importlib.import_module("lib.model.mpnn")
    ref_base = sys.modules[reference_mpnn.__name__ + ".base_model"].base_mp_nn
    ref_enum = sys.modules[reference_mpnn.__name__ + ".mp_nn"].mp_conv_type

    class mp_conv_v2_native(mp_conv_v2, ref_base):
        def __init__(self, nin, nou, nedge_types, bias=True, bn=True, extension=ref_enum.ORIG_WITH_DIFF,
                     activation_fn='relu', aggregtor='softmax'):
            if not isinstance(extension, (ref_enum, mp_conv_type)):
                raise ValueError("extension must one of mp_conv_type")
            mp_conv_v2.__init__(self, nin, nou, nedge_types, bias=bias, bn=bn,
                                extension=mp_conv_type(extension.value), activation_fn=activation_fn,
                                aggregtor=aggregtor)
            self.extension = extension if isinstance(extension, ref_enum) else ref_enum(extension.value)

    mp_conv_v2_native.__name__ = mp_conv_v2_native.__qualname__ = "mp_conv_v2"
    for name in ("mp_nn", "mp_nn_residual", "factor_mpnn_sp", "factor_mpnn", "sequential", "ensemble"):
        mod = sys.modules.get(reference_mpnn.__name__ + "." + name)
        if mod is not None and hasattr(mod, "mp_conv_v2"):
            setattr(mod, "mp_conv_v2", mp_conv_v2_native)
    reference_mpnn.mp_conv_v2 = mp_conv_v2_native
    return mp_conv_v2_native
