"""The callers that define "one FGNN layer", rebuilt around the native message-passing core.

Mirrors of /root/reference/lib/model/mpnn/factor_mpnn_sp.py (`FVModule` :14-22, `FactorNN` :25-178),
base_model.py (`iid_mapping*` :43-90) and sequential.py (`mp_sequential` :21-39) with the same
constructor arguments, forward signatures and state_dict keys, so a checkpoint written by the
reference loads here and vice versa.  Only the `mp_conv_v2` inside them differs: it is
fgnn_b200.mp_conv_v2 (one sm_100a kernel per call).  The per-node 1x1 maps and norms either side
of the core stay in PyTorch.
"""
import torch

from .mp_nn import base_mp_nn, conv1x1, conv1x1_native, mp_conv_residual, mp_conv_type, mp_conv_v2


def _as_index(t):
    return t if t.dtype in (torch.int64, torch.int32) else t.long()


class iid_mapping(torch.nn.Module):
    """1x1 conv + LeakyReLU per node (base_model.py:43-60)."""

    def __init__(self, nin, nout, bias=True):
        super().__init__()
        self.main = torch.nn.Sequential(torch.nn.Conv2d(nin, nout, 1, bias=bias),
                                        torch.nn.LeakyReLU(inplace=True))

    def forward(self, x):
        return self.main(x)


class iid_mapping_bn(torch.nn.Module):
    """1x1 conv + BatchNorm + ReLU per node (base_model.py:63-80)."""

    def __init__(self, nin, nout, bias=True, bn=True):
        super().__init__()
        self.main = torch.nn.Sequential(torch.nn.Conv2d(nin, nout, 1, bias=bias),
                                        torch.nn.BatchNorm2d(nout), torch.nn.ReLU(inplace=True))

    def forward(self, x):
        return self.main(x)


class _InstanceNorm2d(torch.nn.InstanceNorm2d):
    """InstanceNorm2d that also accepts a single spatial element (the LDPC "global" factor has
    one): PyTorch 1.0, which the reference targets, returned (x - x)/sqrt(eps) = 0 there, newer
    versions raise (SURVEY 8c shim 1)."""

    def forward(self, x):
        if x.shape[-1] * x.shape[-2] == 1 and not self.track_running_stats:
            return torch.zeros_like(x)
        return super().forward(x)


def instance_norm_act(x, eps=1e-5, activation=None):
    """InstanceNorm2d (affine = False) + activation (None | 'relu') on a CUDA fp32 [B,C,N,1] tensor in ONE pass of the
    library's kernel (`fgnn_instance_norm_forward`); memory format of x is kept.  PyTorch's instance_norm spends
    2.7 ms per call on the LDPC shapes ([4096, 64..256, 96, 1]: its statistics kernel reduces 96 elements per block);
    it was 47 % of a FactorNN forward once the message-passing cores ran on tensor cores."""
    import ctypes
    from . import _lib
    B, C, N, W = x.shape
    assert W == 1
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.lib().fgnn_instance_norm_forward(
            ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), B, C, N, x.stride(0), x.stride(1), x.stride(2),
            out.stride(0), out.stride(1), out.stride(2), float(eps), _lib.ACT_RELU if activation == "relu" else _lib.ACT_NONE,
            0.0, ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    _lib.check(rc, "instance_norm_forward")
    return out


class iid_mapping_in(torch.nn.Module):
    """1x1 conv + InstanceNorm + ReLU per node (base_model.py:83-90)."""

    def __init__(self, nin, nout, bias=True):
        super().__init__()
        self.main = torch.nn.Sequential(torch.nn.Conv2d(nin, nout, 1, bias=bias),
                                        _InstanceNorm2d(nout), torch.nn.ReLU())

    def forward(self, x):
        norm = self.main[1]
        if (x.dim() == 4 and x.shape[2] * x.shape[3] == 1 and isinstance(norm, _InstanceNorm2d)
                and not norm.track_running_stats and not norm.affine):
            # a single spatial element (the LDPC "global" factor): the norm returns zeros whatever the map computes
            # (_InstanceNorm2d above), so the map is not run at all
            return torch.zeros((x.shape[0], self.main[0].out_channels, 1, 1), dtype=x.dtype, device=x.device)
        y = conv_in_relu(self.main[0], norm, x, self.training)
        return y if y is not None else self.main(x)


def conv_in_relu(conv, norm, x, training=False):
    """Conv2d(1x1) -> InstanceNorm2d(affine = False) -> ReLU on a CUDA fp32 [B,C,N,1] tensor as two native passes (the map
    on the tensor-core kernel when the shape qualifies, norm + ReLU in one kernel); None when the input does not qualify
    (then the caller runs its own nn.Sequential).  Forward-only, like the core."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[3] == 1 and x.shape[2] > 1
            and isinstance(conv, torch.nn.Conv2d) and conv.kernel_size == (1, 1)
            and isinstance(norm, torch.nn.InstanceNorm2d) and not norm.affine and not norm.track_running_stats
            and not (training and torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad))):
        return None
    with torch.no_grad():
        y = conv1x1_native(x, conv.weight, conv.bias)                      # tensor-core pass when the shape qualifies
        if y is None:
            y = conv1x1(x, conv.weight, conv.bias)
    return instance_norm_act(y, norm.eps, "relu")                           # norm + ReLU: one native pass


class mp_sequential(base_mp_nn):
    """nn.Sequential that passes (x, *graph) to message-passing children (sequential.py:9-39)."""

    def __init__(self, *module_list):
        super().__init__()
        self.module_list = []
        for i, mod in enumerate(module_list):
            self.add_module(str(i), mod)
            self.module_list.append(mod)

    def forward(self, node_feature, *argv):
        extra = []
        for m in self.module_list:
            node_feature = m(node_feature, *argv) if isinstance(m, base_mp_nn) else m(node_feature)
            if isinstance(node_feature, tuple):
                extra += list(node_feature[1:])
                node_feature = node_feature[0]
        return node_feature if not extra else (node_feature, extra)


class FVModule(torch.nn.Module):
    """The FV module: mp_conv_v2(NO_EXTENSION, max) (factor_mpnn_sp.py:14-22)."""

    def __init__(self, nin, nout, nedge_types, with_bn=True):
        super().__init__()
        self.main_module = mp_conv_v2(nin, nout, nedge_types, bn=with_bn,
                                      extension=mp_conv_type.NO_EXTENSION, aggregtor='max')

    def forward(self, n_or_f_feature, nn_idx, etype):
        return self.main_module(n_or_f_feature, nn_idx, etype)


class FactorNN(torch.nn.Module):
    """Stacked FGNN with split Variable->Factor / Factor->Variable tables (factor_mpnn_sp.py:25-178).

    Per layer `i` and factor type `j`:  nfeature = v2v_i(x_v) + sum_j f2v_ij(x_fj, idx_f2v_j, et_f2v_j);
    nffeature_j = f2f_ij(x_fj) + v2f_ij(x_v, idx_v2f_j, et_v2f_j); residual when nin == nout; skip links.
    """

    def __init__(self, node_feature_dim, factor_feature_dim_list, dim_mapping_list, netype_list,
                 nclass=2, gnn_immediate_dim=64, max_mpnn_dim=128, final_filter=None, skip_link={},
                 aggregator='max', ret_high=False):
        super().__init__()
        self.node_feature_dim = node_feature_dim
        self.map_dim = dim_mapping_list[0]
        self.final_filter = final_filter
        self.node_mapping_module = iid_mapping(node_feature_dim, self.map_dim)
        self.nfactor_types = len(factor_feature_dim_list)
        self.factor_mapping_modules = []
        for j, dim in enumerate(factor_feature_dim_list):
            m = iid_mapping_bn(dim, self.map_dim)
            self.factor_mapping_modules.append(m)
            self.add_module('factor_mapping_modules_{}'.format(j), m)
        self.f2v_modules, self.v2f_modules, self.f2f_modules, self.v2v_modules = [], [], [], []
        self.dim_mapping_list = dim_mapping_list
        none = mp_conv_type.NO_EXTENSION
        for i in range(len(dim_mapping_list) - 1):
            nin, nout = dim_mapping_list[i], dim_mapping_list[i + 1]
            self.v2v_modules.append(iid_mapping_in(nin, nout))
            self.add_module('v2v_{}'.format(i), self.v2v_modules[-1])
            f2v, v2f, f2f = [], [], []
            for j in range(self.nfactor_types):
                netype = netype_list[j]
                f2f.append(iid_mapping_in(nin, nout))

                def core():
                    if nin == nout:                                    # factor_mpnn_sp.py:79-83
                        return mp_conv_residual(nin, gnn_immediate_dim, netype, extension=none,
                                                with_residual=False, aggregator=aggregator)
                    if nin <= max_mpnn_dim and nout <= max_mpnn_dim:   # :85-89
                        return mp_conv_v2(nin, nout, netype, extension=none, aggregtor=aggregator)
                    return mp_conv_residual(nin, gnn_immediate_dim, netype, extension=none,   # :90-94
                                            with_residual=False, nout=nout, aggregator=aggregator)
                f2v.append(core())
                v2f.append(core())
                self.add_module('f2v_{}_{}'.format(i, j), f2v[-1])
                self.add_module('v2f_{}_{}'.format(i, j), v2f[-1])
                self.add_module('f2f_{}_{}'.format(i, j), f2f[-1])
            self.f2f_modules.append(f2f)
            self.f2v_modules.append(f2v)
            self.v2f_modules.append(v2f)
        self.skip_link = skip_link
        self.ret_high = ret_high
        final_dim = nclass if nclass > 2 else 1
        self.final_classifier = torch.nn.Sequential(
            torch.nn.Conv2d(dim_mapping_list[-1], 128, 1), _InstanceNorm2d(128),
            torch.nn.ReLU(inplace=True), torch.nn.Conv2d(128, final_dim, 1, bias=True))

    def mpnn_forward(self, mpnn, node_feature, nn_idx, efeature):
        if isinstance(mpnn, base_mp_nn):
            return mpnn(node_feature, nn_idx, efeature)
        return mpnn(node_feature)

    @staticmethod
    def _add_core(acc, mpnn, node_feature, nn_idx, efeature):
        """acc + mpnn(...) (factor_mpnn_sp.py:146-151); the native modules add into `acc` inside their last kernel's store
        (`acc` is always a fresh tensor here: the output of this layer's v2v / f2f map)."""
        if isinstance(mpnn, (mp_conv_v2, mp_conv_residual)):
            return mpnn(node_feature, nn_idx, efeature, add_to=acc)
        if isinstance(mpnn, base_mp_nn):
            return acc + mpnn(node_feature, nn_idx, efeature)
        return acc + mpnn(node_feature)

    def forward(self, node_feature, hop_features, nn_idx_f2v, nn_idx_v2f, etype_f2v, etype_v2f):
        x_v = self.node_mapping_module(node_feature)
        x_f = [m(f) for f, m in zip(hop_features, self.factor_mapping_modules)]
        if x_v.is_cuda:
            # node-major (channels_last) from here on: it is what the native cores read and write, their 1x1
            # neighbours then run as row-major GEMMs (mp_nn.conv1x1) and no call needs a layout pass
            x_v = x_v.contiguous(memory_format=torch.channels_last)
            x_f = [f.contiguous(memory_format=torch.channels_last) for f in x_f]
        history = []
        for i in range(len(self.v2f_modules)):
            nin, nout = self.dim_mapping_list[i], self.dim_mapping_list[i + 1]
            new_v = self.v2v_modules[i](x_v)
            new_f = [m(f) for f, m in zip(x_f, self.f2f_modules[i])]
            for j in range(len(self.f2v_modules[i])):
                # the reference casts with .long() (factor_mpnn_sp.py:145,150); int32 tables go through as they are
                # (the core takes both) so the per-table validation / plan caches keep hitting on the caller's object
                new_v = self._add_core(new_v, self.f2v_modules[i][j], x_f[j], _as_index(nn_idx_f2v[j]), etype_f2v[j])
                new_f[j] = self._add_core(new_f[j], self.v2f_modules[i][j], x_v, _as_index(nn_idx_v2f[j]), etype_v2f[j])
            if nin == nout:
                x_v = x_v + new_v
                x_f = [a + b for a, b in zip(new_f, x_f)]
            else:
                x_v, x_f = new_v, new_f
            if i in self.skip_link.keys():
                old_v, old_f = history[self.skip_link[i]]
                x_v = x_v + old_v
                x_f = [a + b for a, b in zip(old_f, x_f)]
            history.append([x_v, x_f])
        fc = self.final_classifier
        head = conv_in_relu(fc[0], fc[1], x_v, self.training)             # Conv -> InstanceNorm -> ReLU natively
        res = fc[3](head) if head is not None else fc(x_v)
        if self.final_filter is not None:
            res = self.final_filter(res, node_feature)
        return (res, x_f) if self.ret_high else res


class factor_mpnn(torch.nn.Module):
    """Stacked FGNN over MERGED tables (factor_mpnn.py:8-133; train_syn_pw_factor.py / train_syn_hop_factor.py): per
    layer and factor type the variable and factor features are concatenated along the node axis and ONE core call
    does the Variable->Factor and the Factor->Variable step together (`nn_idx [B, N+F, K]`: the first N rows name
    factor rows, the last F rows variable rows); the per-type variable features are concatenated along channels and
    merged by a 1x1 map.  Same constructor arguments, forward signature and state_dict keys as the reference; the
    cores are fgnn_b200 modules (ORIG_WITH_DIFF, which runs on the tensor-core kernel through the two-atom row
    [x_m || x_idx] against a transformed filter image)."""

    def __init__(self, node_feature_dim, factor_feature_dim_list, dim_mapping_list, netype_list,
                 gnn_immediate_dim=64, max_mpnn_dim=64, final_filter=None, skip_link={}):
        super().__init__()
        self.node_feature_dim = node_feature_dim
        self.map_dim = dim_mapping_list[0]
        self.mapping_modules = [iid_mapping(node_feature_dim, self.map_dim)]
        self.nfactor_types = len(factor_feature_dim_list)
        for dim in factor_feature_dim_list:
            self.mapping_modules.append(iid_mapping(dim, self.map_dim))
        for i, m in enumerate(self.mapping_modules):
            self.add_module('mapping_modules_{}'.format(i), m)
        self.mp_nn_modules, self.mp_merge_modules = [], []
        self.final_filter = final_filter
        for i in range(len(dim_mapping_list) - 1):
            nin, nout = dim_mapping_list[i], dim_mapping_list[i + 1]
            row = []
            for j in range(self.nfactor_types):
                if nin == nout:                                            # factor_mpnn.py:55-57
                    m = mp_conv_residual(nin, gnn_immediate_dim, netype_list[j])
                elif nin <= max_mpnn_dim and nout <= max_mpnn_dim:         # :59-60
                    m = mp_conv_v2(nin, nout, netype_list[j])
                else:                                                      # :62-65
                    m = torch.nn.Sequential(torch.nn.Conv2d(nin, nout, 1), torch.nn.InstanceNorm2d(nout),
                                            torch.nn.ReLU(inplace=True))
                self.add_module('mp_nn_{}_{}'.format(i, j), m)
                row.append(m)
            self.mp_nn_modules.append(row)
            if i < len(dim_mapping_list) - 2:
                merge = iid_mapping_bn(nout * self.nfactor_types, nout)
            else:
                merge = torch.nn.Sequential(
                    torch.nn.Conv2d(nout * self.nfactor_types, 256, 1, bias=True), torch.nn.BatchNorm2d(256),
                    torch.nn.LeakyReLU(), torch.nn.Conv2d(256, 256, 1, bias=True), torch.nn.LeakyReLU(),
                    torch.nn.Conv2d(256, nout, 1, bias=True))
            self.add_module('merge_module_{}'.format(i), merge)
            self.mp_merge_modules.append(merge)
        self.skip_link = skip_link

    def forward(self, node_features, factor_features, graph_structures):
        nnode = node_features.shape[2]
        x_v = self.mapping_modules[0](node_features)
        x_f = [m(f) for f, m in zip(factor_features, self.mapping_modules[1:])]
        history = []
        for i, modules in enumerate(self.mp_nn_modules):
            per_type_v, new_f = [], []
            for j, (f, m) in enumerate(zip(x_f, modules)):
                merged = torch.cat([x_v, f], dim=2)
                nn_idx, etype = graph_structures[j]
                merged = m(merged, nn_idx, etype) if isinstance(m, base_mp_nn) else m(merged.contiguous())
                per_type_v.append(merged[:, :, :nnode, :])
                new_f.append(merged[:, :, nnode:, :])
            x_v = self.mp_merge_modules[i](torch.cat(per_type_v, dim=1))
            x_f = new_f
            if i in self.skip_link.keys():
                old_v, old_f = history[self.skip_link[i]]
                x_v = x_v + old_v
                x_f = [a + b for a, b in zip(x_f, old_f)]
            history.append([x_v, x_f])
        if self.final_filter is not None:
            x_v = self.final_filter(x_v, node_features)
        return x_v, x_f
