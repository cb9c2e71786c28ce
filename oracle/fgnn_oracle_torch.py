"""PyTorch restatement of the reference hot path in the reference's OWN OP ORDER -- TEST / BENCH INFRASTRUCTURE.

`mp_conv_forward_torch` is /root/reference/lib/model/mpnn/mp_nn.py:92-175 written out op for op
(permute -> mm -> int64 index repeat -> gather -> bmm -> aggregator -> bias -> eval BatchNorm -> ReLU), including
`to_edge_feature`'s `index.repeat` expansion (mp_nn.py:105-111) that makes the reference's gather read an
O*T-wide int64 index per element.  It exists so that the reference's ATen op chain can be timed where
/root/reference is absent (the GPU box): on the host cores (`cpu_baseline.kind = "port-torch"`) and on the B200
itself (`gpu_aten_baseline`).  Only tests/, bench.py's baseline legs and
__graft_entry__.smoke() may import it; the product package never does.

Pinned against the same golden vectors as the numpy / C oracles (tests/test_oracle_golden.py): outputs of the real
reference executed in the build container.  `chunk_rows` evaluates the destinations in slices (bit-identical in
eval mode) so that configurations whose O*T-wide intermediates exceed memory still run.
"""
import torch


def to_edge_feature(node_feature, nn_idx):
    """mp_nn.py:92-113: node_feature [B,N,W], nn_idx [B,M,K] -> [B,M,K,W] through repeat + gather."""
    batch_size, npts, k = nn_idx.shape
    assert batch_size == node_feature.shape[0]                                   # mp_nn.py:100
    nidx = nn_idx.reshape(batch_size, -1).unsqueeze(2).repeat(1, 1, node_feature.shape[2])   # int64 [B, M*K, W]
    pts_knn = node_feature.gather(1, nidx).view(batch_size, npts, k, -1)
    return pts_knn


def _aggregate(nfeature, aggregator, gamma=3.0):
    if aggregator == "max":
        return torch.max(nfeature, dim=3, keepdim=True)[0]                       # mp_nn.py:73-75
    if aggregator == "softmax":
        return 1.0 / gamma * torch.logsumexp(gamma * nfeature, dim=3, keepdim=True)   # mp_nn.py:80-83
    if aggregator == "mean":
        return torch.mean(nfeature, dim=3, keepdim=True)                         # mp_nn.py:87
    if aggregator is None:
        return nfeature
    raise ValueError(aggregator)


def _core(x, nn_idx, etype, filters, extension, nou, nedge_types):
    batch_size, nin, nnodes = x.shape[0], x.shape[1], x.shape[2]
    k = nn_idx.shape[2]
    nedge_type = etype.permute(0, 2, 3, 1).contiguous().view(-1, nedge_types, 1)             # mp_nn.py:122-123
    if extension == 0:                                                                       # mp_nn.py:124-134
        node_feature = x.permute(0, 2, 3, 1).contiguous().view(-1, nin)
        node_feature = node_feature.mm(filters).view(batch_size, nnodes, -1)
        edge_feature = to_edge_feature(node_feature, nn_idx)
        edge_feature = edge_feature.view(-1, nou, nedge_types).bmm(nedge_type)
        edge_feature = edge_feature.view(batch_size, nn_idx.shape[1], k, nou)
    else:                                                                                    # mp_nn.py:136-159
        node_feature = x.permute(0, 2, 3, 1).contiguous()
        edge_feature = to_edge_feature(node_feature.view(batch_size, nnodes, nin), nn_idx)
        if extension == 2:
            edge_feature = node_feature - edge_feature
        edge_feature = torch.cat([node_feature.repeat(1, 1, k, 1), edge_feature], dim=3)
        edge_feature = edge_feature.view(-1, 2 * nin).mm(filters).view(-1, nou, nedge_types).bmm(nedge_type)
        edge_feature = edge_feature.view(batch_size, nnodes, k, nou)
    return edge_feature.permute(0, 3, 1, 2)                                                  # mp_nn.py:160


def mp_conv_forward_torch(x, nn_idx, etype, filters, bias=None, bn=None, extension=2, aggregator="softmax",
                          activation="relu", gamma=3.0, chunk_rows=0):
    """x [B,C,N,1], nn_idx [B,M,K] int64, etype [B,T,M,K], filters [C or 2C, O*T]; bn = dict(weight, bias,
    running_mean, running_var) of torch tensors or None.  Returns [B,O,M,1] ([B,O,M,K] for aggregator None)."""
    nedge_types = etype.shape[1]
    nou = filters.shape[1] // nedge_types
    M = nn_idx.shape[1]
    if extension != 0 or not chunk_rows or chunk_rows >= M:
        nfeature = _aggregate(_core(x, nn_idx, etype, filters, extension, nou, nedge_types), aggregator, gamma)
    else:
        parts = []
        for m0 in range(0, M, chunk_rows):
            sl = slice(m0, min(M, m0 + chunk_rows))
            parts.append(_aggregate(_core(x, nn_idx[:, sl], etype[:, :, sl], filters, 0, nou, nedge_types), aggregator, gamma))
        nfeature = torch.cat(parts, dim=2)
    if bias is not None:
        nfeature = nfeature + bias.view(1, nou, 1, 1)                                         # mp_nn.py:165-168
    if bn is not None:                                                                       # mp_nn.py:169-170 (eval)
        nfeature = torch.nn.functional.batch_norm(nfeature, bn["running_mean"], bn["running_var"], bn.get("weight"),
                                                  bn.get("bias"), False, 0.1, 1e-5)
    if activation == "relu":
        nfeature = torch.relu(nfeature)                                                      # mp_nn.py:172-173
    return nfeature
