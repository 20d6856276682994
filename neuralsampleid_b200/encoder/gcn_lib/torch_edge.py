"""B200 counterparts of the reference's encoder/gcn_lib/torch_edge.py hot-path entries:
``pairwise_distance`` (:7-18), ``dense_knn_matrix`` (:70-103), ``DenseDilated`` (:233-255),
``DenseDilatedKnnGraph`` (:262-284).  One fused kernel (grafp_knn_fwd) does normalise + distance
tiles + top-(k*d) + dilation stride; the N x N distance matrix is never materialised."""
from __future__ import annotations

import torch
from torch import nn

from ... import ops


def _edge_index(nn_idx: torch.Tensor) -> torch.Tensor:
    """(B, N, k) int32 neighbour lists -> the reference's (2, B, N, k) int64 edge_index with the
    centre index in slot 1 (torch_edge.py:102-103)."""
    B, N, k = nn_idx.shape
    center = torch.arange(N, device=nn_idx.device).view(1, N, 1).expand(B, N, k)
    return torch.stack((nn_idx.long(), center), dim=0)


def pairwise_distance(x: torch.Tensor) -> torch.Tensor:
    """Not provided as a standalone product op: the distance matrix only exists tile-by-tile inside
    the kNN kernel.  Use ``dense_knn_matrix`` (or ops.knn(return_dist=True) for the selected
    distances)."""
    raise NotImplementedError("pairwise_distance is fused into the kNN kernel (never materialised)")


def dense_knn_matrix(x: torch.Tensor, k: int = 16, relative_pos=None) -> torch.Tensor:
    """x: (B, C, N, 1) (already normalised by the caller, as in the reference) ->
    edge_index (2, B, N, k) int64, neighbours in ascending-distance order."""
    if relative_pos is not None:
        raise NotImplementedError("relative_pos is never passed on the GraphEncoder path (SURVEY Q7)")
    B, C, N = x.shape[:3]
    nodes = ops.nchw_to_nodes(x.reshape(B, C, N))
    return _edge_index(ops.knn(nodes, B, N, k, 1, normalize=False))


class DenseDilated(nn.Module):
    """edge_index[..., ::dilation] (reference :233-255); the stochastic branch is dead on this path
    (stochastic=False, graph_encoder.py:141) and is rejected."""

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation, self.stochastic, self.epsilon, self.k = dilation, stochastic, epsilon, k

    def forward(self, edge_index):
        if self.stochastic and self.training:
            raise NotImplementedError("stochastic dilation is not supported")
        return edge_index[:, :, :, ::self.dilation]


class DenseDilatedKnnGraph(nn.Module):
    """F.normalize + dense kNN of k*dilation candidates + every dilation-th rank."""

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation, self.stochastic, self.epsilon, self.k = dilation, stochastic, epsilon, k
        self._dilated = DenseDilated(k, dilation, stochastic, epsilon)

    def knn_nodes(self, x_nodes: torch.Tensor, B: int, N: int, row_sumsq=None) -> torch.Tensor:
        """node-major (B*N, C) -> int32 (B, N, k).  ``row_sumsq``: per-node sum of squares already
        produced by the GEMM that wrote x_nodes (skips the kNN's own norm pass)."""
        if self.stochastic and self.training:
            raise NotImplementedError("stochastic dilation is not supported")
        return ops.knn(x_nodes, B, N, self.k, self.dilation, normalize=True, row_sumsq=row_sumsq)

    def forward(self, x, y=None, relative_pos=None):
        if y is not None:
            raise NotImplementedError("r > 1 (pooled y) graphs are not reached by GraphEncoder (r=1)")
        if relative_pos is not None:
            raise NotImplementedError("relative_pos is never passed on the GraphEncoder path")
        B, C, N = x.shape[:3]
        nodes = ops.nchw_to_nodes(x.reshape(B, C, N))
        return _edge_index(self.knn_nodes(nodes, B, N))
