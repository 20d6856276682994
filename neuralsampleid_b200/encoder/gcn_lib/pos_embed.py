"""Relative position table used only to populate the (never read) ``relative_pos``
state_dict entries of ``Grapher`` so reference checkpoints load key-for-key.

Interface of the reference's encoder/gcn_lib/pos_embed.py:9-19 (``get_2d_relative_pos_embed``);
the table is 2 * E E^T / D for the standard 2-D sin-cos embedding E (grid_size^2, D).
Construction-time host code, not part of the hot path."""
import numpy as np


def _axis_embedding(dim: int, coords: np.ndarray) -> np.ndarray:
    if dim % 2:
        raise AssertionError("embedding dimension must be even")
    freq = 10000.0 ** (-np.arange(dim // 2, dtype=np.float64) / (dim / 2.0))
    phase = coords.reshape(-1).astype(np.float64)[:, None] * freq[None, :]
    return np.concatenate([np.sin(phase), np.cos(phase)], axis=1)


def get_2d_sincos_pos_embed(embed_dim: int, grid_size: int, cls_token: bool = False) -> np.ndarray:
    ax = np.arange(grid_size, dtype=np.float32)
    gw, gh = np.meshgrid(ax, ax)                       # w varies fastest
    emb = np.concatenate([_axis_embedding(embed_dim // 2, gw), _axis_embedding(embed_dim // 2, gh)],
                         axis=1)
    if cls_token:
        emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return emb


def get_2d_relative_pos_embed(embed_dim: int, grid_size: int) -> np.ndarray:
    emb = get_2d_sincos_pos_embed(embed_dim, grid_size)
    return 2.0 * (emb @ emb.T) / emb.shape[1]
