"""Social-GAN pooling (`--pool_type sgan`; reference: mggan/model/modules/social_gan.py:157-229 `PoolHiddenNet`).

    pooled_a = max_b  mlp_pre_pool([spatial_embedding(pos_b - pos_a), h_b]),   a, b in the same scene (b = a included)

Same parameter names as the reference (`spatial_embedding`, `mlp_pre_pool.0`, `mlp_pre_pool.2`).  An ablation variant,
not a new kernel: the arithmetic runs in `mggan_linear_*` with the embedding folded into the first layer,

    W1 [We (pos_b - pos_a) + be ; h_b] + b1  =  Q_b - R_a,   Q = [h | pos] [W1h | W1e We]^T + (W1e be + b1),   R = pos (W1e We)^T

so the first layer is evaluated per AGENT (N rows) and only the second layer per ordered in-scene pair (sum n_s^2 rows);
pair gather, ReLU and the max over b are elementwise / index glue.  Only in-scene pairs are evaluated.
"""
import torch
from torch import nn

from mggan import kernels as K
from mggan.utils import make_mlp


class PoolHiddenNet(nn.Module):
    def __init__(self, embedding_dim=64, h_dim=64, mlp_dim=1024, bottleneck_dim=1024, activation="relu",
                 batch_norm=False, dropout=0.0):
        super().__init__()
        if activation != "relu" or batch_norm or dropout:
            raise NotImplementedError("B200 path: PoolHiddenNet with ReLU, no batch norm, no dropout (the reference's use)")
        self.mlp_dim, self.h_dim, self.bottleneck_dim, self.embedding_dim = mlp_dim, h_dim, bottleneck_dim, embedding_dim
        self.spatial_embedding = nn.Linear(2, embedding_dim)
        self.mlp_pre_pool = make_mlp([embedding_dim + h_dim, h_dim, bottleneck_dim], activation=activation,
                                     batch_norm=batch_norm, dropout=dropout)

    def forward(self, in_xy, in_dxdy, h_states, seq_start_end):
        """in_xy (T, N, 2), h_states (N, h_dim), seq_start_end list of [start, end) tiling rows 0..N-1 -> (N, bottleneck)."""
        scenes = K.SceneIndex.get(seq_start_end, h_states.device)
        h = h_states.reshape(-1, self.h_dim)
        if scenes.n_agents != h.shape[0] or (scenes.sub_batches and scenes.sub_batches[0][0] != 0):
            raise NotImplementedError("sub_batches must tile rows 0..N-1")
        pos = in_xy[-1]
        E = self.embedding_dim
        l1, l2 = self.mlp_pre_pool[0], self.mlp_pre_pool[2]
        U = l1.weight[:, :E] @ self.spatial_embedding.weight                        # (h_dim, 2)
        c = l1.weight[:, :E] @ self.spatial_embedding.bias + l1.bias
        zeros = pos.new_zeros(pos.shape[0], 2)                                      # pad the K dimension to a multiple of 4
        Q = K.linear(torch.cat([h, pos, zeros], 1), torch.cat([l1.weight[:, E:], U, U.new_zeros(U.shape)], 1), c)
        R = pos[:, 0:1] * U[:, 0] + pos[:, 1:2] * U[:, 1]                           # (N, h_dim)
        ia, ib = scenes.pair_index()
        y = K.linear(torch.relu(Q.index_select(0, ib) - R.index_select(0, ia)), l2.weight, l2.bias)
        out = y.new_full((h.shape[0], y.shape[1]), float("-inf"))
        return out.scatter_reduce(0, ia[:, None].expand_as(y), y, reduce="amax", include_self=True)
