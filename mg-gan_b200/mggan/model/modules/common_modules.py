"""Trajectory encoder / decoder parameter containers (reference:
mggan/model/modules/common_modules.py).  Same class names, constructor arguments, parameter
names and shapes (so reference checkpoints load), same forward signatures; the arithmetic runs
in the sm_100a kernels (`mggan_lstm_seq_*`, `mggan_decoder_*`)."""
from collections import namedtuple

import torch
from torch import nn

from mggan import kernels as K
from mggan.utils import make_mlp

GeneratorOutput = namedtuple("generator_out", ["rel", "abs"])


def get_input(xy, dxdy, inp_format):
    """reference common_modules.py:12-21"""
    if inp_format == "rel":
        return dxdy
    if inp_format == "abs":
        return xy
    raise NotImplementedError("inp_format='abs_rel' is outside the B200 hot path (default is 'rel')")


class TrajectoryEncoder(nn.Module):
    """Linear(inp_size, embedding_dim) + 1-layer LSTM, returns h_T (reference :24-66)."""

    def __init__(self, hidden_size=128, inp_size=2, num_layers=1, embedding_dim=None, return_hc=False):
        super().__init__()
        if num_layers != 1 or embedding_dim is None or inp_size != 2 or return_hc:
            raise NotImplementedError("B200 path: 1-layer LSTM over 2-D inputs with an embedding, h_T output")
        if hidden_size not in (32, 64):
            raise NotImplementedError("B200 path: encoder hidden size 32 or 64")
        self.embedding_dim, self.inp_size, self.return_hc = embedding_dim, inp_size, return_hc
        self.embedding = nn.Linear(inp_size, embedding_dim)
        self.encoder = nn.LSTM(input_size=embedding_dim, hidden_size=hidden_size, num_layers=num_layers)

    def forward(self, inp, hc=None):
        """inp (T, N, 2) -> (N, hidden)"""
        assert hc is None
        e = self.encoder
        return K.lstm_encode(inp, self.embedding.weight, self.embedding.bias, e.weight_ih_l0, e.weight_hh_l0,
                             e.bias_ih_l0, e.bias_hh_l0)


class RelativeDecoder(nn.Module):
    """One generator's decoder weights (reference :69-131).  `forward` decodes a batch of rows
    with this generator alone; MultiGenerator drives all generators through one launch."""

    def __init__(self, pred_len=12, embedding_dim=128, h_dim=128, num_layers=1, dropout=0.0, inp_format="abs_rel",
                 z_size=64, social_feat_size=128):
        super().__init__()
        if inp_format != "rel" or num_layers != 1 or h_dim != 32 or social_feat_size != 32:
            raise NotImplementedError("B200 path: inp_format='rel', decoder_h_dim=32, social_feat_size=32")
        self.pred_len, self.h_dim, self.embedding_dim, self.inp_format = pred_len, h_dim, embedding_dim, inp_format
        self.decoder = nn.LSTM(embedding_dim, h_dim, num_layers, dropout=dropout)
        self.spatial_embedding = nn.Linear(2, embedding_dim)
        self.hidden2pos = make_mlp([h_dim + social_feat_size, h_dim // 2, 2], "leaky_relu", batch_norm=False)

    def folded(self):
        """Kernel-layout weights of this generator (embedding folded into the input projection)."""
        d = self.decoder
        H = self.h_dim
        w1 = self.hidden2pos[0].weight
        return {
            "wx": d.weight_ih_l0 @ self.spatial_embedding.weight,
            "b": d.weight_ih_l0 @ self.spatial_embedding.bias + d.bias_ih_l0 + d.bias_hh_l0,
            "whh": d.weight_hh_l0, "w1h": w1[:, :H], "w1s": w1[:, H:], "b1": self.hidden2pos[0].bias,
            "w2": self.hidden2pos[2].weight, "b2": self.hidden2pos[2].bias,
        }

    def forward(self, xy, dxdy, noise, social_feats, state_tuple):
        """xy, dxdy (R,2); social_feats (R,32); state_tuple (h0, c0) each (1,R,32), c0 must be 0
        (as in forward_all).  Returns (abs (T,R,2), rel (T,R,2)).  `noise` is unused, as in the
        reference."""
        R = xy.shape[0]
        h0 = state_tuple[0].reshape(R, self.h_dim)
        sel = K.Selection.all_generators(R, 1, 1, xy.device)
        gw = {k: v[None] for k, v in self.folded().items()}
        zero_noise = torch.zeros(R, 1, device=xy.device)
        wz = torch.zeros(self.h_dim, 1, device=xy.device)
        pabs, prel = K.decode(h0, social_feats, xy, dxdy, zero_noise, wz, gw, sel, self.pred_len)
        return pabs, prel
