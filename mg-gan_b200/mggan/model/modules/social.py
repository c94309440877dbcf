"""Social-Ways attention (reference: mggan/model/modules/social.py).  Parameter containers
with the reference's names; forward = `mggan_social_attn_*` over in-scene pairs only."""
import torch
from torch import nn

from mggan import kernels as K


class AttentionPooling(nn.Module):
    def __init__(self, h_dim, f_dim):
        super().__init__()
        self.f_dim, self.h_dim = f_dim, h_dim
        self.W = nn.Linear(h_dim, f_dim, bias=True)


class EmbedSocialFeatures(nn.Module):
    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.fc = nn.Sequential(nn.Linear(input_size, 32), nn.ReLU(), nn.Linear(32, 64), nn.ReLU(),
                                nn.Linear(64, hidden_size))


class SocialAttention(nn.Module):
    """reference social.py:107-123"""

    def __init__(self, social_feat_size, hidden_size):
        super().__init__()
        if hidden_size not in (32, 64):
            raise NotImplementedError("B200 path: social attention over 32- or 64-wide hidden states")
        self.feature_embedder = EmbedSocialFeatures(3, social_feat_size)
        self.attention = AttentionPooling(hidden_size, social_feat_size)

    def forward(self, in_xy, in_dxdy, enc_h, sub_batches):
        """in_xy (T,N,2), in_dxdy (T-1,N,2), enc_h (N,H), sub_batches list of [start,end) -> (N,H).
        Rows outside every range and single-agent scenes get 0."""
        scenes = K.SceneIndex.get(sub_batches, enc_h.device)
        fc = self.feature_embedder.fc
        n = scenes.n_agents
        if n == enc_h.shape[0] and (not scenes.sub_batches or scenes.sub_batches[0][0] == 0):
            return K.social_attention(in_xy[-1], in_dxdy[-1], enc_h, scenes, fc[0], fc[2], fc[4], self.attention.W)
        raise NotImplementedError("sub_batches must tile rows 0..N-1")
