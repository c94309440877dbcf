"""MultiDiscriminatorTrajectory (reference: mggan/model/modules/discriminators.py:12-263).

Observed-trajectory LSTM (H=64) + prediction MLP -> social attention -> scene attention ->
sigmoid head (+ generator-id classifier head for gan_type='mgan').  Same constructor,
parameter names and forward contract; the arithmetic runs in the sm_100a kernels.

Reference behaviour kept on purpose: the reference passes
`seq_start_end * n_samples` -- the SAME index ranges repeated -- to the social module, so
attention pooling only ever fills the rows of sample 0 and every other sample gets zeros.  The
reference still evaluates the pair MLP on all (k N)^2 row pairs and throws the result away;
here only the n_s^2 in-scene pairs of sample 0 are computed.
"""
import contextlib

import torch
import torch.nn as nn

from mggan import kernels as K
from mggan.model.modules.cnn import AttentionGlobal
from mggan.model.modules.common_modules import TrajectoryEncoder
from mggan.model.modules.social import SocialAttention
from mggan.model.modules.social_gan import PoolHiddenNet


class MultiDiscriminatorTrajectory(nn.Module):
    def __init__(self, num_gens, num_discs, unbound_output, h_dim, inp_format, pred_len, gan_type, global_disc,
                 scene_dim, pool_type="sgan"):
        super().__init__()
        assert inp_format in ("rel", "abs", "abs_rel")
        assert gan_type in ("probgan", "mgan", "infogan", "gan")
        assert pool_type in ("sways", "sgan")
        if (inp_format != "rel" or gan_type not in ("mgan", "gan") or not global_disc or num_discs != 1 or h_dim != 64):
            raise NotImplementedError(
                "B200 path covers the default discriminator: inp_format='rel', gan_type in {'mgan','gan'}, "
                "global_disc=1, one head, h_dim=64")
        self.pool_type = pool_type
        if scene_dim not in (0, 64):
            raise NotImplementedError("scene_dim must be 0 or 64")
        self.inp_format, self.unbound_output, self.n_ds = inp_format, unbound_output, num_discs
        self.gan_type, self.global_disc, self.inp_size = gan_type, global_disc, 2
        self.in_encoder = TrajectoryEncoder(hidden_size=h_dim, inp_size=2, num_layers=1, embedding_dim=h_dim,
                                            return_hc=False)
        self.in_encoder_fc = nn.Sequential(nn.Linear(h_dim, h_dim // 2), nn.LeakyReLU(0.2),
                                           nn.Linear(h_dim // 2, h_dim // 2))
        self.pred_encoder = nn.Sequential(nn.Linear(pred_len * 2, h_dim), nn.LeakyReLU(0.2),
                                          nn.Linear(h_dim, h_dim // 2))
        if pool_type == "sways":
            self.social = SocialAttention(h_dim, h_dim)
        else:                                        # reference discriminators.py:62-67
            self.social = PoolHiddenNet(embedding_dim=16, h_dim=h_dim, mlp_dim=h_dim, bottleneck_dim=h_dim)
        h_dim *= 2
        if scene_dim > 0:
            self.scene_encoder = AttentionGlobal(noise_attention_dim=0, PhysFeature=True, num_layers=2, channels_cnn=8)
            h_dim += scene_dim
        self.discs = nn.ModuleList()
        for _ in range(num_discs):
            self.discs.append(nn.Sequential(nn.Linear(h_dim, h_dim // 2), nn.LeakyReLU(0.2), nn.Linear(h_dim // 2, 1),
                                            nn.Sigmoid()))
        if gan_type == "mgan":
            self.gen_id_reconstructor = nn.Sequential(nn.Linear(h_dim, h_dim // 2), nn.LeakyReLU(0.2),
                                                      nn.Linear(h_dim // 2, num_gens))
        self.eps = 1e-7
        self.len_hist = 1.0
        self._obs_memo = None

    @contextlib.contextmanager
    def share_observed(self):
        """While active, the features that depend only on the observed inputs and the weights -- the
        observed-trajectory encoding `in_encoder_fc(in_encoder(in_dxdy))` and the scene features -- are computed
        once per distinct input tensor and shared by every `forward` (the discriminator step runs D on real and
        on fake futures of the SAME observations with the SAME weights, train.py:150,171).  The shared tensors
        are single autograd nodes, so the gradients of all uses add up exactly as in the reference; BatchNorm
        running statistics are still updated once per forward."""
        enc = getattr(self, "scene_encoder", None)
        self._obs_memo = {}
        if enc is not None:
            enc.memo = {}
        try:
            yield self
        finally:
            self._obs_memo = None
            if enc is not None:
                enc.memo = None

    @staticmethod
    def _mlp2(seq, x, last_act=K.ACT_NONE):
        return K.mlp2(x, seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias, K.ACT_LRELU, 0.2, last_act)

    def _in_enc(self, in_dxdy):
        key = (id(in_dxdy), in_dxdy._version, torch.is_grad_enabled())
        in_enc = self._obs_memo.get(key) if self._obs_memo is not None else None
        if in_enc is None:
            in_enc = self._mlp2(self.in_encoder_fc, self.in_encoder(in_dxdy))
            if self._obs_memo is not None:
                self._obs_memo[key] = in_enc
        return in_enc

    def _pred_enc(self, pred_dxdy):
        pred_len, n_samples, b, _ = pred_dxdy.shape
        pv = pred_dxdy.permute(1, 2, 0, 3).reshape(n_samples * b, pred_len * 2)
        return self._mlp2(self.pred_encoder, pv)

    def encode(self, in_xy, in_dxdy, pred_xy, pred_dxdy, mask=None):
        """-> (k * N, 64), row = sample * N + agent (reference :113-142)."""
        in_enc = self._in_enc(in_dxdy)
        pred_len, n_samples, b, _ = pred_xy.shape
        N = in_xy.size(1)
        pred_enc = self._pred_enc(pred_dxdy)
        if mask is not None:
            padded = torch.zeros(N * n_samples, pred_enc.size(1), device=pred_enc.device)
            padded[mask.repeat(n_samples)] = pred_enc
            pred_enc = padded
        return torch.cat([in_enc.repeat(n_samples, 1), pred_enc], dim=1)

    def _scene(self, img, mask):
        rows, rows_key = None, None
        if mask is not None:
            rows = torch.nonzero(mask).flatten().to(torch.int32)
            rows_key = ("mask", id(mask), mask._version)
        return self.scene_encoder(img, rows, rows_key)

    def _heads_frozen(self):
        ps = list(self.discs[0].parameters())
        if self.gan_type == "mgan":
            ps += list(self.gen_id_reconstructor.parameters())
        return not torch.is_grad_enabled() or not any(p.requires_grad for p in ps)

    def _forward_hoisted(self, in_xy, in_dxdy, pred_dxdy, seq_start_end, img, mask):
        """Same function as `forward` for frozen head weights (generator step, evaluation): the first layer of
        both heads is split by input block, its per-agent part ([in_enc | scene] and, for sample 0, soc) is
        evaluated once per agent and `mggan_disc_heads_*` adds the per-sample pred_enc part (6x fewer MACs than
        the k*N x 192 product; no (k*N, 192) classifier input is materialised)."""
        pred_len, n_samples, b, _ = pred_dxdy.shape
        N = in_xy.size(1)
        in_enc = self._in_enc(in_dxdy)                                  # (N, 32)
        pe = self._pred_enc(pred_dxdy)                                  # (k*b, 32), row = s*b + i
        pe0 = pe[:b]
        if mask is not None:
            pad = torch.zeros(N, pe.size(1), device=pe.device)
            pad[mask] = pe0
            pe0 = pad
        soc = self.social(in_xy, in_dxdy, torch.cat([in_enc, pe0], 1), seq_start_end)       # (N, 64): sample 0
        d0 = self.discs[0][0]
        mgan = self.gan_type == "mgan"
        w1 = torch.cat([d0.weight, self.gen_id_reconstructor[0].weight], 0) if mgan else d0.weight
        b1 = torch.cat([d0.bias, self.gen_id_reconstructor[0].bias], 0) if mgan else d0.bias
        hs, he = soc.shape[1], in_enc.shape[1]                           # column blocks: soc | in_enc | pred_enc | scene
        per_agent = in_enc
        w_agent = w1[:, hs:hs + he]
        if mask is not None:
            soc, per_agent = soc[mask], per_agent[mask]
        if img is not None:
            per_agent = torch.cat([per_agent, self._scene(img, mask)], 1)
            w_agent = torch.cat([w_agent, w1[:, hs + 2 * he:]], 1)
        if self.pool_type == "sgan":
            # `seq_start_end * n_samples` makes PoolHiddenNet emit the sample-0 pooling once per sample (it concatenates one
            # block per listed range, social_gan.py:227-228): the social term is per agent, not sample-0 only
            per_agent = torch.cat([soc, per_agent], 1)
            w_agent = torch.cat([w1[:, :hs], w_agent], 1)
            base = K.linear(per_agent, w_agent, b1)
            soc0 = torch.zeros_like(base)
        else:
            base = K.linear(per_agent, w_agent, b1)
            soc0 = K.linear(soc, w1[:, :hs])
        w1p = w1[:, hs + he:hs + 2 * he]
        d2 = self.discs[0][2]
        g2 = self.gen_id_reconstructor[2] if mgan else None
        p, branch = K.disc_heads(pe, base, soc0, w1p, d2.weight, d2.bias, g2.weight if mgan else None,
                                 g2.bias if mgan else None, b, n_samples)
        output = p.reshape(n_samples, b).t()
        if not mgan:
            return output
        return output, branch.reshape(n_samples, b, -1).transpose(0, 1)

    def forward(self, in_xy, in_dxdy, pred_xy, pred_dxdy, seq_start_end, return_all=False, img=None, mask=None):
        """pred_* (pred_len, k, n_act, 2) (3-D inputs are one sample).  Returns output (n_act, k) in
        (1e-7, 1 - 1e-7), and for gan_type='mgan' also branch logits (n_act, k, G)."""
        if pred_xy.dim() == 3:
            pred_xy, pred_dxdy = pred_xy.unsqueeze(1), pred_dxdy.unsqueeze(1)
        pred_len, n_samples, b, _ = pred_xy.shape
        N = in_xy.size(1)
        if self._heads_frozen() and not self.unbound_output:       # the fused heads kernel has the sigmoid built in
            return self._forward_hoisted(in_xy, in_dxdy, pred_dxdy, seq_start_end, img, mask)
        enc = self.encode(in_xy, in_dxdy, pred_xy, pred_dxdy, mask)
        soc0 = self.social(in_xy, in_dxdy, enc[:N], seq_start_end)
        if n_samples > 1 and self.pool_type == "sgan":
            soc = soc0.repeat(n_samples, 1)          # one pooled block per repeated range (social_gan.py:227-228)
        elif n_samples > 1:
            soc = torch.cat([soc0, enc.new_zeros((n_samples - 1) * N, soc0.shape[1])], 0)
        else:
            soc = soc0
        classifier_inp = torch.cat([soc, enc], dim=1)
        if mask is not None:
            classifier_inp = classifier_inp[mask.repeat(n_samples)]
        if img is not None:
            scene = self._scene(img, mask)
            classifier_inp = torch.cat([classifier_inp, scene.repeat(n_samples, 1)], 1)
        # sigmoid * (1 - 2 eps) + eps, or the raw score for gan_obj LS / W (reference discriminators.py:83,203)
        output = self._mlp2(self.discs[0], classifier_inp, K.ACT_NONE if self.unbound_output else K.ACT_SIGMOID_EPS)
        if not return_all:
            output = output.mean(1)
        output = output.reshape(n_samples, b).t()
        if self.gan_type == "gan":
            return output
        branch_out = self._mlp2(self.gen_id_reconstructor, classifier_inp)
        return output, branch_out.reshape(n_samples, b, -1).transpose(0, 1)
