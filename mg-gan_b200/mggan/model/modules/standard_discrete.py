"""DiscreteLatentGenerator (`--experiment discrete`; reference: mggan/model/modules/standard_discrete.py:18-257).

The ablation with ONE decoder: the "generator" index is a discrete latent code -- one-hot -> `one_hot_sample_encoder`
(Linear, ReLU, Linear) -- concatenated to the encoding in front of the noise, h0 = enc_h_to_dec_h([enc | code(g) | z]).
Same constructor, parameter names and forward contract as the reference.  No new kernel: the trunk (encoder, scene and
social attention), the PM-Network and the sampler are MultiGenerator's, the h0 projection runs in `mggan_linear_*` and
the k x n (or, for the PM step, k x G x n) sequences are decoded by one launch of the decoder kernels with the single
decoder's weights (`RelativeDecoder.forward`).
"""
import torch
import torch.nn as nn

from mggan import kernels as K
from mggan.model.modules.cnn import AttentionGlobal
from mggan.model.modules.common_modules import GeneratorOutput, RelativeDecoder, TrajectoryEncoder
from mggan.model.modules.social import SocialAttention
from mggan.model.modules.social_gan import PoolHiddenNet
from mggan.model.modules.standard import MultiGenerator
from mggan.utils import get_global_noise, make_mlp


class _Counts:
    """What the trainer reads from `G.last_selection`: per-generator draw counts (train.py:92-96)."""

    def __init__(self, idx, num_gens):
        # elementwise compare + sum: no host synchronisation (torch.bincount has one), so it can be graph-captured
        gens = torch.arange(num_gens, device=idx.device)
        self.totals = (idx.reshape(-1, 1) == gens).sum(0).to(torch.int32)


class DiscreteLatentGenerator(MultiGenerator):
    # MultiGenerator supplies the trunk (`_trunk`, `share_trunk`, `drop_shared`), `pm_logits` and `get_samples`;
    # construction, decoding and `forward` are this class's own.
    def __init__(self, z_size, encoder_h_dim, decoder_h_dim, social_feat_size, num_gens, pred_len, embedding_dim,
                 inp_format, num_social_modules, pool_type, scene_dim, use_pinet, learn_prior=False):
        nn.Module.__init__(self)
        assert inp_format in ("rel", "abs", "abs_rel")
        assert num_social_modules in (0, 1, num_gens)
        assert pool_type in ("sways", "sgan")
        if inp_format != "rel" or social_feat_size <= 0 or num_social_modules != 1:
            raise NotImplementedError("B200 path covers the default configuration: inp_format='rel', one social module")
        if encoder_h_dim != 32 or decoder_h_dim != 32 or social_feat_size != 32:
            raise NotImplementedError("B200 path: h_dim = decoder_h_dim = 32 (config.py defaults)")
        if scene_dim not in (0, 64):
            raise NotImplementedError("scene_dim must be 0 (no scene encoder) or 64")
        self.use_pinet, self.inp_format, self.z_size = use_pinet, inp_format, z_size
        self.embedding_dim, self.social_feat_size = embedding_dim, social_feat_size
        self.n_social_modules, self.pool_type = num_social_modules, pool_type
        self.decoder_h_dim, self.encoder_h_dim, self.scene_dim = decoder_h_dim, encoder_h_dim, scene_dim

        self.encoder = TrajectoryEncoder(inp_size=2, hidden_size=encoder_h_dim, embedding_dim=embedding_dim,
                                         num_layers=1)
        if scene_dim > 0:
            self.scene_encoder = AttentionGlobal(noise_attention_dim=0, PhysFeature=True, num_layers=2,
                                                 channels_cnn=16)
        if pool_type == "sways":
            self.social = SocialAttention(social_feat_size, encoder_h_dim)
        else:
            self.social = PoolHiddenNet(embedding_dim=embedding_dim, h_dim=encoder_h_dim, mlp_dim=social_feat_size,
                                        bottleneck_dim=encoder_h_dim)
        self.decoder = RelativeDecoder(pred_len=pred_len, embedding_dim=embedding_dim, h_dim=decoder_h_dim, num_layers=1,
                                       social_feat_size=encoder_h_dim, z_size=z_size, dropout=0.0, inp_format=inp_format)
        self.n_gs = num_gens
        self.pred_len = pred_len
        self.enc_h_to_dec_h = make_mlp([encoder_h_dim + z_size + scene_dim + z_size + social_feat_size, decoder_h_dim],
                                       batch_norm=False)
        assert not (use_pinet and learn_prior), "Using conditional distribution already, `learn_prior` has no effect"
        self.net_chooser = nn.Sequential(
            nn.Linear(encoder_h_dim + scene_dim + social_feat_size, encoder_h_dim // 2), nn.ReLU(),
            nn.Linear(encoder_h_dim // 2, encoder_h_dim // 2), nn.ReLU(),
            nn.Linear(encoder_h_dim // 2, num_gens))
        self.one_hot_sample_encoder = make_mlp([num_gens, z_size, z_size])
        self.net_prior = nn.Parameter(torch.zeros(1, self.n_gs), requires_grad=learn_prior)
        self._sample_calls = 0
        self._shared = None

    # ------------------------------------------------------------------ pieces
    def _codes(self, device):
        """(G, z): the latent code of every generator index (one_hot_sample_encoder on the identity)."""
        enc = self.one_hot_sample_encoder
        eye = torch.eye(self.n_gs, device=device)
        return K.linear(K.linear(eye, enc[0].weight, enc[0].bias, K.ACT_RELU), enc[2].weight, enc[2].bias)

    def _decode_rows(self, in_xy, in_dxdy, social_feats, enc_h, code, noise, reps):
        """enc_h (n, C), social (n, 32) repeated `reps` times to match code / noise (reps * n, z) -> (T, reps * n, 2) x 2."""
        w = self.enc_h_to_dec_h[0]
        h0 = K.linear(torch.cat([enc_h.repeat(reps, 1), code, noise], 1), w.weight, w.bias)
        return self.decoder(in_xy[-1].repeat(reps, 1), in_dxdy[-1].repeat(reps, 1), None, social_feats.repeat(reps, 1),
                            (h0[None], None))

    # ------------------------------------------------------------------ reference API
    def forward(self, in_xy, in_dxdy, sub_batches, noise=None, all_gen_out=True, img=None, num_samples=5, mask=None,
                gen_idxs=None):
        """See reference standard_discrete.py:109-237.  Returns (GeneratorOutput(rel, abs), logits, idx) with
        rel / abs (pred_len, k, n_act, 2), or (pred_len, k, G, n_act, 2) under no_grad if all_gen_out.  `gen_idxs`
        (additive) as in MultiGenerator.forward."""
        batch_size = in_xy.size(1)
        enc_h, social_feats = self._trunk(in_xy, in_dxdy, sub_batches, img)
        if noise is not None:
            assert noise.shape == (num_samples, batch_size, self.z_size)
        else:
            noise = get_global_noise(self.z_size, sub_batches, "gaussian", in_xy.device, num_samples)
        if mask is not None:
            in_xy, in_dxdy = in_xy[:, mask], in_dxdy[:, mask]
            enc_h, social_feats, noise = enc_h[mask], social_feats[mask], noise[:, mask]
            batch_size = enc_h.shape[0]
        k, n, G = num_samples, batch_size, self.n_gs

        if all_gen_out:
            with torch.no_grad():
                # row = (sample s, generator g, agent i): the layout of the reference's nested stacks (:170-195)
                code = self._codes(enc_h.device)[None, :, None, :].expand(k, G, n, self.z_size).reshape(k * G * n, -1)
                z = noise[:, None].expand(k, G, n, self.z_size).reshape(k * G * n, -1)
                pred_xy, pred_dxdy = self._decode_rows(in_xy, in_dxdy, social_feats, enc_h, code, z, k * G)
            net_chooser_out, sampled_gen_idxs = self.get_samples(enc_h, num_samples)
            shape = (self.pred_len, k, G, n, 2)
            return GeneratorOutput(pred_dxdy.view(shape), pred_xy.view(shape)), net_chooser_out, sampled_gen_idxs

        with torch.no_grad():
            if gen_idxs is None:
                net_chooser_out, sampled_gen_idxs = self.get_samples(enc_h, num_samples)
            else:
                net_chooser_out = self.pm_logits(enc_h)
                sampled_gen_idxs = gen_idxs(net_chooser_out) if callable(gen_idxs) else gen_idxs
                sampled_gen_idxs = sampled_gen_idxs.to(device=enc_h.device, dtype=torch.int64)
                assert sampled_gen_idxs.shape == (n, k), sampled_gen_idxs.shape
        self.last_selection = _Counts(sampled_gen_idxs, G)
        # row = (sample s, agent i); the code is trainable: gradients reach one_hot_sample_encoder through the gather
        code = self._codes(enc_h.device).index_select(0, sampled_gen_idxs.t().reshape(-1))
        pred_xy, pred_dxdy = self._decode_rows(in_xy, in_dxdy, social_feats, enc_h, code, noise.reshape(k * n, -1), k)
        shape = (self.pred_len, k, n, 2)
        return GeneratorOutput(pred_dxdy.view(shape), pred_xy.view(shape)), net_chooser_out, sampled_gen_idxs

    def forward_all(self, in_xy, in_dxdy, enc_h, noise, social_feats):
        """One decoding of every row with the given encoding (reference :239-257; enc_h already carries the code)."""
        w = self.enc_h_to_dec_h[0]
        h0 = K.linear(torch.cat([enc_h, noise], -1), w.weight, w.bias)
        return self.decoder(in_xy[-1], in_dxdy[-1], noise, social_feats, (h0[None], None))
