"""MultiGenerator (reference: mggan/model/modules/standard.py:17-302).

Same constructor, attributes, state_dict layout (every decoder registered twice, `gs.{i}.*`
and `G_{i}.*`, standard.py:86-87) and forward contract.  What changes is the execution: the
encoder, scene and social attention, the PM-Network MLP, the sampling of generator indices and
the decoding of exactly the selected (agent, sample, generator) sequences are sm_100a kernels;
the reference's `forward_all` over G x M x N sequences followed by a gather becomes one decoder
launch over k x N sequences grouped by generator.
"""
import contextlib

import torch
import torch.nn as nn

from mggan import kernels as K
from mggan.model.modules.cnn import AttentionGlobal
from mggan.model.modules.common_modules import GeneratorOutput, RelativeDecoder, TrajectoryEncoder, get_input
from mggan.model.modules.social import SocialAttention
from mggan.model.modules.social_gan import PoolHiddenNet
from mggan.utils import get_global_noise, make_mlp


class MultiGenerator(nn.Module):
    def __init__(self, z_size, encoder_h_dim, decoder_h_dim, social_feat_size, num_gens, pred_len, embedding_dim,
                 inp_format, num_social_modules, pool_type, scene_dim, use_pinet, learn_prior=False):
        super().__init__()
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
        else:                                        # reference standard.py:66-71
            self.social = PoolHiddenNet(embedding_dim=embedding_dim, h_dim=encoder_h_dim, mlp_dim=social_feat_size,
                                        bottleneck_dim=encoder_h_dim)
        self.gs = nn.ModuleList()
        for i in range(num_gens):
            decoder = RelativeDecoder(pred_len=pred_len, embedding_dim=embedding_dim, h_dim=decoder_h_dim,
                                      num_layers=1, social_feat_size=social_feat_size, z_size=z_size, dropout=0.0,
                                      inp_format=inp_format)
            setattr(self, "G_{}".format(i), decoder)
            self.gs.append(decoder)
        self.n_gs = len(self.gs)
        self.pred_len = pred_len
        self.enc_h_to_dec_h = make_mlp([encoder_h_dim + z_size + scene_dim + social_feat_size, decoder_h_dim],
                                       batch_norm=False)
        assert not (use_pinet and learn_prior), "Using conditional distribution already, `learn_prior` has no effect"
        self.net_chooser = nn.Sequential(
            nn.Linear(encoder_h_dim + scene_dim + social_feat_size, encoder_h_dim // 2), nn.ReLU(),
            nn.Linear(encoder_h_dim // 2, encoder_h_dim // 2), nn.ReLU(),
            nn.Linear(encoder_h_dim // 2, num_gens))
        self.net_prior = nn.Parameter(torch.zeros(1, self.n_gs), requires_grad=learn_prior)
        self._sample_calls = 0
        self._shared = None             # dict while the trainer guarantees constant weights (see share_trunk)

    # ------------------------------------------------------------------ trunk sharing
    def share_trunk(self):
        """Start sharing the observation trunk (trajectory encoder, scene attention, social attention) between
        forwards on the SAME input tensors.  The discriminator step's no-grad forward (train.py:160) and the
        generator step's forward (train.py:46) run the same weights on the same observations; between
        `share_trunk()` and `drop_shared()` the trunk is evaluated once, with its autograd graph, and reused
        (BatchNorm running statistics are still updated once per forward).  The caller must call
        `drop_shared()` before the weights change."""
        self._shared = {}
        if self.scene_dim > 0:
            self.scene_encoder.memo = {}

    def drop_shared(self):
        self._shared = None
        if self.scene_dim > 0:
            self.scene_encoder.memo = None

    def _trunk(self, in_xy, in_dxdy, sub_batches, img):
        """-> (enc_h (N, 32 [+64] + 32), social_feats (N, 32))"""
        if self._shared is None:
            return self._trunk_compute(in_xy, in_dxdy, sub_batches, img)
        scenes = K.SceneIndex.get(sub_batches, in_xy.device)
        key = (id(in_xy), in_xy._version, id(in_dxdy), in_dxdy._version, id(scenes), self.training)
        with torch.enable_grad():           # the graph is built even under the D step's no_grad: the G step reuses it
            hit = self._shared.get(key)
            if hit is None:
                enc_h = self.encoder(get_input(in_xy, in_dxdy, self.inp_format))
                hit = self._shared[key] = (enc_h, self.social(in_xy, in_dxdy, enc_h, scenes))
            enc_h, social_feats = hit
            feats = [enc_h]
            if img is not None:
                feats.append(self.scene_encoder(img))       # memoised by kernels.scene_attention (replays BN updates)
            feats.append(social_feats)
            out = torch.cat(feats, -1)
        if not torch.is_grad_enabled():
            out, social_feats = out.detach(), social_feats.detach()
        return out, social_feats

    def _trunk_compute(self, in_xy, in_dxdy, sub_batches, img):
        enc_h = self.encoder(get_input(in_xy, in_dxdy, self.inp_format))
        enc_features = [enc_h]
        if img is not None:
            enc_features.append(self.scene_encoder(img))
        social_feats = self.social(in_xy, in_dxdy, enc_h, sub_batches)
        enc_features.append(social_feats)
        return torch.cat(enc_features, -1), social_feats

    # ------------------------------------------------------------------ pieces
    def _stacked_decoder_weights(self):
        """Kernel-layout decoder weights of all generators, stacked along dim 0.  The spatial embedding is folded into the
        input projection with two batched products over the generator axis (instead of a product per generator).  While
        the trainer shares the trunk (constant weights between the discriminator step's and the generator step's
        forward) the folded set is built once, with its autograd graph, and reused."""
        if self._shared is not None and "dec_w" in self._shared:
            w = self._shared["dec_w"]
            return w if torch.is_grad_enabled() else {k: v.detach() for k, v in w.items()}
        with torch.enable_grad() if self._shared is not None else contextlib.nullcontext():
            st = lambda f: torch.stack([f(g) for g in self.gs])
            H = self.decoder_h_dim
            w_ih, w_s = st(lambda g: g.decoder.weight_ih_l0), st(lambda g: g.spatial_embedding.weight)
            b_s = st(lambda g: g.spatial_embedding.bias)
            w1 = st(lambda g: g.hidden2pos[0].weight)
            w = {
                "wx": torch.bmm(w_ih, w_s),
                "b": torch.bmm(w_ih, b_s[:, :, None])[:, :, 0] + st(lambda g: g.decoder.bias_ih_l0) + st(lambda g: g.decoder.bias_hh_l0),
                "whh": st(lambda g: g.decoder.weight_hh_l0), "w1h": w1[:, :, :H], "w1s": w1[:, :, H:],
                "b1": st(lambda g: g.hidden2pos[0].bias), "w2": st(lambda g: g.hidden2pos[2].weight),
                "b2": st(lambda g: g.hidden2pos[2].bias),
            }
        if self._shared is not None:
            self._shared["dec_w"] = w
            if not torch.is_grad_enabled():
                return {k: v.detach() for k, v in w.items()}
        return w

    def pm_logits(self, enc_h):
        if not self.use_pinet:
            return self.net_prior.expand(enc_h.size(0), -1)
        nc = self.net_chooser
        x = K.mlp2(enc_h, nc[0].weight, nc[0].bias, nc[2].weight, nc[2].bias, K.ACT_RELU, 0.0, K.ACT_RELU)
        return K.linear(x, nc[4].weight, nc[4].bias)

    def get_samples(self, enc_h, num_samples=5):
        """(logits (n, G), generator indices (n, num_samples) int64).  The reference draws with
        Categorical(logits).sample (standard.py:217-225); here a device Gumbel-max sampler keyed on
        torch's seed and a call counter draws from the same distribution."""
        logits = self.pm_logits(enc_h)
        graph = getattr(self, "_graph", None)          # captured iteration: the Philox offset advances on the device
        seed = torch.initial_seed() & ((1 << 63) - 1)
        if graph is not None:
            # replayed draws live in their own offset space (bit 62 set by the graph's device-side offset; the static part is
            # the call's index inside the iteration), so they never meet the eager draws' `_sample_calls << 20`
            idx = K.gumbel_sample(logits, num_samples, seed, graph.sampler_calls << 20, graph.sampler_offset())
        else:
            self._sample_calls += 1
            idx = K.gumbel_sample(logits, num_samples, seed, self._sample_calls << 20, None)
        return logits, idx

    def _decode(self, in_xy, in_dxdy, enc_h, noise, social_feats, sel):
        """enc_h (n, C): per-agent encoding; noise (k', n, Z).  Decodes the sequences of `sel`."""
        w = self.enc_h_to_dec_h[0]
        C = enc_h.shape[1]
        A = K.linear(enc_h, w.weight[:, :C], w.bias)               # per-agent half of h0, hoisted
        wz = w.weight[:, C:]
        return K.decode(A, social_feats, in_xy[-1], in_dxdy[-1], noise, wz, self._stacked_decoder_weights(), sel,
                        self.pred_len)

    # ------------------------------------------------------------------ reference API
    def forward(self, in_xy, in_dxdy, sub_batches, noise=None, all_gen_out=True, img=None, num_samples=5, mask=None,
                gen_idxs=None):
        """See reference standard.py:111-215.  Returns (GeneratorOutput(rel, abs), logits, idx):
        rel/abs (pred_len, k, n_act, 2), or (pred_len, k, G, n_act, 2) under no_grad if all_gen_out.
        `gen_idxs` (additive): generator indices (n_act, k) or a callable logits -> indices used instead of a
        PM-Network draw when all_gen_out is False -- the prediction strategies of train.py:291-465 decode exactly
        the sequences they keep instead of all G * k * n and gathering."""
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
        noise = noise.contiguous()

        if all_gen_out:
            with torch.no_grad():
                pred_xy, pred_dxdy = self.forward_all(in_xy, in_dxdy, enc_h, noise=noise, social_feats=social_feats)
            net_chooser_out, sampled_gen_idxs = self.get_samples(enc_h, num_samples)
        else:
            with torch.no_grad():
                if gen_idxs is None:
                    net_chooser_out, sampled_gen_idxs = self.get_samples(enc_h, num_samples)
                else:
                    net_chooser_out = self.pm_logits(enc_h)
                    sampled_gen_idxs = gen_idxs(net_chooser_out) if callable(gen_idxs) else gen_idxs
                    sampled_gen_idxs = sampled_gen_idxs.to(device=enc_h.device, dtype=torch.int64)
                    assert sampled_gen_idxs.shape == (batch_size, num_samples), sampled_gen_idxs.shape
                    # caller-supplied indices (prediction strategies, tests): the selection kernel would decode an
                    # out-of-range index with generator 0; this path may synchronise, so check on the host
                    if sampled_gen_idxs.numel() and not torch.cuda.is_current_stream_capturing():
                        lo, hi = int(sampled_gen_idxs.min()), int(sampled_gen_idxs.max())
                        if lo < 0 or hi >= self.n_gs:
                            raise ValueError(f"gen_idxs out of range [0, {self.n_gs}): min {lo}, max {hi}")
            sel = K.Selection.from_indices(sampled_gen_idxs, self.n_gs)
            self.last_selection = sel           # per-generator draw counts for the trainer's reweighting
            pred_xy, pred_dxdy = self._decode(in_xy, in_dxdy, enc_h, noise, social_feats, sel)
            pred_xy = pred_xy.view(self.pred_len, num_samples, batch_size, 2)
            pred_dxdy = pred_dxdy.view(self.pred_len, num_samples, batch_size, 2)
        return GeneratorOutput(pred_dxdy, pred_xy), net_chooser_out, sampled_gen_idxs

    def forward_all(self, in_xy, in_dxdy, enc_h, noise, social_feats):
        """Every generator on every (sample, agent): two tensors (pred_len, k, G, n, 2) (abs, rel)
        (reference standard.py:227-265)."""
        k, n, _ = noise.shape
        sel = K.Selection.all_generators(n, k, self.n_gs, enc_h.device)
        pred_xy, pred_dxdy = self._decode(in_xy, in_dxdy, enc_h, noise.contiguous(), social_feats, sel)
        shape = (self.pred_len, k, self.n_gs, n, 2)
        return pred_xy.view(shape), pred_dxdy.view(shape)
