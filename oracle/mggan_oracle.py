"""TEST INFRASTRUCTURE — CPU oracle for the MG-GAN training-step hot path.

This file is a *restatement* (plain PyTorch fp32 on CPU, functional style, one
function per reference operator) of the algorithm the reference implements in
`mggan/model/modules/*.py`, `mggan/model/train.py` and `mggan/abstract_train.py`
(selflein/MG-GAN @ 5ea1167).  It is the checker for the CUDA path: only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  Nothing under `mg-gan_b200/` does, and the product
fails loudly when its CUDA library is missing instead of falling back here.

Parity pin: the reference ships no tests or golden vectors, so
this oracle is pinned against outputs of the reference itself, executed in the
build container by `oracle/make_golden.py` and frozen under `tests/golden/`
(`tests/test_oracle_golden.py`), plus a live cross-check when `/root/reference`
is mounted (`tests/test_oracle_vs_reference.py`).

All tensors are fp32 (indices int64, masks bool); parameters are held in plain
dicts keyed by the reference's `state_dict` names (decoders under `gs.{g}.`; the
reference registers every decoder a second time as `G_{g}.`, standard.py:86-87).
The arithmetic itself (LSTM cell, Linear, conv, BatchNorm, AdamW) lives in
PyTorch in the reference too (pin: pytorch=1.6.0, environment.yml:102; executed
here with torch 2.11 — see SURVEY.md §7 "Torch-version semantics").
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

PRED_LEN = 12          # model_factory.py:18
BN_EPS = 1e-5          # nn.BatchNorm2d default, cnn.py:140-141
BN_MOMENTUM = 0.1
D_EPS = 1e-7           # discriminators.py:110


# --------------------------------------------------------------------------- helpers
def _lin(sd, pre, x):
    return x @ sd[pre + ".weight"].t() + sd[pre + ".bias"]


def lrelu(x, a):
    return torch.where(x > 0, x, x * a)


def lstm_step(sd, pre, x, h, c):
    """One nn.LSTM layer step, gate rows ordered [i; f; g; o]."""
    gates = (x @ sd[pre + ".weight_ih_l0"].t() + sd[pre + ".bias_ih_l0"]
             + h @ sd[pre + ".weight_hh_l0"].t() + sd[pre + ".bias_hh_l0"])
    H = h.shape[1]
    i = torch.sigmoid(gates[:, :H])
    f = torch.sigmoid(gates[:, H:2 * H])
    g = torch.tanh(gates[:, 2 * H:3 * H])
    o = torch.sigmoid(gates[:, 3 * H:])
    c = f * c + i * g
    h = o * torch.tanh(c)
    return h, c


def trajectory_encoder(sd, pre, inp):
    """TrajectoryEncoder.forward (common_modules.py:48-66): Linear embed, 1-layer LSTM, h_T."""
    T, N, _ = inp.shape
    H = sd[pre + ".encoder.weight_hh_l0"].shape[1]
    h = inp.new_zeros(N, H)
    c = inp.new_zeros(N, H)
    for t in range(T):
        x = _lin(sd, pre + ".embedding", inp[t])
        h, c = lstm_step(sd, pre + ".encoder", x, h, c)
    return h


def relative_decoder(sd, pre, xy, dxdy, social, h0, pred_len=PRED_LEN):
    """RelativeDecoder.forward, inp_format="rel" (common_modules.py:97-131).

    Returns (abs (T,R,2), rel (T,R,2)).  The `noise` argument of the reference
    is unused inside the decoder.
    """
    h, c = h0, torch.zeros_like(h0)
    out_abs, out_rel = [], []
    for _ in range(pred_len):
        x = _lin(sd, pre + ".spatial_embedding", dxdy)
        h, c = lstm_step(sd, pre + ".decoder", x, h, c)
        u = lrelu(_lin(sd, pre + ".hidden2pos.0", torch.cat([h, social], 1)), 0.01)
        dxdy = _lin(sd, pre + ".hidden2pos.2", u)
        xy = xy + dxdy
        out_abs.append(xy)
        out_rel.append(dxdy)
    return torch.stack(out_abs), torch.stack(out_rel)


# --------------------------------------------------------------------------- social attention
def social_features(x4):
    """SocialFeatures / BearingMTX / DCA_MTX (social.py:67-104) on rows x4 = (p, v).

    out[i, j] = (|dp|, cos-bearing, distance of closest approach), dp = p_i - p_j.
    """
    D = x4[:, None, :] - x4[None, :, :]
    dp, dv = D[..., :2], D[..., 2:]
    dist = dp.norm(dim=2)
    v = x4[:, None, 2:].expand(-1, x4.shape[0], -1)
    bearing = (dp * v).sum(-1) / (dist * v.norm(dim=2) + 1e-6)
    ttca = -(dp * dv).sum(-1) / ((dv * dv).sum(-1) + 1e-6)
    dca = (dp + ttca[..., None] * dv).norm(dim=2)
    return torch.stack([dist, bearing, dca], 2)


def _embed_social(sd, pre, f):
    f = torch.relu(_lin(sd, pre + ".feature_embedder.fc.0", f))
    f = torch.relu(_lin(sd, pre + ".feature_embedder.fc.2", f))
    return _lin(sd, pre + ".feature_embedder.fc.4", f)


def social_attention(sd, pre, xy_last, dxdy_last, h, sub_batches, mode="scene"):
    """SocialAttention.forward (social.py:107-123) + AttentionPooling (social.py:14-30).

    mode="reference": the reference's cost structure — features and the 3-layer
      embedding on ALL rows^2 pairs, then a python loop per agent over its scene.
    mode="scene": identical outputs, only in-scene pairs are evaluated.
    `sub_batches` may list the same range several times (the discriminator passes
    `seq_start_end * n_samples`, discriminators.py:179-184); rows outside every
    range and single-agent scenes stay 0 (social.py:19-20).
    """
    x4 = torch.cat([xy_last, dxdy_last], -1)
    Wh = _lin(sd, pre + ".attention.W", h)
    S = torch.zeros_like(h)
    if mode == "reference":
        emb = _embed_social(sd, pre, social_features(x4))
        for a, b in sub_batches:
            n = b - a
            if n == 1:
                continue
            for ii in range(a, b):
                sigma = (emb[ii, a:b] * Wh[a:b]).sum(1).clone()
                sigma[ii - a] = -1000
                att = torch.softmax(sigma, 0)
                S[ii] = att @ h[a:b]
        return S
    rows = []
    for a, b in dict.fromkeys((int(a), int(b)) for a, b in sub_batches):
        n = b - a
        if n == 1:
            continue
        emb = _embed_social(sd, pre, social_features(x4[a:b]))          # (n, n, F)
        sigma = (emb * Wh[a:b][None]).sum(-1)
        sigma = sigma.masked_fill(torch.eye(n, dtype=torch.bool), -1000.0)
        rows.append((a, b, torch.softmax(sigma, 1) @ h[a:b]))
    if rows:
        pieces, cur = [], 0
        for a, b, val in sorted(rows, key=lambda r: r[0]):
            if a > cur:
                pieces.append(h.new_zeros(a - cur, h.shape[1]))
            pieces.append(val)
            cur = b
        if cur < h.shape[0]:
            pieces.append(h.new_zeros(h.shape[0] - cur, h.shape[1]))
        S = torch.cat(pieces, 0)
    return S


def pool_hidden_net(sd, pre, xy_last, h, sub_batches):
    """PoolHiddenNet.forward (social_gan.py:203-229), `--pool_type sgan`: per scene, for every ordered pair (a, b)
    MLP([Linear_2->E(pos_b - pos_a), h_b]) with mlp_pre_pool = Linear, ReLU, Linear (make_mlp of 3 dims, utils.py:134-149),
    then max over b (b = a included).  Output rows follow the order of `sub_batches` (concatenation, :227-228): the
    discriminator's `seq_start_end * n_samples` therefore yields n_samples copies of the sample-0 pooling."""
    out = []
    for a, b in sub_batches:
        n = b - a
        hid = h[a:b].repeat(n, 1)                                      # H1, H2, ..., H1, H2, ...   (row a' * n + b')
        p1 = xy_last[a:b].repeat(n, 1)                                 # pos of b'
        p2 = xy_last[a:b].unsqueeze(1).repeat(1, n, 1).view(-1, 2)     # pos of a'
        emb = _lin(sd, pre + ".spatial_embedding", p1 - p2)
        x = torch.cat([emb, hid], 1)
        y = _lin(sd, pre + ".mlp_pre_pool.2", torch.relu(_lin(sd, pre + ".mlp_pre_pool.0", x)))
        out.append(y.view(n, n, -1).max(1)[0])
    return torch.cat(out, 0)


# --------------------------------------------------------------------------- physical attention
def _bn_train(sd, pre, x, training):
    w, b = sd[pre + ".weight"], sd[pre + ".bias"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        n = x.numel() / x.shape[1]
        with torch.no_grad():
            sd[pre + ".running_mean"].mul_(1 - BN_MOMENTUM).add_(mean.detach() * BN_MOMENTUM)
            sd[pre + ".running_var"].mul_(1 - BN_MOMENTUM).add_(var.detach() * (n / (n - 1)) * BN_MOMENTUM)
            sd[pre + ".num_batches_tracked"].add_(1)
    else:
        mean, var = sd[pre + ".running_mean"], sd[pre + ".running_var"]
    xh = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + BN_EPS)
    return xh * w[None, :, None, None] + b[None, :, None, None]


def attention_global(sd, pre, img, training=True):
    """AttentionGlobal.forward (cnn.py:109-116) over CNN/Conv_Blocks (cnn.py:137-158,205-241).

    2x [conv3x3 pad1 -> BatchNorm2d -> ReLU -> MaxPool2x2], (N,4,33,33) -> (N,C,8,8);
    per position MLP C->32->C (LeakyReLU 0.01), softmax over channels, weighted sum -> (N,64).
    """
    x = img
    for blk in (1, 2):
        p = f"{pre}.CNN.encoder.ConvBlock_{blk}.Block"
        x = F.conv2d(x, sd[p + ".Conv_1.weight"], sd[p + ".Conv_1.bias"], stride=1, padding=1)
        x = _bn_train(sd, p + ".BN_1", x, training)
        x = F.max_pool2d(torch.relu(x), 2, 2)
    v = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1, x.shape[1])        # (N, 64, C)
    s = _lin(sd, pre + ".cnn_attention.2", lrelu(_lin(sd, pre + ".cnn_attention.0", v), 0.01))
    return (torch.softmax(s, 2) * v).sum(-1)


# --------------------------------------------------------------------------- PM-Net / selection
def pm_logits(sd, enc_cat, use_pinet=True):
    """MultiGenerator.get_samples, logits part (standard.py:217-222)."""
    if not use_pinet:
        return sd["net_prior"].expand(enc_cat.shape[0], -1)
    x = torch.relu(_lin(sd, "net_chooser.0", enc_cat))
    x = torch.relu(_lin(sd, "net_chooser.2", x))
    return _lin(sd, "net_chooser.4", x)


def selection_indices(idx):
    """get_selection_indices (utils.py:234-248): occurrence rank of every draw in its row."""
    same = idx[:, :, None] == idx[:, None, :]                # [i, j, j']
    earlier = torch.tril(torch.ones(idx.shape[1], idx.shape[1], dtype=torch.bool), -1)
    return (same & earlier[None]).sum(-1).to(idx.dtype)


def global_noise(z_size, sub_batches, generator=None):
    """get_global_noise (utils.py:160-165): one N(0,1) vector per scene, shared by its agents."""
    out = []
    for a, b in sub_batches:
        out.append(torch.randn(1, z_size, generator=generator).repeat(b - a, 1))
    return torch.cat(out)


# --------------------------------------------------------------------------- generator
def forward_all(sd, n_gens, xy_last, dxdy_last, enc_cat, noise, social):
    """MultiGenerator.forward_all (standard.py:227-265) -> (T, M, G, N, 2) x2 (abs, rel)."""
    M, N, _ = noise.shape
    z = noise.reshape(M * N, -1)
    dec_in = torch.cat([enc_cat.repeat(M, 1), z], -1)
    h0 = _lin(sd, "enc_h_to_dec_h.0", dec_in)
    xy = xy_last.repeat(M, 1)
    dxdy = dxdy_last.repeat(M, 1)
    soc = social.repeat(M, 1)
    A, R = [], []
    for g in range(n_gens):
        a, r = relative_decoder(sd, f"gs.{g}", xy, dxdy, soc, h0)
        A.append(a.reshape(PRED_LEN, M, N, 2))
        R.append(r.reshape(PRED_LEN, M, N, 2))
    return torch.stack(A, 2), torch.stack(R, 2)


def generator_trunk(sd, in_xy, in_dxdy, sub_batches, img, training, social_mode="scene"):
    """Encoder + scene attention + social attention (standard.py:143-155)."""
    enc_h = trajectory_encoder(sd, "encoder", in_dxdy)
    feats = [enc_h]
    if img is not None:
        feats.append(attention_global(sd, "scene_encoder", img, training))
    if "social.mlp_pre_pool.0.weight" in sd:                               # --pool_type sgan (standard.py:63-71)
        social = pool_hidden_net(sd, "social", in_xy[-1], enc_h, sub_batches)
    else:
        social = social_attention(sd, "social", in_xy[-1], in_dxdy[-1], enc_h, sub_batches, social_mode)
    feats.append(social)
    return torch.cat(feats, -1), social


def discrete_generator_forward(sd, n_gens, in_xy, in_dxdy, sub_batches, noise, all_gen_out, img, num_samples,
                               mask=None, gen_idxs=None, training=True, use_pinet=True, social_mode="scene",
                               generator=None):
    """DiscreteLatentGenerator.forward (`--experiment discrete`, standard_discrete.py:109-257): ONE decoder; the generator
    index is a discrete latent code, one_hot -> Linear, ReLU, Linear (`one_hot_sample_encoder`), concatenated to the
    encoding in front of the noise: h0 = enc_h_to_dec_h([enc | code(g) | z]) (:150-153,201-212,239-257)."""
    N = in_xy.shape[1]
    enc_cat, social = generator_trunk(sd, in_xy, in_dxdy, sub_batches, img, training, social_mode)
    if noise is None:
        noise = torch.stack([global_noise(8, sub_batches, generator) for _ in range(num_samples)])
    assert noise.shape[:2] == (num_samples, N)
    if mask is not None:
        in_xy, in_dxdy = in_xy[:, mask], in_dxdy[:, mask]
        enc_cat, social, noise = enc_cat[mask], social[mask], noise[:, mask]
    n_act = enc_cat.shape[0]

    def code(one_hot):
        return _lin(sd, "one_hot_sample_encoder.2", torch.relu(_lin(sd, "one_hot_sample_encoder.0", one_hot)))

    def decode(inp_h, z):
        h0 = _lin(sd, "enc_h_to_dec_h.0", torch.cat([inp_h, z], -1))
        return relative_decoder(sd, "decoder", in_xy[-1], in_dxdy[-1], social, h0)

    def draw(logits):
        if gen_idxs is not None:
            return gen_idxs
        p = torch.softmax(logits.detach(), 1)
        return torch.multinomial(p, num_samples, replacement=True, generator=generator)

    if all_gen_out:
        A, R = [], []
        with torch.no_grad():
            for i in range(num_samples):
                ga, gr = [], []
                for j in range(n_gens):
                    oh = F.one_hot(torch.tensor(j), n_gens)[None].repeat(n_act, 1).float()
                    a, r = decode(torch.cat([enc_cat, code(oh)], 1), noise[i])
                    ga.append(a)
                    gr.append(r)
                A.append(torch.stack(ga, 1))
                R.append(torch.stack(gr, 1))
        logits = pm_logits(sd, enc_cat, use_pinet)
        return (torch.stack(R, 1), torch.stack(A, 1)), logits, draw(logits)
    with torch.no_grad():
        logits = pm_logits(sd, enc_cat, use_pinet)
        idx = draw(logits)
    oh = F.one_hot(idx, n_gens).float()
    A, R = [], []
    for i in range(num_samples):
        a, r = decode(torch.cat([enc_cat, code(oh[:, i])], 1), noise[i])
        A.append(a)
        R.append(r)
    return (torch.stack(R, 1), torch.stack(A, 1)), logits, idx


def generator_forward(sd, n_gens, in_xy, in_dxdy, sub_batches, noise, all_gen_out, img,
                      num_samples, mask=None, gen_idxs=None, training=True, use_pinet=True,
                      social_mode="scene", generator=None):
    """MultiGenerator.forward (standard.py:111-215).

    `gen_idxs` (N_act, k) int64 injects the PM-Net draws (the reference samples them with
    Categorical under no_grad, standard.py:187-188,223-224); None draws them here.
    Returns ((rel, abs), logits, gen_idxs).
    """
    if "one_hot_sample_encoder.0.weight" in sd:                          # --experiment discrete (model_factory.py:50-65)
        return discrete_generator_forward(sd, n_gens, in_xy, in_dxdy, sub_batches, noise, all_gen_out, img, num_samples,
                                          mask, gen_idxs, training, use_pinet, social_mode, generator)
    N = in_xy.shape[1]
    enc_cat, social = generator_trunk(sd, in_xy, in_dxdy, sub_batches, img, training, social_mode)
    if noise is None:
        noise = torch.stack([global_noise(8, sub_batches, generator) for _ in range(num_samples)])
    assert noise.shape[:2] == (num_samples, N)
    if mask is not None:
        in_xy, in_dxdy = in_xy[:, mask], in_dxdy[:, mask]
        enc_cat, social, noise = enc_cat[mask], social[mask], noise[:, mask]
    n_act = enc_cat.shape[0]

    def draw(logits):
        if gen_idxs is not None:
            return gen_idxs
        p = torch.softmax(logits.detach(), 1)
        return torch.multinomial(p, num_samples, replacement=True, generator=generator)

    if all_gen_out:
        with torch.no_grad():
            pabs, prel = forward_all(sd, n_gens, in_xy[-1], in_dxdy[-1], enc_cat, noise, social)
        logits = pm_logits(sd, enc_cat, use_pinet)
        idx = draw(logits)
        return (prel, pabs), logits, idx
    with torch.no_grad():
        logits = pm_logits(sd, enc_cat, use_pinet)
        idx = draw(logits)
    offs = selection_indices(idx)
    M = int(offs.max()) + 1
    pabs, prel = forward_all(sd, n_gens, in_xy[-1], in_dxdy[-1], enc_cat, noise[:M], social)
    pabs = pabs.reshape(PRED_LEN, M * n_gens, n_act, 2)
    prel = prel.reshape(PRED_LEN, M * n_gens, n_act, 2)
    sel = idx + offs * n_gens                                            # (n_act, k)
    ar = torch.arange(n_act)[:, None]
    pabs = pabs[:, sel, ar].transpose(1, 2)
    prel = prel[:, sel, ar].transpose(1, 2)
    return (prel, pabs), logits, idx


# --------------------------------------------------------------------------- discriminator
def discriminator_forward(sd, in_xy, in_dxdy, pred_xy, pred_dxdy, sub_batches, img=None,
                          mask=None, training=True, social_mode="scene", unbound_output=False):
    """MultiDiscriminatorTrajectory.forward, gan_type="mgan" (discriminators.py:113-219).

    Returns (output (N_act, k) in (1e-7, 1-1e-7), branch (N_act, k, G)); with unbound_output
    (gan_obj LS / W, model_factory.py:14, discriminators.py:83,203) the head has no sigmoid.
    """
    if pred_xy.dim() == 3:
        pred_xy, pred_dxdy = pred_xy[:, None], pred_dxdy[:, None]
    T, k, n_act, _ = pred_xy.shape
    N = in_xy.shape[1]
    in_enc = trajectory_encoder(sd, "in_encoder", in_dxdy)
    in_enc = _lin(sd, "in_encoder_fc.2", lrelu(_lin(sd, "in_encoder_fc.0", in_enc), 0.2))
    pv = pred_dxdy.permute(1, 2, 0, 3).reshape(k * n_act, T * 2)
    pred_enc = _lin(sd, "pred_encoder.2", lrelu(_lin(sd, "pred_encoder.0", pv), 0.2))
    if mask is not None:
        full = torch.zeros(k * N, pred_enc.shape[1])
        full[mask.repeat(k)] = pred_enc
        pred_enc = full
    enc = torch.cat([in_enc.repeat(k, 1), pred_enc], 1)                  # row = s*N + i
    if "social.mlp_pre_pool.0.weight" in sd:                               # --pool_type sgan (discriminators.py:59-69)
        soc = pool_hidden_net(sd, "social", in_xy[-1].repeat(k, 1), enc, list(sub_batches) * k)
    elif social_mode == "reference":
        soc = social_attention(sd, "social", in_xy[-1].repeat(k, 1), in_dxdy[-1].repeat(k, 1), enc,
                               list(sub_batches) * k, "reference")
    else:
        # `seq_start_end * k` repeats the SAME ranges: only sample-0 rows are ever written.
        soc0 = social_attention(sd, "social", in_xy[-1], in_dxdy[-1], enc[:N], sub_batches, "scene")
        soc = torch.cat([soc0, enc.new_zeros((k - 1) * N, enc.shape[1])], 0)
    c = torch.cat([soc, enc], 1)
    if mask is not None:
        c = c[mask.repeat(k)]
    if img is not None:
        if mask is not None:
            img = img[mask]
        c = torch.cat([c, attention_global(sd, "scene_encoder", img, training).repeat(k, 1)], 1)
    out = _lin(sd, "discs.0.2", lrelu(_lin(sd, "discs.0.0", c), 0.2))
    if not unbound_output:
        out = torch.sigmoid(out) * (1 - 2 * D_EPS) + D_EPS
    out = out.mean(1).reshape(k, n_act).t()
    if "gen_id_reconstructor.0.weight" not in sd:            # gan_type "gan": no generator-id head (:210-211)
        return out, None
    br = _lin(sd, "gen_id_reconstructor.2", lrelu(_lin(sd, "gen_id_reconstructor.0", c), 0.2))
    return out, br.reshape(k, n_act, -1).transpose(0, 1)


# --------------------------------------------------------------------------- optimiser
def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (L2) over the tensors that have a gradient."""
    gs = [g for g in grads.values() if g is not None]
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in gs]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in gs:
        g.mul_(coef)
    return total


class AdamW:
    """torch.optim.AdamW(lr, betas=(beta1, .999), eps=1e-8, weight_decay=0.01), per-tensor
    `step`; tensors whose grad is None are skipped (abstract_train.py:45-50, torch 2.x
    `zero_grad(set_to_none=True)` semantics)."""

    def __init__(self, names, lr=1e-3, beta1=0.5, beta2=0.999, eps=1e-8, wd=0.01):
        self.lr0, self.lr, self.b1, self.b2, self.eps, self.wd = lr, lr, beta1, beta2, eps, wd
        self.state = {n: None for n in names}

    def step(self, sd, grads):
        with torch.no_grad():
            for n in self.state:
                g = grads.get(n)
                if g is None:
                    continue
                if self.state[n] is None:
                    self.state[n] = {"step": 0, "m": torch.zeros_like(sd[n]), "v": torch.zeros_like(sd[n])}
                st = self.state[n]
                st["step"] += 1
                t = st["step"]
                sd[n].mul_(1 - self.lr * self.wd)
                st["m"].mul_(self.b1).add_(g, alpha=1 - self.b1)
                st["v"].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                bc1 = 1 - self.b1 ** t
                bc2 = 1 - self.b2 ** t
                denom = st["v"].sqrt() / math.sqrt(bc2) + self.eps
                sd[n].addcdiv_(st["m"], denom, value=-(self.lr / bc1))

    def cosine(self, epoch, t_max):
        """CosineAnnealingLR(T_max, eta_min=0) closed form after `epoch` scheduler steps."""
        self.lr = self.lr0 * (1 + math.cos(math.pi * epoch / t_max)) / 2


# --------------------------------------------------------------------------- trainer
def _bce(p, y):
    return -(y * torch.clamp(torch.log(p), min=-100) + (1 - y) * torch.clamp(torch.log(1 - p), min=-100))


def _is_param(name):
    return not (name.endswith("running_mean") or name.endswith("running_var")
                or name.endswith("num_batches_tracked"))


class OracleTrainer:
    """PiNetMultiGeneratorGAN.{discriminator,generator,net_chooser}_step (train.py:137-213,
    23-135, 578-658) for the default flags: gan_type=mgan, gan_obj=NS, weighting_target=ml,
    l2_loss_type=min_g_z, inp_format=rel.  All randomness is injected."""

    def __init__(self, sdG, sdD, n_gens, num_samples=20, sigma=1.0, l2_w=1.0, clf_w=1.0,
                 pi_w=1.0, clip_g=500, clip_d=100, lr=1e-3, beta1=0.5, use_pinet=True,
                 social_mode="scene", gan_obj="NS", weighting_target="ml", epoch=1, l2_loss_type="min_g_z"):
        self.l2_loss_type = l2_loss_type         # "mse": squared per-step distances (train.py:62-63)
        # non-default objectives (abstract_train.py:61-79) and PM-step targets (train.py:604-647)
        self.gan_obj, self.weighting_target, self.epoch = gan_obj, weighting_target, epoch
        self.G = {k: v.clone() for k, v in sdG.items() if not k.startswith("G_")}
        self.D = {k: v.clone() for k, v in sdD.items()}
        for sd in (self.G, self.D):
            for n, v in sd.items():
                if _is_param(n) and n != "net_prior":
                    v.requires_grad_(True)
        self.n_gens, self.k, self.sigma = n_gens, num_samples, sigma
        self.l2_w, self.clf_w, self.pi_w = l2_w, clf_w, pi_w
        self.clip_g, self.clip_d = clip_g, clip_d
        self.use_pinet, self.social_mode = use_pinet, social_mode
        self.optG = AdamW([n for n in self.G if _is_param(n)], lr, beta1)
        self.optD = AdamW([n for n in self.D if _is_param(n)], lr, beta1)

    # -- plumbing
    def _grads(self, sd, loss):
        names = [n for n, v in sd.items() if v.requires_grad]
        gs = torch.autograd.grad(loss, [sd[n] for n in names], allow_unused=True)
        return {n: (g.clone() if g is not None else None) for n, g in zip(names, gs)}

    def _G(self, b, noise, all_gen_out, k, mask, gen_idxs, training=True):
        return generator_forward(self.G, self.n_gens, b["in_xy"], b["in_dxdy"], b["seq_start_end"],
                                 noise, all_gen_out, b.get("features"), k, mask, gen_idxs, training,
                                 self.use_pinet, self.social_mode)

    def _D(self, b, pxy, pdxdy, mask, training=True):
        return discriminator_forward(self.D, b["in_xy"], b["in_dxdy"], pxy, pdxdy, b["seq_start_end"],
                                     b.get("features"), mask, training, self.social_mode,
                                     unbound_output=self.gan_obj in ("W", "LS"))

    @staticmethod
    def loss_mask(b):
        """abstract_train.py:130-132."""
        mask = ~torch.isnan(b["gt_xy"]).any(2).any(0)
        return mask, b["gt_xy"][:, mask], b["gt_dxdy"][:, mask]

    def _phi(self, which, d, l_real, l_fake):
        """phi_1 / phi_2 / phi_3 of abstract_train.py:61-79, reduction="none"."""
        full = lambda v: torch.full_like(d, v)
        if self.gan_obj == "LS":
            return (d - full(l_fake if which == 2 else l_real)) ** 2
        if which == 1:
            return _bce(d, full(l_real))
        if which == 2:
            return _bce(d, full(l_fake))
        return _bce(d, full(l_real)) if self.gan_obj == "NS" else -_bce(d, full(l_fake))

    # -- steps
    def discriminator_step(self, b, noise, gen_idxs, labels_real, labels_fake):
        """labels_real = (l_real, l_fake) drawn for the real pass, labels_fake for the fake pass
        (utils.py:18-25: fake is drawn first, then real)."""
        mask, gt_xy, gt_dxdy = self.loss_mask(b)
        real, _ = self._D(b, gt_xy, gt_dxdy, mask)
        real_loss = self._phi(1, real, *labels_real).mean()
        with torch.no_grad():
            (rel, ab), _, idx = self._G(b, noise, False, 1, mask, gen_idxs)
        fake, branch = self._D(b, ab, rel, mask)
        # gan_type "gan" has no classifier term (train.py:181-186)
        ce = F.cross_entropy(branch.flatten(0, 1), idx.flatten()) if branch is not None else fake.new_zeros(())
        fake_loss = self._phi(2, fake, *labels_fake).mean()
        loss = ce + real_loss + fake_loss
        grads = self._grads(self.D, loss)
        norm = clip_grad_norm(grads, self.clip_d)
        self.optD.step(self.D, grads)
        return {"loss": loss.detach(), "ce": ce.detach(), "real": real_loss.detach(),
                "fake": fake_loss.detach(), "grads": grads, "grad_norm": norm,
                "d_real": real.detach(), "d_fake": fake.detach(), "fake_abs": ab}

    def generator_step(self, b, noise, gen_idxs, labels, count_scale=1):
        """count_scale (tests only): multiplies the per-generator draw counts of the reweighting (train.py:92-96), as if
        the batch were `count_scale` copies of itself -- the reweighted terms are sums of loss / count divided by the number
        of draws, so they are NOT invariant under replicating the batch (they shrink by 1 / copies)."""
        mask, gt_xy, gt_dxdy = self.loss_mask(b)
        N = b["in_xy"].shape[1]
        (rel, ab), _, idx = self._G(b, noise, False, self.k, mask, gen_idxs)
        l2 = (ab - gt_xy[:, None]).norm(dim=-1)
        if self.l2_loss_type == "mse":
            l2 = l2 ** 2
        l2 = l2.sum(0)                                                   # (k, N_act)
        min_l2 = 0.0
        for a, e in b["seq_start_end"]:                                  # un-adjusted ranges, train.py:67-71
            min_l2 = min_l2 + l2[:, a:e].sum(1).min()
        min_l2 = min_l2 / N
        out, branch = self._D(b, ab, rel, mask)
        counts = torch.bincount(idx.flatten(), minlength=self.n_gens).to(out.dtype) * count_scale
        w = 1.0 / counts[idx]                                            # train.py:92-96
        adv = (self._phi(3, out, *labels) * w).mean()
        clf = ((F.cross_entropy(branch.flatten(0, 1), idx.reshape(-1), reduction="none").reshape(idx.shape) * w).mean()
               if branch is not None else out.new_zeros(()))                # train.py:101-113
        loss = self.l2_w * min_l2 + adv + self.clf_w * clf
        grads = self._grads(self.G, loss)
        norm = clip_grad_norm(grads, self.clip_g)
        self.optG.step(self.G, grads)
        return {"loss": loss.detach(), "l2": min_l2.detach(), "adv": adv.detach(), "clf": clf.detach(),
                "grads": grads, "grad_norm": norm, "abs": ab.detach(), "rel": rel.detach(),
                "d_out": out.detach(), "branch": branch.detach() if branch is not None else None}

    def net_chooser_step(self, b, noise, k_exp=1):
        mask, gt_xy, gt_dxdy = self.loss_mask(b)
        (rel, ab), logits, _ = self._G(b, noise, True, k_exp, mask,
                                        torch.zeros(int(mask.sum()), k_exp, dtype=torch.long))
        if self.weighting_target == "ml":
            d = ab - gt_xy[:, None, None]
            logp = (-(d * d) / (2 * self.sigma ** 2) - math.log(self.sigma) - math.log(math.sqrt(2 * math.pi)))
            logp = logp.sum([0, -1]).mean(0).t()                             # (N_act, G)
            target = torch.softmax(logp, 1)
            loss = -(target * torch.softmax(logits, 1).log()).sum(1).mean()
        elif self.weighting_target in ("l2", "endpoint"):                    # train.py:618-624, :641-647
            if self.weighting_target == "l2":
                dist = torch.norm(ab - gt_xy[:, None, None], p=2, dim=-1).mean(0)
            else:
                dist = torch.norm(ab[-1] - gt_xy[-1, None, None], p=2, dim=-1)
            loss = F.cross_entropy(logits, torch.argmin(dist.min(0)[0].transpose(0, 1), dim=1))
        elif self.weighting_target == "mgan":                                # train.py:604-614
            _, branch = self._D(b, gt_xy, gt_dxdy, mask)
            out_probs = torch.softmax(logits, 1)
            loss = -(torch.softmax(branch, 1) * out_probs.log()).sum(1).mean()
            loss = loss - (0.9 ** self.epoch) * -(out_probs * out_probs.log()).sum(1).mean()
        else:
            raise ValueError(self.weighting_target)
        grads = self._grads(self.G, loss * self.pi_w)
        self.optG.step(self.G, grads)
        return {"loss": loss.detach(), "grads": grads, "logits": logits.detach(), "abs": ab.detach()}


# --------------------------------------------------------------------------- metrics
def ade_fde(preds, gt, sub_batches):
    """compute_metrics_from_batch(mode="raw") (metrics.py:99-141): scene-level min over k.

    preds (T,k,N,2), gt (T,N,2) -> dict name -> (sum, count)."""
    err = (preds - gt[:, None]).norm(dim=-1)                            # (T, k, N)
    ade, fde = err.sum(0), err[-1]
    T, k, N = err.shape
    sa = sum(float(ade[:, a:b].sum(1).min()) for a, b in sub_batches)
    sf = sum(float(fde[:, a:b].sum(1).min()) for a, b in sub_batches)
    return {"ADE": (sa, T * N), "FDE": (sf, N)}
