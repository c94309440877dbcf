"""torch.autograd.Function wrappers over the C ABI (include/mggan_b200.h).

PyTorch allocates the outputs / workspaces and carries the autograd graph; all arithmetic on
the path happens in the sm_100a kernels of `csrc/`.  Nothing here runs without the library.
"""
import math
import os

import torch

from . import cuda_ext as X
from .cuda_ext import call, ptr

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_SIGMOID_EPS = 0, 1, 2, 3
TILE = 64


def _f32(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _zeros(device, *shapes):
    """Zero-filled fp32 tensors of the given shapes carved out of ONE allocation (one fill launch instead of one
    per gradient accumulator); every tensor starts on a 16-byte boundary."""
    sizes = [int(math.prod(sh)) for sh in shapes]
    padded = [(n + 3) & ~3 for n in sizes]
    flat = torch.zeros(sum(padded), device=device, dtype=torch.float32)
    out, off = [], 0
    for sh, n, pn in zip(shapes, sizes, padded):
        out.append(flat[off:off + n].view(sh))
        off += pn
    return out


# --------------------------------------------------------------------------- dense layer
class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, act, slope, grad_on):
        x, w = _f32(x), _f32(w)
        b = _f32(b) if b is not None else None
        M, K = x.shape
        O = w.shape[0]
        y = torch.empty(M, O, device=x.device, dtype=torch.float32)
        call("mggan_linear_fwd", ptr(x), M, K, ptr(w), ptr(b), O, act, float(slope), ptr(y))
        if grad_on:
            ctx.save_for_backward(x, w, y)
        ctx.act, ctx.slope, ctx.has_bias = act, float(slope), b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = _f32(dy)
        M, K = x.shape
        O = w.shape[0]
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        dx = torch.empty_like(x) if need_x else None
        dw, db = _zeros(x.device, w.shape, (O,)) if (need_w or need_b) else (None, None)
        call("mggan_linear_bwd", ptr(x), M, K, ptr(w), O, ctx.act, ctx.slope, ptr(y), ptr(dy), ptr(dx), ptr(dw), ptr(db))
        return dx, (dw if need_w else None), (db if need_b else None), None, None, None


def linear(x, w, b=None, act=ACT_NONE, slope=0.0):
    """act(x @ w.T + b) for 2-D x."""
    return _Linear.apply(x, w, b, act, slope, torch.is_grad_enabled())


# --------------------------------------------------------------------------- fused two-layer perceptron
class _Mlp2(torch.autograd.Function):
    """act2(W2 act1(W1 x + b1) + b2) in one launch; the backward recomputes the hidden layer (csrc/mlp2.cu)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, act1, slope1, act2, slope2, grad_on):
        x, w1, w2 = _f32(x), _f32(w1), _f32(w2)
        b1 = _f32(b1) if b1 is not None else None
        b2 = _f32(b2) if b2 is not None else None
        M, K = x.shape
        H, O = w1.shape[0], w2.shape[0]
        y = torch.empty(M, O, device=x.device, dtype=torch.float32)
        call("mggan_mlp2_fwd", ptr(x), M, K, ptr(w1), ptr(b1), H, act1, float(slope1), ptr(w2), ptr(b2), O, act2,
             float(slope2), ptr(y))
        if grad_on:
            ctx.save_for_backward(x, w1, b1, w2, y)
        ctx.cfg = (act1, float(slope1), act2, float(slope2), b1 is not None, b2 is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w1, b1, w2, y = ctx.saved_tensors
        act1, slope1, act2, slope2, has_b1, has_b2 = ctx.cfg
        M, K = x.shape
        H, O = w1.shape[0], w2.shape[0]
        need_x = ctx.needs_input_grad[0]
        need_w = any(ctx.needs_input_grad[1:5])
        dx = torch.empty_like(x) if need_x else None
        dw1 = db1 = dw2 = db2 = None
        if need_w:
            dw1, db1, dw2, db2 = _zeros(x.device, w1.shape, (H,), w2.shape, (O,))
        if need_x or need_w:
            call("mggan_mlp2_bwd", ptr(x), M, K, ptr(w1), ptr(b1), H, act1, slope1, ptr(w2), O, act2, slope2, ptr(y),
                 ptr(_f32(dy)), ptr(dx), ptr(dw1), ptr(db1), ptr(dw2), ptr(db2))
        g = ctx.needs_input_grad
        return (dx, dw1 if g[1] else None, db1 if (g[2] and has_b1) else None, dw2 if g[3] else None,
                db2 if (g[4] and has_b2) else None, None, None, None, None, None)


def mlp2_supported(K, H, O):
    return 4 <= K <= 192 and K % 4 == 0 and 4 <= H <= 96 and H % 4 == 0 and 1 <= O <= 32


def mlp2(x, w1, b1, w2, b2, act1=ACT_NONE, slope1=0.0, act2=ACT_NONE, slope2=0.0):
    """act2(act1(x @ w1.T + b1) @ w2.T + b2) for 2-D x: one fused launch when the sizes fit the kernel (they do for every
    chain of the default model), else two dense-layer launches."""
    if not mlp2_supported(x.shape[1], w1.shape[0], w2.shape[0]) or x.shape[0] == 0:
        return linear(linear(x, w1, b1, act1, slope1), w2, b2, act2, slope2)
    return _Mlp2.apply(x, w1, b1, w2, b2, act1, slope1, act2, slope2, torch.is_grad_enabled())


# --------------------------------------------------------------------------- discriminator heads (frozen weights)
class _DiscHeads(torch.autograd.Function):
    """p, branch = heads(base[i] + (s == 0) soc0[i] + W1p pe[s, i]) -- see include/mggan_b200.h.  Gradients flow to
    pe, base and soc0 only: the weights must not require grad (generator step / evaluation)."""

    @staticmethod
    def forward(ctx, pe, base, soc0, w1p, wd2, bd2, wg2, bg2, n, k, grad_on):
        ctx.set_materialize_grads(False)        # an unused output's gradient arrives as None (the kernel takes NULL), not as a zero fill
        assert not (grad_on and any(ctx.needs_input_grad[3:8])), \
            "disc_heads: weight gradients are not produced (frozen discriminator only)"
        pe, base, soc0, w1p, wd2, bd2 = map(_f32, (pe, base, soc0, w1p, wd2, bd2))
        G = 0 if wg2 is None else wg2.shape[0]
        HH = wd2.shape[-1]
        if G:
            wg2, bg2 = _f32(wg2), _f32(bg2)
        assert pe.shape == (n * k, 32) and base.shape == soc0.shape == (n, w1p.shape[0])
        p = torch.empty(n * k, device=pe.device, dtype=torch.float32)
        branch = torch.empty(n * k, G, device=pe.device, dtype=torch.float32) if G else None
        call("mggan_disc_heads_fwd", ptr(pe), n, k, HH, ptr(base), ptr(soc0), ptr(w1p), ptr(wd2), ptr(bd2), ptr(wg2),
             ptr(bg2), G, ptr(p), ptr(branch))
        if grad_on:
            ctx.save_for_backward(pe, base, soc0, w1p, wd2, wg2, p)
            ctx.dims = (n, k, HH, G)
        if G:
            return p, branch
        return p, p.new_zeros(0)

    @staticmethod
    def backward(ctx, dp, dbranch):
        pe, base, soc0, w1p, wd2, wg2, p = ctx.saved_tensors
        n, k, HH, G = ctx.dims
        if dp is None and dbranch is None:
            return (None,) * 11
        d_pe = torch.empty_like(pe)
        d_soc0 = torch.empty_like(soc0)
        d_base = torch.empty_like(base) if ctx.needs_input_grad[1] else None     # written in full by the kernel (plain stores)
        dp = _f32(dp) if dp is not None else None
        dbranch = _f32(dbranch) if (G and dbranch is not None) else None
        call("mggan_disc_heads_bwd", ptr(pe), n, k, HH, ptr(base), ptr(soc0), ptr(w1p), ptr(wd2), ptr(wg2), G, ptr(p),
             ptr(dp), ptr(dbranch), ptr(d_pe), ptr(d_soc0), ptr(d_base))
        return (d_pe, d_base, d_soc0) + (None,) * 8


def disc_heads(pe, base, soc0, w1p, wd2, bd2, wg2, bg2, n, k):
    """-> (p (k*n,), branch (k*n, G) or None)."""
    p, br = _DiscHeads.apply(pe, base, soc0, w1p, wd2, bd2, wg2, bg2, n, k, torch.is_grad_enabled())
    return p, (br if wg2 is not None else None)


# --------------------------------------------------------------------------- encoder LSTM
class _LstmSeq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wx, b, whh, grad_on):
        x, wx, b, whh = _f32(x), _f32(wx), _f32(b), _f32(whh)
        T, N, _ = x.shape
        H = whh.shape[1]
        hT = torch.empty(N, H, device=x.device, dtype=torch.float32)
        need = grad_on and any(ctx.needs_input_grad[1:4])
        acts = torch.empty(T, N, 6, H, device=x.device, dtype=torch.float32) if need else None
        call("mggan_lstm_seq_fwd", ptr(x), T, N, H, ptr(wx), ptr(b), ptr(whh), ptr(hT), ptr(acts))
        ctx.save_for_backward(x, whh, acts)
        return hT

    @staticmethod
    def backward(ctx, dh):
        x, whh, acts = ctx.saved_tensors
        T, N, _ = x.shape
        H = whh.shape[1]
        dwx, db, dwhh = _zeros(x.device, (4 * H, 2), (4 * H,), whh.shape)
        call("mggan_lstm_seq_bwd", ptr(x), T, N, H, ptr(whh), ptr(acts), ptr(_f32(dh)), ptr(dwx), ptr(db), ptr(dwhh))
        return None, dwx, db, dwhh, None


def lstm_encode(x, w_emb, b_emb, w_ih, w_hh, b_ih, b_hh):
    """TrajectoryEncoder: Linear(2,E) then 1-layer LSTM, returns h_T (N,H).

    The embedding is folded into the input projection (tiny host-side products on the
    weights, tracked by autograd); the recurrence itself is one kernel."""
    wx = w_ih @ w_emb                               # (4H, 2)
    b = w_ih @ b_emb + b_ih + b_hh                  # (4H,)
    return _LstmSeq.apply(x, wx, b, w_hh, torch.is_grad_enabled())


# --------------------------------------------------------------------------- scenes (CSR)
class SceneIndex:
    """`seq_start_end` (python list of [start, end)) as device CSR arrays, built once per batch."""

    def __init__(self, sub_batches, device):
        off = [int(sub_batches[0][0])] if len(sub_batches) else [0]
        for a, b in sub_batches:
            assert int(a) == off[-1], "scenes must be contiguous and ordered"
            off.append(int(b))
        self.sub_batches = [(int(a), int(b)) for a, b in sub_batches]
        self.n_scenes = len(sub_batches)
        self.n_agents = off[-1] - off[0]
        sizes = [b - a for a, b in self.sub_batches]
        pair = [0]
        for s in sizes:
            pair.append(pair[-1] + s * s)
        self.n_pairs = pair[-1]
        self.max_size = max(sizes) if sizes else 0
        self.scene_off = torch.tensor(off, dtype=torch.int32, device=device)
        self.pair_off = torch.tensor(pair, dtype=torch.int32, device=device)
        self._pairs = None

    def pair_index(self):
        """(ia, ib) int64 device tensors over the sum n_s^2 ordered in-scene pairs, scene-major then a-major (pair
        a * n_s + b of a scene), built on first use (the pooling variant `--pool_type sgan` gathers with them)."""
        if self._pairs is None:
            import numpy as np
            ia, ib = [], []
            for a, b in self.sub_batches:
                r = np.arange(a, b, dtype=np.int64)
                ia.append(np.repeat(r, b - a))
                ib.append(np.tile(r, b - a))
            cat = (lambda v: np.concatenate(v) if v else np.zeros(0, np.int64))
            dev = self.scene_off.device
            self._pairs = (torch.from_numpy(cat(ia)).to(dev), torch.from_numpy(cat(ib)).to(dev))
        return self._pairs

    _cache = {}

    @classmethod
    def get(cls, sub_batches, device):
        if isinstance(sub_batches, SceneIndex):
            return sub_batches
        key = (tuple((int(a), int(b)) for a, b in sub_batches), str(device))
        hit = cls._cache.get(key)
        if hit is None:
            if len(cls._cache) > 1024:       # ragged datasets: one entry per batch structure (a few hundred bytes each)
                cls._cache.clear()
            hit = cls._cache[key] = SceneIndex(sub_batches, device)
        return hit


# --------------------------------------------------------------------------- social attention
class _SocialAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x4, h, us, w1, b1, w2, b2, scenes, grad_on):
        x4, h, us, w1, b1, w2, b2 = map(_f32, (x4, h, us, w1, b1, w2, b2))
        N, HD = h.shape
        S = torch.empty_like(h)
        att = torch.empty(max(scenes.n_pairs, 1), device=h.device, dtype=torch.float32)
        call("mggan_social_attn_fwd", ptr(x4), ptr(h), HD, ptr(us), ptr(scenes.scene_off), ptr(scenes.pair_off),
             scenes.n_scenes, ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(S), ptr(att))
        if grad_on:
            ctx.save_for_backward(x4, h, us, w1, b1, w2, b2, att)
            ctx.scenes = scenes
        return S

    @staticmethod
    def backward(ctx, dS):
        x4, h, us, w1, b1, w2, b2, att = ctx.saved_tensors
        sc = ctx.scenes
        N, HD = h.shape
        dsig = torch.empty_like(att)
        dh, dus, dw1, db1, dw2, db2 = _zeros(h.device, h.shape, us.shape, w1.shape, b1.shape, w2.shape, b2.shape)
        call("mggan_social_attn_bwd", ptr(x4), ptr(h), HD, ptr(us), ptr(sc.scene_off), ptr(sc.pair_off), sc.n_scenes,
             ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(att), ptr(_f32(dS)), ptr(dsig), ptr(dh), ptr(dus), ptr(dw1),
             ptr(db1), ptr(dw2), ptr(db2))
        return None, dh, dus, dw1, db1, dw2, db2, None, None


def social_attention(xy_last, dxdy_last, h, scenes, fc0, fc2, fc4, att_w):
    """SocialAttention.forward on the rows covered by `scenes`.

    fc0/fc2/fc4: the three Linear layers of EmbedSocialFeatures, att_w: AttentionPooling.W.
    The last embedding layer and W are folded into per-agent vectors with one dense-layer
    kernel: Us = [fc4.weight^T ; fc4.bias] (W h + b) (N,65)."""
    x4 = torch.cat([xy_last, dxdy_last], -1)
    # us = [fc4.weight^T ; fc4.bias] (W h + b): the two weight matrices are folded on the host side (a (65, F) x (F, H) product,
    # tracked by autograd), so the per-agent work is ONE dense layer instead of two
    wc3 = torch.cat([fc4.weight.t(), fc4.bias[None]], 0)        # (65, F)
    us = linear(h, wc3 @ att_w.weight, wc3 @ att_w.bias)
    return _SocialAttn.apply(x4, h, us, fc0.weight, fc0.bias, fc2.weight, fc2.bias, scenes, torch.is_grad_enabled())


# --------------------------------------------------------------------------- scene (physical) attention
BN_EPS, BN_MOMENTUM = 1e-5, 0.1


def _allreduce(t, group):
    if group is not None:
        from .distributed import allreduce_tensor
        allreduce_tensor(t, group)


class PatchStats:
    """Data-only statistics of a batch of crops (R 36x36, P 36; see csrc/scene.cu), computed once and
    shared by every scene-CNN forward / backward that sees the same `img` tensor (and row selection)."""
    _cache = []          # [(weakref(img), version, rows_key, PatchStats)]
    _n_total = {}        # (local crops, group) -> global crop count

    def __init__(self, img, rows, group):
        dev = img.device
        n = int(rows.numel()) if rows is not None else img.shape[0]
        buf = torch.zeros(36 * 36 + 36 + 1, device=dev, dtype=torch.float64)
        self.buf = buf
        self.R_local, self.P_local = buf[:1296], buf[1296:1332]
        call("mggan_scene_patch_stats", ptr(img), ptr(rows), n, ptr(self.R_local), ptr(self.P_local))
        buf[1332:1333].fill_(float(n))          # a fill kernel (a python-scalar assignment is a host copy: not capturable)
        if group is not None:
            g = buf.clone()
            _allreduce(g, group)
            self.R, self.P = g[:1296], g[1296:1332]
            key = (n, id(group))
            if torch.cuda.is_current_stream_capturing():      # no collective on the host inside a CUDA-graph capture: the
                self.n_total = PatchStats._n_total[key]       # global crop count of this structure was summed when it ran eagerly
            else:                                             # the crop count is a host number: summed on the host (gloo)
                from .distributed import host_sum
                self.n_total = PatchStats._n_total[key] = host_sum([n], group)[0]
        else:
            self.R, self.P, self.n_total = self.R_local, self.P_local, float(n)

    @classmethod
    def get(cls, img, rows, rows_key, group):
        import weakref
        for ref, ver, rk, ps in cls._cache:
            if ref() is img and ver == img._version and rk == rows_key:
                return ps
        ps = PatchStats(img, rows, group)
        cls._cache.append((weakref.ref(img), img._version, rows_key, ps))
        cls._cache[:] = [e for e in cls._cache if e[0]() is not None][-4:]
        return ps


class _SceneAttn(torch.autograd.Function):
    """AttentionGlobal.forward.  Non-tensor state: the two BatchNorm modules (running buffers are
    updated in place, like nn.BatchNorm2d in train mode), `rows` (int32 gather or None), training
    flag and the process group whose ranks share BatchNorm statistics (None = local)."""

    @staticmethod
    def forward(ctx, img, c1w, c1b, g1, be1, c2w, c2b, g2, be2, a0w, a0b, a2w, a2b, bn1, bn2, rows, rows_key,
                training, group, grad_on, holder=None):
        ws = [_f32(t) for t in (c1w, c1b, g1, be1, c2w, c2b, g2, be2, a0w, a0b, a2w, a2b)]
        c1w, c1b, g1, be1, c2w, c2b, g2, be2, a0w, a0b, a2w, a2b = ws
        assert img.dtype == torch.float32 and img.is_contiguous()
        dev = img.device
        C = c1w.shape[0]
        N = int(rows.numel()) if rows is not None else img.shape[0]
        need = grad_on and any(ctx.needs_input_grad)
        x2 = torch.empty(N, C, 16, 16, device=dev, dtype=torch.float32)
        out = torch.empty(N, 64, device=dev, dtype=torch.float32)
        ab1, mi1 = torch.empty(2 * C, device=dev), torch.empty(2 * C, device=dev)
        ab2, mi2 = torch.empty(2 * C, device=dev), torch.empty(2 * C, device=dev)
        tr = 1 if training else 0
        ps = PatchStats.get(img, rows, rows_key, group) if training else None
        n_total = ps.n_total if ps is not None else float(N)
        call("mggan_scene_bn1_from_patches", ptr(ps.R) if ps else None, ptr(ps.P) if ps else None, n_total * 1089.0, C,
             ptr(c1w), ptr(c1b), ptr(g1), ptr(be1), ptr(bn1.running_mean), ptr(bn1.running_var),
             ptr(bn1.num_batches_tracked), BN_MOMENTUM, BN_EPS, tr, ptr(ab1), ptr(mi1))
        st2 = torch.zeros(2 * C, device=dev, dtype=torch.float64) if training else None
        e1 = torch.empty(N, C, 256, device=dev, dtype=torch.float32) if need else None
        idx1 = torch.empty(N, C, 256, device=dev, dtype=torch.uint8) if need else None
        call("mggan_scene_fused12_fwd", ptr(img), ptr(rows), N, C, ptr(c1w), ptr(c1b), ptr(ab1), ptr(c2w), ptr(c2b),
             ptr(x2), ptr(st2), ptr(e1), ptr(idx1))
        if training:
            _allreduce(st2, group)
        call("mggan_scene_bn_finalize", ptr(st2), n_total * 256.0, C, ptr(g2), ptr(be2), ptr(bn2.running_mean),
             ptr(bn2.running_var), ptr(bn2.num_batches_tracked), BN_MOMENTUM, BN_EPS, tr, ptr(ab2), ptr(mi2))
        call("mggan_scene_attn_fwd", ptr(x2), N, C, ptr(ab2), ptr(a0w), ptr(a0b), ptr(a2w), ptr(a2b), ptr(out))
        if holder is not None and training:
            holder.update(ps=ps, st2=st2, n_total=n_total, C=C)
        if need:
            assert training, "scene-attention backward needs train-mode BatchNorm (batch statistics)"
            ctx.save_for_backward(img, c1w, c1b, c2w, a0w, a0b, a2w, a2b, x2, e1, idx1, ab1, mi1, ab2, mi2)
            ctx.rows, ctx.group, ctx.n_total, ctx.N, ctx.C, ctx.ps = rows, group, n_total, N, C, ps
        return out

    @staticmethod
    def backward(ctx, dout):
        img, c1w, c1b, c2w, a0w, a0b, a2w, a2b, x2, e1, idx1, ab1, mi1, ab2, mi2 = ctx.saved_tensors
        N, C, dev, group, ps = ctx.N, ctx.C, img.device, ctx.group, ctx.ps
        dout = _f32(dout)
        z = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)
        da0w, da0b, da2w, da2b, dg2, db2, dc2w, dc2b, S1, dc1b = _zeros(dev, (32, C), (32,), (C, 32), (C,), (C,), (C,),
                                                                        (C, C, 3, 3), (C,), (C, 36), (C,))
        sums2, sums1 = torch.zeros(2, 2 * C, device=dev, dtype=torch.float64).unbind(0)
        dy2 = torch.empty(N, C, 64, device=dev)
        idx2 = torch.empty(N, C, 64, device=dev, dtype=torch.uint8)
        call("mggan_scene_attn_bwd", ptr(x2), N, C, ptr(ab2), ptr(mi2), ptr(a0w), ptr(a0b), ptr(a2w), ptr(a2b),
             ptr(dout), ptr(da0w), ptr(da0b), ptr(da2w), ptr(da2b), ptr(dy2), ptr(idx2), ptr(sums2))
        # BatchNorm affine gradients are sums over the local shard; the means use global sums
        m12_2 = torch.empty(2 * C, device=dev)
        loc2 = sums2.clone() if group is not None else sums2
        _allreduce(sums2, group)
        call("mggan_scene_bn_bwd_finalize", ptr(sums2), ctx.n_total * 256.0, C, ptr(m12_2), ptr(dg2), ptr(db2))
        if group is not None:
            dg2, db2 = loc2[C:].float(), loc2[:C].float()
        call("mggan_scene_fused12_bwd", ptr(img), ptr(ctx.rows), N, C, ptr(x2), ptr(e1), ptr(idx1), ptr(ab1), ptr(mi1),
             ptr(ab2), ptr(mi2), ptr(m12_2), ptr(c2w), ptr(dy2), ptr(idx2), ptr(dc2w), ptr(dc2b), ptr(S1), ptr(sums1))
        glob1 = sums1
        if group is not None:
            glob1 = sums1.clone()
            _allreduce(glob1, group)
        dc1w, dg1, db1 = torch.empty(C, 4, 3, 3, device=dev), torch.empty(C, device=dev), torch.empty(C, device=dev)
        call("mggan_scene_bn1_bwd_finalize", ptr(glob1), ptr(sums1), ctx.n_total * 1089.0, C, ptr(S1), ptr(ps.R_local),
             ptr(ps.P_local), ptr(c1w), ptr(c1b), ptr(ab1), ptr(mi1), ptr(dc1w), ptr(dg1), ptr(db1))
        # dc1b stays zero: conv1's bias gradient vanishes identically under train-mode BatchNorm
        return (None, dc1w, dc1b, dg1, db1, dc2w, dc2b, dg2, db2, da0w, da0b, da2w, da2b) + (None,) * 8


def _replay_bn_updates(mod, holder):
    """A second forward of the same crops through the same weights: only its side effect is left to do --
    the train-mode running-statistics update of both BatchNorm layers (momentum applied once more)."""
    b1, b2 = mod.CNN.encoder.ConvBlock_1.Block, mod.CNN.encoder.ConvBlock_2.Block
    C, ps, dev = holder["C"], holder["ps"], holder["st2"].device
    scratch = torch.empty(4 * C, device=dev)
    call("mggan_scene_bn1_from_patches", ptr(ps.R), ptr(ps.P), holder["n_total"] * 1089.0, C, ptr(_f32(b1.Conv_1.weight)),
         ptr(_f32(b1.Conv_1.bias)), ptr(_f32(b1.BN_1.weight)), ptr(_f32(b1.BN_1.bias)), ptr(b1.BN_1.running_mean),
         ptr(b1.BN_1.running_var), ptr(b1.BN_1.num_batches_tracked), BN_MOMENTUM, BN_EPS, 1, ptr(scratch[:2 * C]),
         ptr(scratch[2 * C:]))
    call("mggan_scene_bn_finalize", ptr(holder["st2"]), holder["n_total"] * 256.0, C, ptr(_f32(b2.BN_1.weight)),
         ptr(_f32(b2.BN_1.bias)), ptr(b2.BN_1.running_mean), ptr(b2.BN_1.running_var), ptr(b2.BN_1.num_batches_tracked),
         BN_MOMENTUM, BN_EPS, 1, ptr(scratch[:2 * C]), ptr(scratch[2 * C:]))


def scene_attention(img, mod, rows=None, group=None, rows_key=None):
    """mod: an AttentionGlobal parameter container (mggan.model.modules.cnn).  `rows_key`: hashable
    identity of the row selection (None = all rows), part of the patch-statistics cache key.

    When `mod.memo` is a dict (set by a caller that knows the weights do not change between calls, e.g.
    the discriminator step's real / fake passes), a repeated call on the same crops returns the SAME output
    tensor (one autograd node: gradients of both uses add up) and only replays the BatchNorm running-statistics
    update."""
    b1, b2 = mod.CNN.encoder.ConvBlock_1.Block, mod.CNN.encoder.ConvBlock_2.Block
    a = mod.cnn_attention
    if img.dtype != torch.float32 or not img.is_contiguous():
        img = img.float().contiguous()
    memo = getattr(mod, "memo", None)
    key = (id(img), img._version, rows_key, torch.is_grad_enabled(), mod.training)
    if memo is not None and key in memo:
        out, holder = memo[key]
        if mod.training:
            _replay_bn_updates(mod, holder)
        return out
    holder = {} if memo is not None else None
    out = _SceneAttn.apply(img, b1.Conv_1.weight, b1.Conv_1.bias, b1.BN_1.weight, b1.BN_1.bias,
                           b2.Conv_1.weight, b2.Conv_1.bias, b2.BN_1.weight, b2.BN_1.bias,
                           a[0].weight, a[0].bias, a[2].weight, a[2].bias, b1.BN_1, b2.BN_1, rows, rows_key,
                           mod.training, group, torch.is_grad_enabled(), holder)
    if memo is not None:
        memo[key] = (out, holder)
    return out


# --------------------------------------------------------------------------- generator selection
class Selection:
    """Decoder work list (device arrays) for a set of (agent, generator, noise sample, slot) sequences."""

    def __init__(self, n_agents, k, num_gens, n_tiles, tile_gen, seq_agent, seq_noise, seq_out, n_cols, totals=None):
        self.n_agents, self.k, self.num_gens, self.n_tiles = n_agents, k, num_gens, n_tiles
        self.tile_gen, self.seq_agent, self.seq_noise, self.seq_out = tile_gen, seq_agent, seq_noise, seq_out
        self.n_cols, self.totals = n_cols, totals

    @staticmethod
    def from_indices(idx, num_gens):
        """idx (n_act, k) int64: sampled generator per agent and sample (get_selection_indices + gather)."""
        idx = idx.contiguous()
        n, k = idx.shape
        dev = idx.device
        n_tiles = X.selection_tiles(n * k, num_gens)
        i32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.int32)
        cnt, base_row, totals = i32(max(n * num_gens, 1)), i32(num_gens + 1), i32(num_gens)
        rank = torch.empty(max(n * k, 1), device=dev, dtype=torch.uint8)
        err = torch.zeros(1, device=dev, dtype=torch.int32)
        tile_gen, seq_agent, seq_noise, seq_out = i32(n_tiles), i32(n_tiles * TILE), i32(n_tiles * TILE), i32(n_tiles * TILE)
        call("mggan_selection_build", ptr(idx), n, k, num_gens, n_tiles, ptr(cnt), ptr(rank), ptr(base_row), ptr(err),
             ptr(totals), ptr(tile_gen), ptr(seq_agent), ptr(seq_noise), ptr(seq_out))
        return Selection(n, k, num_gens, n_tiles, tile_gen, seq_agent, seq_noise, seq_out, k * n, totals)

    @staticmethod
    def all_generators(n, k, num_gens, device):
        tiles_per_gen = 2 * ((n * k + 2 * TILE - 1) // (2 * TILE))      # 128-row groups (mggan_selection_all)
        n_tiles = num_gens * tiles_per_gen
        i32 = lambda *s: torch.empty(*s, device=device, dtype=torch.int32)
        tile_gen = i32(max(n_tiles, 1))
        seq_agent, seq_noise, seq_out = (i32(max(n_tiles * TILE, 1)) for _ in range(3))
        call("mggan_selection_all", n, k, num_gens, ptr(tile_gen), ptr(seq_agent), ptr(seq_noise), ptr(seq_out))
        return Selection(n, k, num_gens, n_tiles, tile_gen, seq_agent, seq_noise, seq_out, k * num_gens * n)


def gumbel_sample(logits, k, seed, offset, dyn_offset=None):
    """dyn_offset: optional device int64 scalar added to `offset` inside the kernel (CUDA-graph replay)."""
    logits = _f32(logits.detach())
    n, G = logits.shape
    idx = torch.empty(n, k, device=logits.device, dtype=torch.int64)
    call("mggan_gumbel_sample", ptr(logits), n, k, G, int(seed), int(offset), ptr(dyn_offset), ptr(idx))
    return idx


# --------------------------------------------------------------------------- decoder
# Forward entry point: the tcgen05 (tensor-core, 3 x TF32) kernel by default; MGGAN_DECODER=fp32 selects the FP32-FMA
# kernel with the same contract (kept for A/B measurements; both are CUDA, neither is a fallback).
DECODER_FWD = "mggan_decoder_fwd" if os.environ.get("MGGAN_DECODER", "tc") == "fp32" else "mggan_decoder_fwd_tc"


class _Decoder(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, social, last_xy, last_dxdy, noise, wz, wx, b, whh, w1h, w1s, b1, w2, b2, sel, pred_len, grad_on):
        ctx.set_materialize_grads(False)        # d_abs / d_rel of an unused output arrive as None (the kernel takes NULL)
        ts = [_f32(t) for t in (A, social, last_xy, last_dxdy, noise, wz, wx, b, whh, w1h, w1s, b1, w2, b2)]
        A, social, last_xy, last_dxdy, noise, wz, wx, b, whh, w1h, w1s, b1, w2, b2 = ts
        dev = A.device
        Z = wz.shape[1]
        noise = noise.reshape(-1, Z)
        out_abs = torch.empty(pred_len, sel.n_cols, 2, device=dev, dtype=torch.float32)
        out_rel = torch.empty_like(out_abs)
        need = grad_on and any(ctx.needs_input_grad)
        R = sel.n_tiles * TILE
        acts = torch.empty(pred_len, R, 32, 2, device=dev) if need else None       # (h_t, c_t): the backward recomputes the gates
        u1 = torch.empty(pred_len, R, 16, device=dev) if need else None
        h0 = torch.empty(R, 32, device=dev) if need else None
        call(DECODER_FWD, sel.n_tiles, ptr(sel.tile_gen), ptr(sel.seq_agent), ptr(sel.seq_noise),
             ptr(sel.seq_out), ptr(A), ptr(social), ptr(last_xy), ptr(last_dxdy), ptr(noise), Z, ptr(wz), ptr(wx),
             ptr(b), ptr(whh), ptr(w1h), ptr(w1s), ptr(b1), ptr(w2), ptr(b2), pred_len, sel.n_cols, ptr(out_abs),
             ptr(out_rel), ptr(acts), ptr(u1), ptr(h0))
        if need:
            ctx.save_for_backward(social, last_dxdy, noise, wz, wx, b, whh, w1h, w1s, b1, w2, b2, out_rel, acts, u1, h0)
            ctx.sel, ctx.pred_len, ctx.n_agents = sel, pred_len, A.shape[0]
        return out_abs, out_rel

    @staticmethod
    def backward(ctx, d_abs, d_rel):
        if d_abs is None and d_rel is None:
            return (None,) * 17
        social, last_dxdy, noise, wz, wx, b, whh, w1h, w1s, b1, w2, b2, out_rel, acts, u1, h0 = ctx.saved_tensors
        sel = ctx.sel
        Z = wz.shape[1]
        dwz, dwx, db, dwhh, dw1h, dw1s, db1, dw2, db2, dA, dsoc = _zeros(
            wz.device, wz.shape, wx.shape, b.shape, whh.shape, w1h.shape, w1s.shape, b1.shape, w2.shape, b2.shape,
            (ctx.n_agents, 32), (ctx.n_agents, 32))
        d_abs = _f32(d_abs) if d_abs is not None else None
        d_rel = _f32(d_rel) if d_rel is not None else None
        call("mggan_decoder_bwd", sel.n_tiles, ptr(sel.tile_gen), ptr(sel.seq_agent), ptr(sel.seq_noise),
             ptr(sel.seq_out), ptr(social), ptr(last_dxdy), ptr(noise), Z, ptr(wz), ptr(wx), ptr(b), ptr(whh),
             ptr(w1h), ptr(w1s), ptr(b1), ptr(w2), ptr(b2), ctx.pred_len, sel.n_cols, ptr(out_rel), ptr(acts), ptr(u1),
             ptr(h0), ptr(d_abs), ptr(d_rel), ptr(dwz), ptr(dwx), ptr(db), ptr(dwhh), ptr(dw1h), ptr(dw1s), ptr(db1),
             ptr(dw2), ptr(db2), ptr(dA), ptr(dsoc))
        return dA, dsoc, None, None, None, dwz, dwx, db, dwhh, dw1h, dw1s, db1, dw2, db2, None, None, None


def decode(A, social, last_xy, last_dxdy, noise, wz, gen_weights, sel, pred_len):
    """gen_weights: dict of stacked per-generator tensors (wx, b, whh, w1h, w1s, b1, w2, b2)."""
    g = gen_weights
    return _Decoder.apply(A, social, last_xy, last_dxdy, noise, wz, g["wx"], g["b"], g["whh"], g["w1h"], g["w1s"],
                          g["b1"], g["w2"], g["b2"], sel, pred_len, torch.is_grad_enabled())


# --------------------------------------------------------------------------- losses
class _L2SceneMin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, abs_, gt, scenes, inv_norm, squared=False):
        abs_, gt = _f32(abs_), _f32(gt)
        T, k, n, _ = abs_.shape
        loss = torch.zeros(1, device=abs_.device)
        best = torch.empty(max(scenes.n_scenes, 1), device=abs_.device, dtype=torch.int32)
        d_abs = torch.zeros_like(abs_) if ctx.needs_input_grad[0] else None
        call("mggan_l2_scene_min", ptr(abs_), ptr(gt), T, k, n, ptr(scenes.scene_off), scenes.n_scenes,
             float(inv_norm), 1 if squared else 0, ptr(loss), ptr(best), ptr(d_abs))
        ctx.save_for_backward(d_abs)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (d_abs,) = ctx.saved_tensors
        return d_abs * g, None, None, None, None


def l2_scene_min(abs_, gt, scenes, inv_norm, squared=False):
    """(1/N) sum_scenes min_s sum_{i in scene, t} |abs - gt|  (train.py:57-75); inv_norm = 1/N.
    squared: l2_loss_type "mse" (the per-step distances are squared, train.py:62-63)."""
    return _L2SceneMin.apply(abs_, gt, scenes, inv_norm, squared)


def _scalar_label_loss(entry):
    class _ScalarLabelLoss(torch.autograd.Function):
        @staticmethod
        def forward(ctx, p, label, gen_idx, counts, inv_denom):
            p = _f32(p)
            loss = torch.zeros(1, device=p.device)
            dp = torch.empty_like(p) if ctx.needs_input_grad[0] else None
            gi = gen_idx.contiguous() if gen_idx is not None else None
            call(entry, ptr(p), p.numel(), float(label), ptr(gi), ptr(counts), float(inv_denom), ptr(loss), ptr(dp))
            ctx.save_for_backward(dp)
            return loss[0]

        @staticmethod
        def backward(ctx, g):
            (dp,) = ctx.saved_tensors
            return dp * g, None, None, None, None

    return _ScalarLabelLoss


_BceScalar = _scalar_label_loss("mggan_bce_scalar_label")
_MseScalar = _scalar_label_loss("mggan_mse_scalar_label")


def bce_scalar_label(p, label, gen_idx=None, counts=None, inv_denom=None):
    """inv_denom * sum_i BCE(p_i, label) / counts[gen_idx_i]  (default inv_denom = 1/numel: a mean)."""
    if inv_denom is None:
        inv_denom = 1.0 / max(p.numel(), 1)
    return _BceScalar.apply(p, label, gen_idx, counts, inv_denom)


def mse_scalar_label(p, label, gen_idx=None, counts=None, inv_denom=None):
    """inv_denom * sum_i (p_i - label)^2 / counts[gen_idx_i]  (gan_obj LS)."""
    if inv_denom is None:
        inv_denom = 1.0 / max(p.numel(), 1)
    return _MseScalar.apply(p, label, gen_idx, counts, inv_denom)


class _CeGen(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, counts, inv_denom):
        logits = _f32(logits)
        n, G = logits.shape
        loss = torch.zeros(1, device=logits.device)
        dl = torch.empty_like(logits) if ctx.needs_input_grad[0] else None
        call("mggan_ce_generators", ptr(logits), n, G, ptr(target.contiguous()), ptr(counts), float(inv_denom),
             ptr(loss), ptr(dl))
        ctx.save_for_backward(dl)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None, None


def ce_generators(logits, target, counts=None, inv_denom=None):
    """Cross-entropy of (n, G) logits against generator ids, optionally weighted by 1/counts[target]."""
    if inv_denom is None:
        inv_denom = 1.0 / max(logits.shape[0], 1)
    return _CeGen.apply(logits, target.reshape(-1), counts, inv_denom)


class _PmMl(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, abs_all, gt, sigma, weight, inv_n):
        logits, abs_all, gt = _f32(logits), _f32(abs_all), _f32(gt)
        T, ks, G, n, _ = abs_all.shape
        loss = torch.zeros(1, device=logits.device)
        dl = torch.empty_like(logits)
        target = torch.empty_like(logits)
        call("mggan_pm_ml_loss", ptr(abs_all), ptr(gt), T, ks, G, n, ptr(logits), float(sigma), float(weight),
             float(inv_n), ptr(loss), ptr(dl), ptr(target))
        ctx.save_for_backward(dl)
        ctx.mark_non_differentiable(target)
        return loss[0], target

    @staticmethod
    def backward(ctx, g, _):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None, None, None, None


def pm_ml_loss(logits, abs_all, gt, sigma, weight=1.0, inv_n=None):
    """PM-Net "ml" objective (train.py:626-639).  Returns (unweighted loss, target); the gradient
    carries `weight` (= pi_net_loss_weight), as the reference backpropagates loss * weight."""
    if inv_n is None:
        inv_n = 1.0 / max(logits.shape[0], 1)
    return _PmMl.apply(logits, abs_all, gt, sigma, weight, inv_n)


# --------------------------------------------------------------------------- optimiser
def _tables(entries, dyn=None):
    """entries: list of (p, g, m, v, n, bc1, bc2_sqrt) -> ctypes tables of <= TABLE_MAX rows.
    dyn: optional list of device pointers (one per table) to [lr, bc1[64], bc2_sqrt[64]] overrides."""
    out = []
    for s in range(0, len(entries), X.TABLE_MAX):
        chunk = entries[s:s + X.TABLE_MAX]
        tb = X.TensorTable()
        tb.dyn = dyn[s // X.TABLE_MAX] if dyn is not None else None
        for i, (p, g, m, v, n, bc1, bc2s) in enumerate(chunk):
            tb.p[i], tb.g[i], tb.m[i], tb.v[i] = p, g, m, v
            tb.n[i], tb.bc1[i], tb.bc2_sqrt[i] = n, bc1, bc2s
        out.append((tb, len(chunk)))
    return out


def grad_sqnorm(grads, out=None):
    """Squared L2 norm of a list of gradient tensors, as a device double (no host sync)."""
    dev = grads[0].device
    if out is None:
        out = torch.zeros(1, device=dev, dtype=torch.float64)
    ent = [(None, ptr(g), None, None, g.numel(), 1.0, 1.0) for g in grads]
    for tb, n in _tables(ent):
        call("mggan_grad_sqnorm", tb, n, ptr(out))
    return out


def clip_adamw(params, grads, exp_avg, exp_avg_sq, steps, sqnorm, max_norm, lr, beta1, beta2, eps, wd,
               grad_scale=1.0, dyn=None):
    """One fused clip + AdamW update over parallel lists; `steps` are the per-tensor step counts
    AFTER this update (bias corrections are computed from them).  dyn: per-table device pointers to
    [lr, bc1[64], bc2_sqrt[64]] that override the by-value scalars (CUDA-graph replay)."""
    ent = []
    for p, g, m, v, t in zip(params, grads, exp_avg, exp_avg_sq, steps):
        ent.append((ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), 1.0 - beta1 ** t, math.sqrt(1.0 - beta2 ** t)))
    for tb, n in _tables(ent, dyn):
        call("mggan_clip_adamw", tb, n, ptr(sqnorm), float(max_norm), float(grad_scale), float(lr), float(beta1),
             float(beta2), float(eps), float(wd))


def multi_copy(dsts, srcs):
    ent = [(ptr(d), ptr(s), None, None, d.numel(), 1.0, 1.0) for d, s in zip(dsts, srcs)]
    for tb, n in _tables(ent):
        call("mggan_multi_copy", tb, n)


# --------------------------------------------------------------------------- data side / evaluation (no autograd)
def scene_crop(atlas, img_off, img_wh, img_scale, agent_img, last_xy, out=None):
    """(N, 4, 33, 33) crop features cut from the resident u8 scene-image atlas (mggan_scene_crop)."""
    n = int(agent_img.numel())
    last_xy = _f32(last_xy)
    assert last_xy.shape == (n, 2) and agent_img.dtype == torch.int32 and atlas.dtype == torch.uint8
    if out is None:
        out = torch.empty(n, 4, 33, 33, device=last_xy.device, dtype=torch.float32)
    assert out.shape == (n, 4, 33, 33) and out.dtype == torch.float32 and out.is_contiguous()
    call("mggan_scene_crop", ptr(atlas), ptr(img_off), ptr(img_wh), ptr(img_scale), int(img_scale.numel()),
         ptr(agent_img.contiguous()), ptr(last_xy), n, ptr(out))
    torch.autograd.graph.increment_version(out)      # written through a raw pointer: caches keyed on `_version` must see it
    return out


def tube_inside(traj, radius, desc, man_list):
    """traj (P, T, 2) fp32, radius (T) float64, desc (n, 3) int32, man_list int32 -> (n,) bool (mggan_tube_inside)."""
    traj = _f32(traj)
    assert traj.dim() == 3 and traj.shape[2] == 2 and radius.dtype == torch.float64 and radius.numel() == traj.shape[1]
    assert desc.dtype == torch.int32 and man_list.dtype == torch.int32 and desc.dim() == 2 and desc.shape[1] == 3
    n = desc.shape[0]
    inside = torch.empty(n, device=traj.device, dtype=torch.uint8)
    if man_list.numel() == 0:           # keep a valid pointer for descriptors with empty manifolds
        man_list = torch.zeros(1, device=traj.device, dtype=torch.int32)
    call("mggan_tube_inside", ptr(traj), traj.shape[1], ptr(radius.contiguous()), ptr(desc.contiguous()), n,
         ptr(man_list.contiguous()), ptr(inside))
    return inside.bool()


def min_ade_fde(preds, gt, scene_off, scene_scale=None, mode_thresh=3.0):
    """preds (T, K, n, 2), gt (T, n, 2), scene_off (S+1) int32 -> ade, fde (S, K) float64 prefix minima over the samples,
    mode (S, K) int32 (mggan_min_ade_fde)."""
    preds, gt = _f32(preds), _f32(gt)
    T, K, n, _ = preds.shape
    assert gt.shape == (T, n, 2) and scene_off.dtype == torch.int32
    S = scene_off.numel() - 1
    ade = torch.empty(S, K, device=preds.device, dtype=torch.float64)
    fde = torch.empty_like(ade)
    mode = torch.empty(S, K, device=preds.device, dtype=torch.int32)
    call("mggan_min_ade_fde", ptr(preds), ptr(gt), T, K, n, ptr(scene_off.contiguous()), S,
         ptr(_f32(scene_scale)) if scene_scale is not None else None, float(mode_thresh), ptr(ade), ptr(fde), ptr(mode))
    return ade, fde, mode
