"""GPU parity: every sm_100a kernel (through the C ABI + its autograd wrapper) against the CPU
oracle (`oracle/mggan_oracle.py`, pinned to the reference by tests/test_oracle_golden.py) on the
same seeded inputs.  fp32 everywhere; the north-star tolerance is 1e-3 relative on outputs, the
kernel-level checks here are tighter."""
import math

import numpy as np
import pytest
import torch

import mggan_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


def rel_err(a, b):
    a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def check(a, b, tol, what):
    e = rel_err(a, b)
    assert math.isfinite(e) and e <= tol, f"{what}: relative-to-max error {e:.3e} > {tol:.1e}"
    return e


def rand_sd(shapes, seed, scale=0.3):
    g = torch.Generator().manual_seed(seed)
    return {k: (torch.randn(*s, generator=g) * scale) for k, s in shapes.items()}


@pytest.fixture(scope="module")
def K():
    from mggan import kernels
    return kernels


# ------------------------------------------------------------------------------------ fused two-layer perceptron
def _act_ref(z, act, slope):
    if act == 1:
        return torch.relu(z)
    if act == 2:
        return torch.where(z > 0, z, slope * z)
    if act == 3:
        return torch.sigmoid(z) * (1 - 2e-7) + 1e-7
    return z


@pytest.mark.parametrize("M,Kd,Hd,Od,act1,act2,frozen", [
    (300, 24, 64, 32, 2, 0, False),        # pred_encoder (discriminators.py:46-50)
    (5000, 64, 32, 32, 2, 0, False),       # in_encoder_fc (:52-56)
    (1000, 192, 96, 1, 2, 3, False),       # discriminator head (:76-85): sigmoid with the eps squash
    (333, 192, 96, 8, 2, 0, False),        # generator-id head (:103-108)
    (130, 128, 64, 3, 2, 0, False),        # heads without the scene CNN, G = 3
    (77, 128, 16, 16, 1, 1, False),        # PM-Network layers 1-2 (standard.py:99-105)
    (1, 64, 16, 16, 1, 1, False),
    (4099, 24, 64, 32, 2, 0, True),        # generator step: frozen discriminator, input gradient only
])
def test_mlp2_matches_pytorch_fp32(K, M, Kd, Hd, Od, act1, act2, frozen):
    """Plain PyTorch fp32 reference of the same two dense layers (CPU), outputs and every gradient."""
    g = torch.Generator().manual_seed(M * 7 + Kd)
    x = torch.randn(M, Kd, generator=g)
    w1, b1 = torch.randn(Hd, Kd, generator=g) * 0.2, torch.randn(Hd, generator=g) * 0.1
    w2, b2 = torch.randn(Od, Hd, generator=g) * 0.2, torch.randn(Od, generator=g) * 0.1
    dy = torch.randn(M, Od, generator=g)
    ref_in = [t.clone().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    yr = _act_ref(_act_ref(ref_in[0] @ ref_in[1].t() + ref_in[2], act1, 0.2) @ ref_in[3].t() + ref_in[4], act2, 0.2)
    (yr * dy).sum().backward()
    got_in = [t.clone().to(DEV).requires_grad_(not frozen or i == 0) for i, t in enumerate((x, w1, b1, w2, b2))]
    yg = K.mlp2(got_in[0], got_in[1], got_in[2], got_in[3], got_in[4], act1, 0.2, act2, 0.2)
    (yg * dy.to(DEV)).sum().backward()
    check(yg, yr, 2e-5, "y")
    names = ["dx", "dw1", "db1", "dw2", "db2"]
    for i, (a, b) in enumerate(zip(got_in, ref_in)):
        if frozen and i > 0:
            assert a.grad is None
        else:
            check(a.grad, b.grad, 1e-4, names[i])


def test_mlp2_falls_back_to_two_layers_for_unsupported_sizes(K):
    g = torch.Generator().manual_seed(3)
    x, w1, w2 = torch.randn(20, 65, generator=g), torch.randn(200, 65, generator=g) * 0.1, torch.randn(5, 200, generator=g) * 0.1
    y = K.mlp2(x.to(DEV), w1.to(DEV), None, w2.to(DEV), None, 1, 0.0, 0)
    check(y, torch.relu(x @ w1.t()) @ w2.t(), 2e-5, "y")


# ------------------------------------------------------------------------------------ linear
@pytest.mark.parametrize("M,Kd,Od,act", [(1, 16, 16, 0), (37, 24, 64, 2), (300, 192, 96, 2), (129, 96, 1, 3),
                                          (1000, 128, 16, 1), (65, 65, 8, 0)])
def test_linear(K, M, Kd, Od, act):
    g = torch.Generator().manual_seed(M + Kd)
    x = torch.randn(M, Kd, generator=g)
    w = torch.randn(Od, Kd, generator=g) * 0.2
    b = torch.randn(Od, generator=g) * 0.1
    dy = torch.randn(M, Od, generator=g)

    def ref(x, w, b):
        z = x @ w.t() + b
        if act == 1:
            return torch.relu(z)
        if act == 2:
            return O.lrelu(z, 0.2)
        if act == 3:
            return torch.sigmoid(z) * (1 - 2e-7) + 1e-7
        return z

    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    yr = ref(xr, wr, br)
    yr.backward(dy)
    xg, wg, bg = (t.clone().to(DEV).requires_grad_(True) for t in (x, w, b))
    yg = K.linear(xg, wg, bg, act, 0.2)
    yg.backward(dy.to(DEV))
    check(yg, yr, 2e-5, "y")
    check(xg.grad, xr.grad, 5e-5, "dx")
    check(wg.grad, wr.grad, 5e-5, "dw")
    check(bg.grad, br.grad, 5e-5, "db")


# ------------------------------------------------------------------------------------ encoder LSTM
@pytest.mark.parametrize("H,E,N", [(32, 16, 5), (32, 16, 200), (64, 64, 3), (64, 64, 97)])
def test_lstm_encoder(K, H, E, N):
    T = 7
    sd = rand_sd({"e.embedding.weight": (E, 2), "e.embedding.bias": (E,), "e.encoder.weight_ih_l0": (4 * H, E),
                  "e.encoder.weight_hh_l0": (4 * H, H), "e.encoder.bias_ih_l0": (4 * H,),
                  "e.encoder.bias_hh_l0": (4 * H,)}, seed=H + N)
    x = torch.randn(T, N, 2, generator=torch.Generator().manual_seed(1)) * 0.5
    dh = torch.randn(N, H, generator=torch.Generator().manual_seed(2))
    sr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    hr = O.trajectory_encoder(sr, "e", x)
    hr.backward(dh)
    sg = {k: v.clone().to(DEV).requires_grad_(True) for k, v in sd.items()}
    hg = K.lstm_encode(x.to(DEV), sg["e.embedding.weight"], sg["e.embedding.bias"], sg["e.encoder.weight_ih_l0"],
                       sg["e.encoder.weight_hh_l0"], sg["e.encoder.bias_ih_l0"], sg["e.encoder.bias_hh_l0"])
    hg.backward(dh.to(DEV))
    check(hg, hr, 2e-5, "h_T")
    for k in sd:
        check(sg[k].grad, sr[k].grad, 2e-4, "grad " + k)


# ------------------------------------------------------------------------------------ social attention
def _social_sd(HD, F, seed):
    return rand_sd({"s.feature_embedder.fc.0.weight": (32, 3), "s.feature_embedder.fc.0.bias": (32,),
                    "s.feature_embedder.fc.2.weight": (64, 32), "s.feature_embedder.fc.2.bias": (64,),
                    "s.feature_embedder.fc.4.weight": (F, 64), "s.feature_embedder.fc.4.bias": (F,),
                    "s.attention.W.weight": (F, HD), "s.attention.W.bias": (F,)}, seed)


@pytest.mark.parametrize("HD,sizes", [(32, [4]), (32, [1, 3, 2, 5, 1, 4]), (64, [32, 32]), (64, [2, 70, 1, 9]),
                                      (32, [33])])
def test_social_attention(K, HD, sizes):
    from mggan.model.modules.social import SocialAttention
    g = torch.Generator().manual_seed(sum(sizes) + HD)
    N = sum(sizes)
    sse, c = [], 0
    for s in sizes:
        sse.append([c, c + s])
        c += s
    xy = torch.randn(N, 2, generator=g) * 3
    dxdy = torch.randn(N, 2, generator=g) * 0.4
    h = torch.randn(N, HD, generator=g)
    dS = torch.randn(N, HD, generator=g)
    sd = _social_sd(HD, HD, seed=N)
    sr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    hr = h.clone().requires_grad_(True)
    Sr = O.social_attention(sr, "s", xy, dxdy, hr, sse, "scene")
    Sr.backward(dS)
    mod = SocialAttention(HD, HD).to(DEV)
    mod.load_state_dict({k[2:]: v for k, v in sd.items()})
    hg = h.clone().to(DEV).requires_grad_(True)
    Sg = mod(xy.to(DEV)[None], dxdy.to(DEV)[None], hg, sse)
    Sg.backward(dS.to(DEV))
    check(Sg, Sr, 5e-5, "S")
    check(hg.grad, hr.grad, 2e-4, "dh")
    for k, p in mod.named_parameters():
        check(p.grad, sr["s." + k].grad, 5e-4, "grad " + k)


# ------------------------------------------------------------------------------------ scene attention
def _patch_stats_reference(img):
    """R = sum patch patch^T, P = sum patch over (agent, pixel) in float64: the 36 taps (ci, ky, kx) of the zero-padded
    3 x 3 window (csrc/scene.cu; the quantities BatchNorm-1 of cnn.py:137-158 is derived from)."""
    x = np.pad(img.astype(np.float64), ((0, 0), (0, 0), (1, 1), (1, 1)))
    taps = np.stack([x[:, ci, ky:ky + 33, kx:kx + 33] for ci in range(4) for ky in range(3) for kx in range(3)], 0)
    taps = taps.reshape(36, -1)
    return taps @ taps.T, taps.sum(1)


@pytest.mark.parametrize("N,exact", [(1, True), (7, True), (700, True), (5, False), (333, False)])
def test_scene_patch_stats(K, N, exact):
    """Tensor-pipe Gram kernel against float64: crops cut from 8-bit images (TF32-exact operands, one MMA per tile) and
    arbitrary fp32 crops (3 x TF32 split products), with and without the row gather."""
    from mggan.cuda_ext import call, ptr
    from mggan.synthetic import make_batch
    rng = np.random.default_rng(N)
    if exact:
        img = make_batch([N], seed=N, with_img=True)["features"]
    else:
        img = rng.standard_normal((N, 4, 33, 33)).astype(np.float32) * rng.uniform(0.1, 3.0, (N, 4, 1, 1)).astype(np.float32)
    for rows in (None, rng.permutation(N)[: max(1, N // 2)].astype(np.int32)):
        sel = img if rows is None else img[rows]
        Rr, Pr = _patch_stats_reference(sel)
        buf = torch.zeros(36 * 36 + 36, device=DEV, dtype=torch.float64)
        d_img = torch.from_numpy(img).to(DEV)
        d_rows = None if rows is None else torch.from_numpy(rows).to(DEV)
        call("mggan_scene_patch_stats", ptr(d_img), ptr(d_rows), sel.shape[0], ptr(buf[:1296]), ptr(buf[1296:]))
        R, P = buf[:1296].reshape(36, 36).cpu().numpy(), buf[1296:].cpu().numpy()
        assert np.array_equal(R, R.T)
        # the tensor pipe accumulates with truncation: a one-sided error of ~2^-24 per accumulation on the all-positive diagonal
        # sums, 17 accumulations per crop with exact operands and 51 with the split products (measured 1e-6 / 3e-6)
        tol = 2e-6 if exact else 8e-6
        assert np.abs(R - Rr).max() <= tol * np.abs(Rr).max(), np.abs(R - Rr).max() / np.abs(Rr).max()
        assert np.abs(P - Pr).max() <= tol * max(np.abs(Pr).max(), np.sqrt(np.abs(Rr).max() * sel.shape[0] * 1089))


@pytest.mark.parametrize("C,N", [(16, 3), (8, 5), (16, 70)])
def test_scene_attention(K, C, N):
    from mggan.model.modules.cnn import AttentionGlobal
    from mggan.synthetic import make_batch
    torch.manual_seed(C * 100 + N)
    mod = AttentionGlobal(channels_cnn=C)
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if "BN_1.weight" in n:
                p.copy_(torch.empty_like(p).uniform_(0.5, 1.5) * torch.where(torch.rand_like(p) < 0.3, -1.0, 1.0))
            elif "BN_1.bias" in n:
                p.uniform_(-0.3, 0.3)
    sd = {"a." + k: v.clone() for k, v in mod.state_dict().items()}
    img = torch.from_numpy(make_batch([N], seed=N, with_img=True)["features"])
    dout = torch.randn(N, 64)
    sr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in sd.items()}
    outr = O.attention_global(sr, "a", img, training=True)
    outr.backward(dout)
    mod = mod.to(DEV).train()
    outg = mod(img.to(DEV))
    outg.backward(dout.to(DEV))
    check(outg, outr, 1e-4, "out")
    for k, p in mod.named_parameters():
        ref = sr["a." + k].grad
        if k.endswith("Conv_1.bias"):        # exactly-zero true gradient under train-mode BatchNorm: both are round-off
            assert p.grad.abs().max().item() < 1e-4 * max(1.0, dout.abs().sum().item())
            continue
        check(p.grad, ref, 1e-3, "grad " + k)
    for k, b in mod.named_buffers():
        check(b.float(), sr["a." + k].float(), 1e-5, "buffer " + k)
    # eval mode uses the running statistics
    mod.eval()
    with torch.no_grad():
        oe = mod(img.to(DEV))
        orf = O.attention_global({k: v.detach() for k, v in sr.items()}, "a", img, training=False)
    check(oe, orf, 1e-4, "eval out")
    # row gather == indexing the images first
    mod.train()
    rows = torch.tensor([N - 1, 0], dtype=torch.int32, device=DEV)
    with torch.no_grad():
        o1 = mod(img.to(DEV), rows)
        sd2 = {k: v.detach().clone() for k, v in sr.items()}
        o2 = O.attention_global(sd2, "a", img[[N - 1, 0]], training=True)
    check(o1, o2, 1e-4, "rows gather")


# ------------------------------------------------------------------------------------ selection + decoder
def test_selection_matches_reference_ranks(K):
    g = torch.Generator().manual_seed(5)
    for n, k, G in [(1, 1, 1), (7, 20, 8), (300, 20, 4), (65, 3, 2)]:
        idx = torch.randint(0, G, (n, k), generator=g)
        sel = K.Selection.from_indices(idx.to(DEV), G)
        offs = O.selection_indices(idx)
        agent, noise, out, tg = (t.cpu() for t in (sel.seq_agent, sel.seq_noise, sel.seq_out, sel.tile_gen))
        assert sel.totals.cpu().tolist() == torch.bincount(idx.flatten(), minlength=G).tolist()
        seen = set()
        for r in range(sel.n_tiles * 64):
            if agent[r] < 0:
                continue
            i = int(agent[r])
            j = int(out[r]) // n
            assert int(out[r]) % n == i
            gsel = int(tg[r // 64])
            assert gsel == int(idx[i, j])
            assert int(noise[r]) == int(offs[i, j]) * n + i
            seen.add((i, j))
        assert len(seen) == n * k


def _decoder_sd(G, C, Z, seed):
    shapes = {"enc_h_to_dec_h.0.weight": (32, C + Z), "enc_h_to_dec_h.0.bias": (32,)}
    for g in range(G):
        p = f"gs.{g}."
        shapes.update({p + "decoder.weight_ih_l0": (128, 16), p + "decoder.weight_hh_l0": (128, 32),
                       p + "decoder.bias_ih_l0": (128,), p + "decoder.bias_hh_l0": (128,),
                       p + "spatial_embedding.weight": (16, 2), p + "spatial_embedding.bias": (16,),
                       p + "hidden2pos.0.weight": (16, 64), p + "hidden2pos.0.bias": (16,),
                       p + "hidden2pos.2.weight": (2, 16), p + "hidden2pos.2.bias": (2,)})
    return rand_sd(shapes, seed, 0.25)


def test_tensor_core_gate_product_is_fp32_accurate(K):
    """tcgen05.mma kind::tf32 with the hi/lo operand split (decoder_tc.cu): D = A B^T against float64."""
    from mggan.cuda_ext import call, ptr
    g = torch.Generator().manual_seed(11)
    for scale in (1.0, 0.05):
        A = (torch.randn(128, 32, generator=g) * scale).to(DEV)
        B = (torch.randn(128, 32, generator=g) * 0.3).to(DEV)
        D = torch.full((128, 128), float("nan"), device=DEV)
        call("mggan_tc_selftest", ptr(A), ptr(B), ptr(D))
        ref = A.double() @ B.double().T
        check(D, ref, 2e-6, "3xTF32 product")


@pytest.mark.parametrize("variant", ["mggan_decoder_fwd_tc", "mggan_decoder_fwd"])
@pytest.mark.parametrize("G,n,k,C", [(1, 4, 3, 128), (4, 17, 20, 64), (8, 70, 20, 128)])
def test_decoder_selected_and_all(K, G, n, k, C, variant, monkeypatch):
    from mggan.model.modules.standard import MultiGenerator
    monkeypatch.setattr(K, "DECODER_FWD", variant)      # tensor-core (default) and FP32-FMA forward kernels
    Z = 8
    gen = torch.Generator().manual_seed(G * 1000 + n)
    sd = _decoder_sd(G, C, Z, seed=n)
    enc = torch.randn(n, C, generator=gen)
    social = torch.randn(n, 32, generator=gen)
    xy = torch.randn(n, 2, generator=gen) * 5
    dxdy = torch.randn(n, 2, generator=gen) * 0.4
    noise = torch.randn(k, n, Z, generator=gen)
    idx = torch.randint(0, G, (n, k), generator=gen)
    d_abs = torch.randn(12, k, n, 2, generator=gen)
    d_rel = torch.randn(12, k, n, 2, generator=gen)

    # oracle: forward_all on M samples + gather (standard.py:190-214)
    sr = {kk: v.clone().requires_grad_(True) for kk, v in sd.items()}
    er, socr = enc.clone().requires_grad_(True), social.clone().requires_grad_(True)
    offs = O.selection_indices(idx)
    M = int(offs.max()) + 1
    pabs, prel = O.forward_all(sr, G, xy, dxdy, er, noise[:M], socr)
    pabs_f, prel_f = pabs.reshape(12, M * G, n, 2), prel.reshape(12, M * G, n, 2)
    selr = idx + offs * G
    ar = torch.arange(n)[:, None]
    ra, rr = pabs_f[:, selr, ar].transpose(1, 2), prel_f[:, selr, ar].transpose(1, 2)
    ((ra * d_abs).sum() + (rr * d_rel).sum()).backward()

    mod = MultiGenerator(z_size=Z, encoder_h_dim=32, decoder_h_dim=32, social_feat_size=32, num_gens=G, pred_len=12,
                         embedding_dim=16, inp_format="rel", num_social_modules=1, pool_type="sways",
                         scene_dim=C - 64, use_pinet=True).to(DEV)
    mod.load_state_dict(sd, strict=False)
    eg, socg = enc.clone().to(DEV).requires_grad_(True), social.clone().to(DEV).requires_grad_(True)
    sel = K.Selection.from_indices(idx.to(DEV), G)
    ga, gr = mod._decode(xy.to(DEV)[None], dxdy.to(DEV)[None], eg, noise.to(DEV), socg, sel)
    ga, gr = ga.view(12, k, n, 2), gr.view(12, k, n, 2)
    check(ga, ra, 5e-5, "abs")
    check(gr, rr, 5e-5, "rel")
    ((ga * d_abs.to(DEV)).sum() + (gr * d_rel.to(DEV)).sum()).backward()
    check(eg.grad, er.grad, 3e-4, "d enc")
    check(socg.grad, socr.grad, 3e-4, "d social")
    params = dict(mod.named_parameters())
    for kk in sd:
        if sr[kk].grad is None:
            continue
        check(params[kk].grad, sr[kk].grad, 5e-4, "grad " + kk)

    # all generators, no grad
    with torch.no_grad():
        aa, arl = mod.forward_all(xy.to(DEV)[None], dxdy.to(DEV)[None], eg, noise[:2].to(DEV), socg)
        oa, orl = O.forward_all({kk: v.detach() for kk, v in sr.items()}, G, xy, dxdy, enc, noise[:2], social)
    check(aa, oa, 5e-5, "forward_all abs")
    check(arl, orl, 5e-5, "forward_all rel")


# ------------------------------------------------------------------------------------ losses
def test_losses(K):
    g = torch.Generator().manual_seed(9)
    T, k, G = 12, 20, 8
    sizes = [3, 1, 6, 4]
    n = sum(sizes)
    sse, c = [], 0
    for s in sizes:
        sse.append([c, c + s])
        c += s
    ab = torch.randn(T, k, n, 2, generator=g)
    gt = torch.randn(T, n, 2, generator=g)
    N_total = n + 3
    abr = ab.clone().requires_grad_(True)
    l2 = (abr - gt[:, None]).norm(dim=-1).sum(0)
    lr = sum(l2[:, a:e].sum(1).min() for a, e in sse) / N_total
    (lr * 1.7).backward()
    abg = ab.clone().to(DEV).requires_grad_(True)
    sc = K.SceneIndex.get(sse, DEV)
    lg = K.l2_scene_min(abg, gt.to(DEV), sc, 1.0 / N_total)
    (lg * 1.7).backward()
    check(lg, lr, 1e-5, "l2")
    check(abg.grad, abr.grad, 1e-5, "d abs")

    p = torch.rand(n, k, generator=g) * 0.98 + 0.01
    idx = torch.randint(0, G, (n, k), generator=g)
    counts = torch.bincount(idx.flatten(), minlength=G)
    pr = p.clone().requires_grad_(True)
    w = 1.0 / counts[idx].float()
    ref = (O._bce(pr, torch.full_like(pr, 0.93)) * w).mean()
    ref.backward()
    pg = p.clone().to(DEV).requires_grad_(True)
    out = K.bce_scalar_label(pg, 0.93, idx.to(DEV), counts.to(torch.int32).to(DEV))
    out.backward()
    check(out, ref, 1e-5, "bce")
    check(pg.grad, pr.grad, 1e-5, "d bce")

    z = torch.randn(n, k, G, generator=g)
    zr = z.clone().requires_grad_(True)
    ref = (torch.nn.functional.cross_entropy(zr.flatten(0, 1), idx.reshape(-1), reduction="none").reshape(n, k) * w).mean()
    ref.backward()
    zg = z.clone().to(DEV).requires_grad_(True)
    out = K.ce_generators(zg.flatten(0, 1), idx.to(DEV).reshape(-1), counts.to(torch.int32).to(DEV))
    out.backward()
    check(out, ref, 1e-5, "ce")
    check(zg.grad, zr.grad, 1e-5, "d ce")

    allp = torch.randn(T, 2, G, n, 2, generator=g)
    logits = torch.randn(n, G, generator=g)
    lr_ = logits.clone().requires_grad_(True)
    sigma = 1.3
    d = allp - gt[:, None, None]
    logp = (-(d * d) / (2 * sigma ** 2) - math.log(sigma) - math.log(math.sqrt(2 * math.pi))).sum([0, -1]).mean(0).t()
    target = torch.softmax(logp, 1)
    ref = -(target * torch.softmax(lr_, 1).log()).sum(1).mean()
    (ref * 0.7).backward()
    lgt = logits.clone().to(DEV).requires_grad_(True)
    out, tg = K.pm_ml_loss(lgt, allp.to(DEV), gt.to(DEV), sigma, 0.7)
    out.backward()
    check(out, ref, 1e-5, "pm loss")
    check(tg, target, 1e-4, "pm target")
    check(lgt.grad, lr_.grad, 1e-4, "d pm")


def test_gumbel_sampler_distribution(K):
    logits = torch.tensor([[0.0, 1.0, -1.0, 2.0]] * 64, device=DEV)
    idx = K.gumbel_sample(logits, 500, seed=123, offset=0)
    freq = torch.bincount(idx.flatten().cpu(), minlength=4).double()
    freq /= freq.sum()
    p = torch.softmax(logits[0].cpu().double(), 0)
    assert (freq - p).abs().max() < 0.01, (freq, p)
    idx2 = K.gumbel_sample(logits, 500, seed=123, offset=0)
    assert torch.equal(idx, idx2)


def test_fused_adamw_matches_oracle(K):
    g = torch.Generator().manual_seed(3)
    shapes = [(128, 32), (7,), (16, 4, 3, 3)] * 30            # 90 tensors -> two table launches
    ps = [torch.randn(*s, generator=g) for s in shapes]
    gs = [torch.randn(*s, generator=g) * 3 for s in shapes]
    names = [str(i) for i in range(len(ps))]
    sd = {n: p.clone() for n, p in zip(names, ps)}
    opt = O.AdamW(names)
    pg = [p.clone().to(DEV) for p in ps]
    m = [torch.zeros_like(p) for p in pg]
    v = [torch.zeros_like(p) for p in pg]
    for step in (1, 2, 3):
        grads = {n: (x * step).clone() for n, x in zip(names, gs)}
        O.clip_grad_norm(grads, 50.0)
        opt.step(sd, grads)
        gg = [(x * step).to(DEV) for x in gs]
        sq = K.grad_sqnorm(gg)
        K.clip_adamw(pg, gg, m, v, [step] * len(pg), sq, 50.0, 1e-3, 0.5, 0.999, 1e-8, 0.01)
    for n, p in zip(names, pg):
        check(p, sd[n], 1e-5, "param " + n)


# ------------------------------------------------------------------------------------ discriminator, frozen (hoisted heads)
@pytest.mark.parametrize("G,sizes,with_img,masked,k", [(8, [4, 6, 3], True, True, 20), (4, [1, 3, 2, 5], False, False, 5),
                                                       (2, [32, 32, 7], True, False, 3), (3, [70], True, True, 1)])
def test_discriminator_frozen_heads_match_oracle(K, G, sizes, with_img, masked, k):
    """Generator-step discriminator pass (weights frozen): `mggan_disc_heads_*` with the per-agent part of the
    first head layer hoisted must give the oracle's outputs and the oracle's gradient w.r.t. the predictions."""
    from mggan.model.modules.discriminators import MultiDiscriminatorTrajectory
    from mggan.synthetic import make_batch
    torch.manual_seed(G + k)
    D = MultiDiscriminatorTrajectory(num_gens=G, num_discs=1, unbound_output=False, h_dim=64, inp_format="rel", pred_len=12,
                                     gan_type="mgan", global_disc=1, scene_dim=64 if with_img else 0, pool_type="sways")
    sd = {kk: v.detach().clone() for kk, v in D.state_dict().items()}
    b = make_batch(sizes, seed=77 + G, with_img=with_img)
    sse = b.pop("seq_start_end")
    bt = {kk: torch.from_numpy(v) for kk, v in b.items()}
    N = bt["in_xy"].shape[1]
    g = torch.Generator().manual_seed(5)
    mask = torch.rand(N, generator=g) > 0.3 if masked else None
    n_act = int(mask.sum()) if masked else N
    rel = torch.randn(12, k, n_act, 2, generator=g) * 0.4
    ab = rel.cumsum(0)
    d_out = torch.randn(n_act, k, generator=g)
    d_br = torch.randn(n_act, k, G, generator=g)
    img = bt.get("features")

    relr = rel.clone().requires_grad_(True)
    oo, ob = O.discriminator_forward(sd, bt["in_xy"], bt["in_dxdy"], ab, relr, sse, img, mask)
    ((oo * d_out).sum() + (ob * d_br).sum()).backward()

    D = D.to(DEV).train()
    for p_ in D.parameters():
        p_.requires_grad_(False)
    relg = rel.clone().to(DEV).requires_grad_(True)
    mg = mask.to(DEV) if masked else None
    go, gb = D(bt["in_xy"].to(DEV), bt["in_dxdy"].to(DEV), ab.to(DEV), relg, sse, img=img.to(DEV) if with_img else None,
               mask=mg)
    check(go, oo, 1e-4, "disc output")
    check(gb, ob, 1e-4, "branch logits")
    ((go * d_out.to(DEV)).sum() + (gb * d_br.to(DEV)).sum()).backward()
    check(relg.grad, relr.grad, 5e-4, "d pred_dxdy")

    # the trainable path (dense-layer kernels on the materialised classifier input) agrees with the hoisted one
    for p_ in D.parameters():
        p_.requires_grad_(True)
    relt = rel.clone().to(DEV).requires_grad_(True)
    to, tb = D(bt["in_xy"].to(DEV), bt["in_dxdy"].to(DEV), ab.to(DEV), relt, sse, img=img.to(DEV) if with_img else None,
               mask=mg)
    check(to, go, 1e-5, "trainable vs hoisted output")
    check(tb, gb, 1e-5, "trainable vs hoisted branch")
    ((to * d_out.to(DEV)).sum() + (tb * d_br.to(DEV)).sum()).backward()
    check(relt.grad, relg.grad, 1e-4, "trainable vs hoisted d pred")
