"""Host logic of the prediction strategies: the vectorised index builders of mggan/utils.py against literal
transcriptions of the reference's per-agent loops (mggan/model/train.py:313-340, 375-403 of the reference)."""
import numpy as np
import torch

from mggan.utils import expected_sample_indices, get_selection_indices, threshold_sample_indices, uniform_sample_indices


def ref_expected(probs, num):
    """Reference train.py:313-340, with np.int -> int."""
    probs = probs.numpy()
    expected_num = np.round(probs * num).astype(int)
    sort_idxs = np.argsort(-expected_num, axis=-1, kind="stable")
    num_samples_missing = num - np.sum(expected_num, 1)
    filler = np.zeros_like(expected_num)
    for b, num_missing in enumerate(num_samples_missing):
        num_missing_abs = np.abs(num_missing)
        uniq, counts = np.unique(np.tile(sort_idxs[b], num_missing_abs)[:num_missing_abs], return_counts=True)
        filler[b, uniq] += np.sign(num_missing) * counts
    expected_num += filler
    assert (np.sum(expected_num, 1) == num).all()
    sample_idxs = []
    for b_idx in range(expected_num.shape[0]):
        idxs = []
        for i in range(num):
            for idx in sort_idxs[b_idx]:
                if expected_num[b_idx, idx] > 0:
                    idxs.append(idx)
                    expected_num[b_idx, idx] -= 1
        sample_idxs.append(torch.tensor(idxs[:num]))
    return torch.stack(sample_idxs, 0)


def ref_uniform(probs, num, eps):
    """Reference train.py:375-403 (index part)."""
    num_gens = probs.shape[1]
    over_thresh = probs > eps
    over_thresh_sum = torch.sum(over_thresh, 1)
    over_thresh[over_thresh_sum < 1.0] = torch.ones(over_thresh.shape[1]).bool()
    out = []
    for b, gen_selector in enumerate(over_thresh):
        sort_idxs = torch.argsort(-probs[b, gen_selector])
        out.append(torch.arange(num_gens)[gen_selector][sort_idxs].repeat(num)[:num])
    return torch.stack(out, 0)


def test_expected_indices_match_reference_loops():
    g = torch.Generator().manual_seed(0)
    for G, num, b in [(2, 5, 40), (4, 7, 64), (8, 20, 200), (8, 19, 100), (3, 1, 30), (16, 20, 50)]:
        probs = torch.softmax(torch.randn(b, G, generator=g) * 1.5, 1)
        got = expected_sample_indices(probs, num)
        assert got.dtype == torch.int64 and got.shape == (b, num)
        assert torch.equal(got, ref_expected(probs.clone(), num).to(torch.int64)), (G, num)
    # degenerate: one generator takes everything; uniform probabilities (all ties)
    one = torch.zeros(3, 4); one[:, 2] = 1.0
    assert torch.equal(expected_sample_indices(one, 6), torch.full((3, 6), 2))
    uni = torch.full((2, 4), 0.25)
    assert torch.equal(expected_sample_indices(uni, 6), ref_expected(uni.clone(), 6).to(torch.int64))


def test_uniform_indices_match_reference_loops():
    g = torch.Generator().manual_seed(1)
    for G, num in [(2, 5), (4, 7), (8, 20), (8, 3)]:
        probs = torch.softmax(torch.randn(50, G, generator=g) * 2.0, 1)
        for eps in (0.0, 1.0 / G, 0.9999):           # the last one: nobody over the threshold -> everybody
            got = uniform_sample_indices(probs, num, eps)
            assert torch.equal(got, ref_uniform(probs.clone(), num, eps)), (G, num, eps)


def test_threshold_sampling_only_draws_allowed_generators():
    torch.manual_seed(2)
    probs = torch.tensor([[0.7, 0.2, 0.05, 0.05], [0.25, 0.25, 0.25, 0.25], [0.01, 0.01, 0.97, 0.01]])
    idx = threshold_sample_indices(probs, 400, eps=0.1)
    assert idx.shape == (3, 400)
    assert set(idx[0].tolist()) == {0, 1} and set(idx[1].tolist()) == {0, 1, 2, 3} and set(idx[2].tolist()) == {2}
    assert abs((idx[0] == 0).float().mean().item() - 0.5) < 0.1          # uniform over the allowed ones
    idx = threshold_sample_indices(probs, 50, eps=0.99)                    # none over the threshold -> all allowed
    assert len(set(idx.flatten().tolist())) == 4


def test_gather_is_the_selected_decode():
    """out[t, j, i] = all[t, rank(i, j), idx[i, j], i] (train.py:342-350) is what the selection work list encodes:
    noise sample = occurrence rank, generator = idx."""
    g = torch.Generator().manual_seed(3)
    b, G, num = 5, 3, 6
    idx = torch.randint(0, G, (b, num), generator=g)
    allp = torch.randn(2, num, G, b, 2, generator=g)
    offs = get_selection_indices(idx)
    flat = allp.reshape(2, num * G, b, 2)
    ref = flat[:, idx + offs * G, torch.arange(b).unsqueeze(1)].transpose(1, 2)
    mine = torch.stack([torch.stack([allp[:, offs[i, j], idx[i, j], i] for i in range(b)], 1) for j in range(num)], 1)
    assert torch.equal(ref, mine)


def test_pool_hidden_net_host_logic_matches_oracle(monkeypatch):
    """`--pool_type sgan`: the folded first layer, pair gather and segment max of the product's PoolHiddenNet (host logic;
    `mggan_linear_*` replaced by torch here, the GPU path is checked by tests/test_gpu_ze_variants.py) against the oracle's
    literal restatement of the reference loop (social_gan.py:203-229), values and gradients."""
    import torch
    import torch.nn.functional as F
    import mggan_oracle as O
    from mggan import kernels as K
    from mggan.model.modules.social_gan import PoolHiddenNet

    def fake_linear(x, w, b=None, act=K.ACT_NONE, slope=0.0):
        y = F.linear(x, w, b)
        return torch.relu(y) if act == K.ACT_RELU else y

    monkeypatch.setattr(K, "linear", fake_linear)
    torch.manual_seed(5)
    for h_dim, emb in ((32, 16), (64, 16)):
        net = PoolHiddenNet(embedding_dim=emb, h_dim=h_dim, mlp_dim=h_dim, bottleneck_dim=h_dim)
        sse = [(0, 3), (3, 4), (4, 9), (9, 11)]
        in_xy = torch.randn(8, 11, 2) * 5
        h = torch.randn(11, h_dim, requires_grad=True)
        out = net(in_xy, None, h, sse)
        sd = {"social." + k: v for k, v in net.state_dict().items()}
        h2 = h.detach().clone().requires_grad_(True)
        ref = O.pool_hidden_net(sd, "social", in_xy[-1], h2, sse)
        assert out.shape == ref.shape == (11, h_dim)
        assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5), float((out - ref).abs().max())
        g = torch.randn_like(ref)
        out.backward(g)
        ref.backward(g)
        assert torch.allclose(h.grad, h2.grad, rtol=1e-4, atol=1e-5)


def test_discrete_latent_generator_host_logic_matches_oracle(monkeypatch):
    """`--experiment discrete`: row layouts, code gather and autograd wiring of the product's DiscreteLatentGenerator (host
    logic) with every kernel wrapper replaced by a plain torch stand-in; the real kernels are checked on the GPU
    (tests/test_gpu_ze_variants.py) against vectors frozen from the reference."""
    import types
    import torch
    import torch.nn.functional as F
    import mggan_oracle as O
    from mggan import kernels as K
    from mggan.model.modules.standard_discrete import DiscreteLatentGenerator

    def fake_linear(x, w, b=None, act=K.ACT_NONE, slope=0.0):
        y = F.linear(x, w, b)
        return torch.relu(y) if act == K.ACT_RELU else y

    def fake_lstm(x, w_emb, b_emb, w_ih, w_hh, b_ih, b_hh):
        h = x.new_zeros(x.shape[1], w_hh.shape[1])
        c = torch.zeros_like(h)
        for t in range(x.shape[0]):
            i, f, g, o = (F.linear(F.linear(x[t], w_emb, b_emb), w_ih, b_ih) + F.linear(h, w_hh, b_hh)).chunk(4, 1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
        return h

    def fake_social(xy_last, dxdy_last, h, scenes, fc0, fc2, fc4, att_w):
        sd = {"s.feature_embedder.fc.0.weight": fc0.weight, "s.feature_embedder.fc.0.bias": fc0.bias,
              "s.feature_embedder.fc.2.weight": fc2.weight, "s.feature_embedder.fc.2.bias": fc2.bias,
              "s.feature_embedder.fc.4.weight": fc4.weight, "s.feature_embedder.fc.4.bias": fc4.bias,
              "s.attention.W.weight": att_w.weight, "s.attention.W.bias": att_w.bias}
        return O.social_attention(sd, "s", xy_last, dxdy_last, h, scenes.sub_batches)

    def fake_decode(A, social, last_xy, last_dxdy, noise, wz, gw, sel, pred_len):
        g = {k: v[0] for k, v in gw.items()}
        h, c, d, xy = A, torch.zeros_like(A), last_dxdy, last_xy
        out_abs, out_rel = [], []
        for _ in range(pred_len):
            i, f, gg, o = (F.linear(d, g["wx"], g["b"]) + F.linear(h, g["whh"])).chunk(4, 1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            u = F.leaky_relu(F.linear(h, g["w1h"]) + F.linear(social, g["w1s"]) + g["b1"], 0.01)
            d = F.linear(u, g["w2"], g["b2"])
            xy = xy + d
            out_abs.append(xy)
            out_rel.append(d)
        return torch.stack(out_abs), torch.stack(out_rel)

    class FakeScenes:
        def __init__(self, sub_batches):
            self.sub_batches = [(int(a), int(b)) for a, b in sub_batches]
            self.n_agents = self.sub_batches[-1][1]

    def fake_mlp2(x, w1, b1, w2, b2, act1=K.ACT_NONE, slope1=0.0, act2=K.ACT_NONE, slope2=0.0):
        return fake_linear(fake_linear(x, w1, b1, act1, slope1), w2, b2, act2, slope2)

    monkeypatch.setattr(K, "linear", fake_linear)
    monkeypatch.setattr(K, "mlp2", fake_mlp2)
    monkeypatch.setattr(K, "lstm_encode", fake_lstm)
    monkeypatch.setattr(K, "social_attention", fake_social)
    monkeypatch.setattr(K, "decode", fake_decode)
    monkeypatch.setattr(K.SceneIndex, "get", classmethod(lambda cls, sb, dev: sb if isinstance(sb, FakeScenes) else FakeScenes(sb)))
    monkeypatch.setattr(K.Selection, "all_generators", staticmethod(lambda n, k, g, dev: types.SimpleNamespace()))

    torch.manual_seed(3)
    Gn, k = 3, 4
    net = DiscreteLatentGenerator(z_size=8, encoder_h_dim=32, decoder_h_dim=32, social_feat_size=32, num_gens=Gn, pred_len=12,
                                  embedding_dim=16, inp_format="rel", num_social_modules=1, pool_type="sways", scene_dim=0,
                                  use_pinet=True)
    sse = [(0, 3), (3, 4), (4, 6)]
    in_xy = torch.randn(8, 6, 2).cumsum(0)
    in_dxdy = in_xy[1:] - in_xy[:-1]
    noise = torch.randn(k, 6, 8)
    idx = torch.randint(0, Gn, (6, k))
    sd = {n: v for n, v in net.state_dict().items()}

    (rel, ab), logits, got_idx = net(in_xy, in_dxdy, sse, noise=noise, all_gen_out=False, num_samples=k, gen_idxs=idx)
    (orel, oab), ologits, _ = O.generator_forward(sd, Gn, in_xy, in_dxdy, sse, noise, False, None, k, None, idx)
    assert ab.shape == oab.shape == (12, k, 6, 2) and torch.equal(got_idx, idx)
    assert torch.allclose(ab, oab, rtol=1e-4, atol=1e-5) and torch.allclose(rel, orel, rtol=1e-4, atol=1e-5)
    assert torch.allclose(logits, ologits, rtol=1e-4, atol=1e-5)
    assert net.last_selection.totals.tolist() == torch.bincount(idx.flatten(), minlength=Gn).tolist()
    # gradients reach the code encoder and the trunk
    ab.square().mean().backward()
    gw = net.one_hot_sample_encoder[0].weight.grad
    params = {n: p for n, p in net.named_parameters()}
    osd = {n: (v.detach().clone().requires_grad_(True) if n in params and params[n].requires_grad else v) for n, v in sd.items()}
    (_, oab2), _, _ = O.generator_forward(osd, Gn, in_xy, in_dxdy, sse, noise, False, None, k, None, idx)
    oab2.square().mean().backward()
    assert gw is not None and torch.allclose(gw, osd["one_hot_sample_encoder.0.weight"].grad, rtol=1e-3, atol=1e-6)
    assert torch.allclose(net.encoder.embedding.weight.grad, osd["encoder.embedding.weight"].grad, rtol=1e-3, atol=1e-6)

    monkeypatch.setattr(DiscreteLatentGenerator, "get_samples", lambda self, enc_h, num_samples=5: (self.pm_logits(enc_h), idx))
    (rel, ab), logits, _ = net(in_xy, in_dxdy, sse, noise=noise, all_gen_out=True, num_samples=k)
    (orel, oab), _, _ = O.generator_forward(sd, Gn, in_xy, in_dxdy, sse, noise, True, None, k, None, idx)
    assert ab.shape == oab.shape == (12, k, Gn, 6, 2)
    assert torch.allclose(ab, oab, rtol=1e-4, atol=1e-5) and torch.allclose(rel, orel, rtol=1e-4, atol=1e-5)


def test_discriminator_sgan_pooling_host_logic_matches_oracle(monkeypatch):
    """`--pool_type sgan` in the discriminator: the replicated pooling block (one copy of the sample-0 pooling per sample,
    social_gan.py:227-228 under `seq_start_end * n_samples`) in both forward paths of the product (trainable heads, and
    the hoisted form used with frozen heads), with masked agents, against the oracle -- kernel wrappers replaced by torch."""
    import torch
    import torch.nn.functional as F
    import mggan_oracle as O
    from mggan import kernels as K
    from mggan.model.modules.discriminators import MultiDiscriminatorTrajectory

    def fake_linear(x, w, b=None, act=K.ACT_NONE, slope=0.0):
        y = F.linear(x, w, b)
        if act == K.ACT_RELU:
            return torch.relu(y)
        if act == K.ACT_LRELU:
            return F.leaky_relu(y, slope)
        if act == K.ACT_SIGMOID_EPS:
            return torch.sigmoid(y) * (1 - 2e-7) + 1e-7
        return y

    def fake_lstm(x, w_emb, b_emb, w_ih, w_hh, b_ih, b_hh):
        h = x.new_zeros(x.shape[1], w_hh.shape[1])
        c = torch.zeros_like(h)
        for t in range(x.shape[0]):
            i, f, g, o = (F.linear(F.linear(x[t], w_emb, b_emb), w_ih, b_ih) + F.linear(h, w_hh, b_hh)).chunk(4, 1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
        return h

    def fake_heads(pe, base, soc0, w1p, wd2, bd2, wg2, bg2, n, k):
        HH = wd2.shape[1]
        z = base.repeat(k, 1) + F.linear(pe, w1p)
        z = z + torch.cat([soc0, soc0.new_zeros((k - 1) * n, soc0.shape[1])], 0)
        a = F.leaky_relu(z, 0.2)
        p = torch.sigmoid(F.linear(a[:, :HH], wd2, bd2)) * (1 - 2e-7) + 1e-7
        return p.reshape(-1), (F.linear(a[:, HH:], wg2, bg2) if wg2 is not None else None)

    class FakeScenes:
        def __init__(self, sub_batches):
            self.sub_batches = [(int(a), int(b)) for a, b in sub_batches]
            self.n_agents = self.sub_batches[-1][1]

        def pair_index(self):
            ia = torch.cat([torch.arange(a, b).repeat_interleave(b - a) for a, b in self.sub_batches])
            ib = torch.cat([torch.arange(a, b).repeat(b - a) for a, b in self.sub_batches])
            return ia, ib

    def fake_mlp2(x, w1, b1, w2, b2, act1=K.ACT_NONE, slope1=0.0, act2=K.ACT_NONE, slope2=0.0):
        return fake_linear(fake_linear(x, w1, b1, act1, slope1), w2, b2, act2, slope2)

    monkeypatch.setattr(K, "linear", fake_linear)
    monkeypatch.setattr(K, "mlp2", fake_mlp2)
    monkeypatch.setattr(K, "lstm_encode", fake_lstm)
    monkeypatch.setattr(K, "disc_heads", fake_heads)
    monkeypatch.setattr(K.SceneIndex, "get", classmethod(lambda cls, sb, dev: sb if isinstance(sb, FakeScenes) else FakeScenes(sb)))

    torch.manual_seed(11)
    Gn, k, N = 3, 4, 7
    D = MultiDiscriminatorTrajectory(num_gens=Gn, num_discs=1, unbound_output=False, h_dim=64, inp_format="rel", pred_len=12,
                                     gan_type="mgan", global_disc=1, scene_dim=0, pool_type="sgan")
    sse = [(0, 3), (3, 4), (4, 7)]
    in_xy = torch.randn(8, N, 2).cumsum(0)
    in_dxdy = in_xy[1:] - in_xy[:-1]
    mask = torch.tensor([True, True, False, True, True, False, True])
    n_act = int(mask.sum())
    pred_dxdy = torch.randn(12, k, n_act, 2) * 0.3
    pred_xy = in_xy[-1, mask][None, None] + pred_dxdy.cumsum(0)
    sd = dict(D.state_dict())
    want, want_br = O.discriminator_forward(sd, in_xy, in_dxdy, pred_xy, pred_dxdy, sse, None, mask)
    got, got_br = D(in_xy, in_dxdy, pred_xy, pred_dxdy, sse, mask=mask)                 # trainable heads: plain path
    assert got.shape == want.shape == (n_act, k) and got_br.shape == want_br.shape == (n_act, k, Gn)
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-6) and torch.allclose(got_br, want_br, rtol=1e-4, atol=1e-5)
    with torch.no_grad():                                                               # frozen heads: hoisted path
        got, got_br = D(in_xy, in_dxdy, pred_xy, pred_dxdy, sse, mask=mask)
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-6) and torch.allclose(got_br, want_br, rtol=1e-4, atol=1e-5)
    # and the default attention variant keeps its sample-0-only semantics through the same two paths
    D2 = MultiDiscriminatorTrajectory(num_gens=Gn, num_discs=1, unbound_output=False, h_dim=64, inp_format="rel", pred_len=12,
                                      gan_type="mgan", global_disc=1, scene_dim=0, pool_type="sways")

    def fake_social(xy_last, dxdy_last, h, scenes, fc0, fc2, fc4, att_w):
        s = {"s.feature_embedder.fc.0.weight": fc0.weight, "s.feature_embedder.fc.0.bias": fc0.bias,
             "s.feature_embedder.fc.2.weight": fc2.weight, "s.feature_embedder.fc.2.bias": fc2.bias,
             "s.feature_embedder.fc.4.weight": fc4.weight, "s.feature_embedder.fc.4.bias": fc4.bias,
             "s.attention.W.weight": att_w.weight, "s.attention.W.bias": att_w.bias}
        return O.social_attention(s, "s", xy_last, dxdy_last, h, scenes.sub_batches)

    monkeypatch.setattr(K, "social_attention", fake_social)
    want, want_br = O.discriminator_forward(dict(D2.state_dict()), in_xy, in_dxdy, pred_xy, pred_dxdy, sse, None, mask)
    for frozen in (False, True):
        with torch.set_grad_enabled(not frozen):
            got, got_br = D2(in_xy, in_dxdy, pred_xy, pred_dxdy, sse, mask=mask)
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-6) and torch.allclose(got_br, want_br, rtol=1e-4, atol=1e-5), frozen
