"""GPU parity of the flag-reachable variants whose golden vectors were frozen from the reference after this round's GPU
budget was spent (gan_type "gan", pool_type "sgan", experiment "discrete").  Same checks as tests/test_gpu_golden.py::test_training_iterations_objective_variants."""
import pytest
import torch

from test_gpu_golden import _run_iterations, injected  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R", [1, 70, 300])
def test_single_relative_decoder_forward_backward(R):
    """`RelativeDecoder.forward` on its own (one decoder, given initial hidden states) -- the call the discrete-latent
    generator builds on -- against the oracle's step-by-step restatement (common_modules.py:97-131), with gradients."""
    import mggan_oracle as O
    from mggan.model.modules.common_modules import RelativeDecoder
    from test_gpu_golden import DEV
    torch.manual_seed(R)
    dec = RelativeDecoder(pred_len=12, embedding_dim=16, h_dim=32, num_layers=1, social_feat_size=32, z_size=8, dropout=0.0,
                          inp_format="rel").to(DEV)
    xy, dxdy = torch.randn(R, 2) * 3, torch.randn(R, 2) * 0.4
    social, h0 = torch.randn(R, 32) * 0.5, torch.randn(R, 32) * 0.5
    sd = {"d." + k: v.detach().cpu().clone().requires_grad_(True) for k, v in dec.state_dict().items()}
    h0_o, soc_o = h0.clone().requires_grad_(True), social.clone().requires_grad_(True)
    want_abs, want_rel = O.relative_decoder(sd, "d", xy, dxdy, soc_o, h0_o)
    h0_g, soc_g = h0.to(DEV).requires_grad_(True), social.to(DEV).requires_grad_(True)
    got_abs, got_rel = dec(xy.to(DEV), dxdy.to(DEV), None, soc_g, (h0_g[None], None))
    assert got_abs.shape == want_abs.shape == (12, R, 2)
    for got, want, what in ((got_abs, want_abs, "abs"), (got_rel, want_rel, "rel")):
        err = float((got.detach().cpu() - want.detach()).abs().max() / want.detach().abs().max())
        assert err <= 1e-3, (what, err)
    w = torch.randn(12, R, 2)
    (want_abs * w).sum().backward()
    (got_abs * w.to(DEV)).sum().backward()
    for got, want, what in ((h0_g.grad, h0_o.grad, "d h0"), (soc_g.grad, soc_o.grad, "d social"),
                            (dec.decoder.weight_hh_l0.grad, sd["d.decoder.weight_hh_l0"].grad, "d W_hh"),
                            (dec.hidden2pos[0].weight.grad, sd["d.hidden2pos.0.weight"].grad, "d hidden2pos.0")):
        err = float((got.detach().cpu() - want).abs().max() / (want.abs().max() + 1e-12))
        assert err <= 2e-3, (what, err)


def test_training_iterations_late_variants(golden_late_variant, injected, tmp_path, monkeypatch):  # noqa: F811
    import mggan.model.modules.standard_discrete as SD
    monkeypatch.setattr(SD, "get_global_noise", injected.global_noise)     # the discrete generator draws through its own import
    m = golden_late_variant["meta"]
    assert (m["gan_type"] == "gan" or m["pool_type"] == "sgan" or m["experiment"] == "discrete"
            or m.get("l2_loss_type") == "mse")
    _run_iterations(golden_late_variant, injected, tmp_path)
