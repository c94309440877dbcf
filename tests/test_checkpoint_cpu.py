"""CPU: a checkpoint written by the UNMODIFIED reference trainer (tests/golden/ref_checkpoint, made by
oracle/make_golden_checkpoint.py) loads into the B200 package's modules and optimisers: same file layout, same
state_dict keys and shapes (strict), same optimiser state layout."""
import os
from argparse import Namespace

import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_checkpoint")
VERSION_DIR = os.path.join(GOLD, "refckpt", "version_3")


def _config():
    from mggan.model.config import get_parser
    from mggan.utils import get_argparse_defaults, load_hparams_from_tags_csv
    defaults = get_argparse_defaults(get_parser())
    tags = load_hparams_from_tags_csv(os.path.join(VERSION_DIR, "meta_tags.csv"))
    assert tags["num_gens"] == 2 and tags["gan_type"] == "mgan"
    # use_pinet / num_gen_parameters are attributes construct_model adds to the config (model_factory.py:16,84), not flags
    unknown = set(tags) - set(defaults) - {"use_pinet", "num_gen_parameters"}
    assert not unknown, f"reference flags the B200 parser does not know: {sorted(unknown)}"
    defaults.update(tags)
    return Namespace(**defaults)


def test_reference_checkpoint_loads_strictly():
    from mggan.model.model_factory import construct_model
    from mggan.optim import FusedAdamW
    ck = torch.load(os.path.join(VERSION_DIR, "checkpoints", "checkpoint_best.pth"), map_location="cpu")
    assert set(ck) == {"generator", "discriminator", "gen_opt", "disc_opt"}
    cfg = _config()
    G, D = construct_model(cfg)
    assert list(G.state_dict()) == list(ck["generator"]), "generator state_dict keys / order differ from the reference's"
    assert list(D.state_dict()) == list(ck["discriminator"])
    G.load_state_dict(ck["generator"], strict=True)
    D.load_state_dict(ck["discriminator"], strict=True)
    for k, v in ck["generator"].items():
        assert torch.equal(G.state_dict()[k], v), k
    # the decoder aliases stay aliases after loading (reference: standard.py registers each decoder twice)
    assert G.state_dict()["G_1.decoder.weight_hh_l0"].data_ptr() == G.state_dict()["gs.1.decoder.weight_hh_l0"].data_ptr()
    for opt_key, mod, lr in (("gen_opt", G, cfg.g_lr), ("disc_opt", D, cfg.d_lr)):
        opt = FusedAdamW(mod.parameters(), lr=lr, betas=(cfg.beta1, 0.999))
        ref_groups = ck[opt_key]["param_groups"]
        assert [len(g["params"]) for g in ref_groups] == [len(g["params"]) for g in opt.state_dict()["param_groups"]]
        opt.load_state_dict(ck[opt_key])
        sd = opt.state_dict()
        assert set(sd["state"]) == set(ck[opt_key]["state"])
        i = next(iter(sd["state"]))
        assert set(sd["state"][i]) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][i]["step"]) >= 1
        assert sd["param_groups"][0]["betas"] == ref_groups[0]["betas"] and sd["param_groups"][0]["weight_decay"] == 0.01


def test_oracle_on_the_reference_checkpoint_reproduces_its_predictions():
    """The weights stored in the reference-written checkpoint, run through the oracle generator in eval mode with the
    fixture's noise and PM-Network draws, give the predictions the reference trainer made after saving."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import mggan_oracle as O
    ck = torch.load(os.path.join(VERSION_DIR, "checkpoints", "checkpoint_best.pth"), map_location="cpu")
    z = np.load(os.path.join(GOLD, "expected.npz"))
    sd = {k: v for k, v in ck["generator"].items() if not k.startswith("G_")}
    sse = [tuple(int(x) for x in r) for r in z["seq_start_end"]]
    with torch.no_grad():
        (_, ab), logits, _ = O.generator_forward(sd, 2, torch.from_numpy(z["in_xy"]), torch.from_numpy(z["in_dxdy"]), sse,
                                                 torch.from_numpy(z["noise"]), False, torch.from_numpy(z["features"]), 4, None,
                                                 torch.from_numpy(z["idx"]), training=False)
    assert float((ab - torch.from_numpy(z["abs"])).abs().max()) <= 1e-5 * float(np.abs(z["abs"]).max())
    assert float((torch.softmax(logits, 1) - torch.from_numpy(z["probs"])).abs().max()) <= 1e-6
