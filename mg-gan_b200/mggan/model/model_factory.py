"""Builds (G, D) from the parsed flags (reference: mggan/model/model_factory.py:7-86)."""
from mggan.model.modules.discriminators import MultiDiscriminatorTrajectory
from mggan.model.modules.standard import MultiGenerator
from mggan.model.modules.standard_discrete import DiscreteLatentGenerator
from mggan.utils import count_parameters

PRED_LEN = 12          # reference model_factory.py:18


def construct_model(config):
    """-> (generator, discriminator); also sets config.use_pinet and config.num_gen_parameters
    like the reference.  `--experiment discrete` builds the DiscreteLatentGenerator ablation (one decoder +
    a discrete latent code, reference model_factory.py:50-79)."""
    unbound_output = config.gan_obj in ["W", "LS"]
    num_discs = 5 if config.gan_type == "probgan" else 1
    config.use_pinet = config.weighting_target != "none" and not config.unconditional
    scene_dim = getattr(config, "scene_dim", 8 * 8)
    if config.experiment not in ("multi_generator", "discrete"):
        raise ValueError("Requested model not implemented.")
    discrete = config.experiment == "discrete"
    G = (DiscreteLatentGenerator if discrete else MultiGenerator)(
        z_size=config.noise_dim, inp_format=config.inp_format, encoder_h_dim=config.h_dim,
        decoder_h_dim=config.decoder_h_dim,
        social_feat_size=config.h_dim if config.n_social_modules > 0 else 0,
        embedding_dim=16 if discrete else int(config.decoder_h_dim // 2), num_gens=config.num_gens, pred_len=PRED_LEN,
        pool_type=config.pool_type, num_social_modules=config.n_social_modules, scene_dim=scene_dim,
        use_pinet=config.use_pinet, learn_prior=config.unconditional)
    D = MultiDiscriminatorTrajectory(
        num_discs=num_discs, num_gens=config.num_gens, unbound_output=unbound_output, h_dim=config.h_dim * 2,
        pred_len=PRED_LEN, inp_format=config.inp_format, gan_type=config.gan_type, scene_dim=scene_dim,
        global_disc=config.global_disc, pool_type=config.pool_type)
    print("G #parameters: ", count_parameters(G))
    print("D #parameters: ", count_parameters(D))
    config.num_gen_parameters = count_parameters(G)
    return G, D
