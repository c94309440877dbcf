"""Builds (G, D) from the parsed flags (reference: mggan/model/model_factory.py:7-86)."""
from mggan.model.modules.discriminators import MultiDiscriminatorTrajectory
from mggan.model.modules.standard import MultiGenerator
from mggan.model.modules.standard_discrete import DiscreteLatentGenerator
from mggan.utils import count_parameters

PRED_LEN = 12          # reference model_factory.py:18
GENERATORS = {"multi_generator": MultiGenerator, "discrete": DiscreteLatentGenerator}


def construct_model(config):
    """-> (generator, discriminator).  Like the reference it also records `use_pinet` and `num_gen_parameters` on the
    config.  `--experiment discrete` selects the DiscreteLatentGenerator ablation (one decoder + a discrete latent code,
    reference model_factory.py:50-79); both experiments share the discriminator."""
    if config.experiment not in GENERATORS:
        raise ValueError("Requested model not implemented.")
    config.use_pinet = config.weighting_target != "none" and not config.unconditional
    scene_dim = getattr(config, "scene_dim", 8 * 8)          # additive flag; the reference hard-codes 8 * 8
    social = config.h_dim if config.n_social_modules > 0 else 0
    # the discrete experiment fixes the embedding width (reference :57), the default ties it to the decoder width (:27)
    embedding = 16 if config.experiment == "discrete" else config.decoder_h_dim // 2
    shared = dict(inp_format=config.inp_format, pred_len=PRED_LEN, num_gens=config.num_gens, scene_dim=scene_dim,
                  pool_type=config.pool_type)
    generator = GENERATORS[config.experiment](
        z_size=config.noise_dim, encoder_h_dim=config.h_dim, decoder_h_dim=config.decoder_h_dim, social_feat_size=social,
        embedding_dim=int(embedding), num_social_modules=config.n_social_modules, use_pinet=config.use_pinet,
        learn_prior=config.unconditional, **shared)
    discriminator = MultiDiscriminatorTrajectory(
        num_discs=5 if config.gan_type == "probgan" else 1, unbound_output=config.gan_obj in ("W", "LS"),
        h_dim=2 * config.h_dim, gan_type=config.gan_type, global_disc=config.global_disc, **shared)
    config.num_gen_parameters = count_parameters(generator)
    print("G #parameters: ", config.num_gen_parameters)
    print("D #parameters: ", count_parameters(discriminator))
    return generator, discriminator
