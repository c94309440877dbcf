"""PiNetMultiGeneratorGAN: the three optimisation steps of one MG-GAN training iteration
(reference: mggan/model/train.py:18-662; D step :137-213, G step :23-135, PM-Network step
:578-658) plus `predict` / `get_predictions` / `check_accuracy` (:215-289).

The step functions keep the reference's signatures and order of random draws (numpy label
smoothing: fake then real, once per `get_gan_labels` call; torch noise: one vector per scene,
sample-major).  What they execute is different: the modules are sm_100a kernels, the losses and
their gradients are single kernels (`mggan_l2_scene_min`, `mggan_bce_scalar_label`,
`mggan_ce_generators`, `mggan_pm_ml_loss`), the per-generator reweighting uses the draw counts
the selection kernel already produced, and no step synchronises with the host.

On the path: gan_type in {mgan, gan}, gan_obj in {NS, MM, LS}, weighting_target in {ml, l2, endpoint, mgan, none},
every l2_loss_type (mse squares the per-step distances), num_unrolling_steps >= 0, both experiments (multi_generator, discrete) and both pooling types (sways, sgan).  What the
reference cannot run either raises NotImplementedError (gan_obj W, gan_type infogan / probgan, weighting_target
disc_scores).  All prediction strategies of `get_predict_func` (train.py:291-576) are built:
they decode only the sequences they keep.
"""
import random
import sys
import time
from functools import partial
from pathlib import Path

import numpy as np
import torch

if __package__ in (None, ""):          # `python mggan/model/train.py <flags>` (the reference's README command line)
    sys.path.insert(0, str(Path(__file__).resolve().parents[2]))

from mggan import kernels as K
from mggan.abstract_train import MultiGeneratorGAN
from mggan.evaluation import evaluate_ade_fde
from mggan.logging import Experiment
from mggan.model.config import get_parser
from mggan.model.model_factory import construct_model
from mggan.utils import (expected_sample_indices, get_gan_labels, get_global_noise, threshold_sample_indices,
                         to_numpy, uniform_sample_indices)


def _label_scalars(shape):
    """(real, fake) scalars of one `get_gan_labels` call (patchable, like the reference's import)."""
    real, fake = get_gan_labels((1,) if shape is None else (1,))
    return float(real.flatten()[0]), float(fake.flatten()[0])


class _frozen:
    """Temporarily stop gradient tracking for a module's parameters (the G step needs D's input
    gradients only; the reference computes D's weight gradients there and discards them)."""

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]

    def __enter__(self):
        for p in self.params:
            p.requires_grad_(False)

    def __exit__(self, *exc):
        for p in self.params:
            p.requires_grad_(True)


class PiNetMultiGeneratorGAN(MultiGeneratorGAN):
    def __init__(self, generator, discriminator, config, writer, dist_ctx=None):
        super().__init__(generator, discriminator, config, writer, dist_ctx)
        assert self.gan_type in ("mgan", "gan"), self.gan_type
        if config.weighting_target not in ("ml", "none", "l2", "endpoint", "mgan"):
            raise NotImplementedError("weighting_target='%s' is outside the B200 hot path" % config.weighting_target)
        if config.weighting_target == "mgan":
            assert self.gan_type == "mgan"

    # ------------------------------------------------------------------ helpers
    def _noise(self, sub_batches, num_samples=None):
        return get_global_noise(self.config.noise_dim, sub_batches, "gaussian", self.device, num_samples)

    def _global(self, value, name=None):
        """Sum of a python number over data-parallel ranks (identity on one GPU).  `name`: the quantity ("agents",
        "active"), answered from the sums the loop body prefetched without stalling the compute stream."""
        if self.dist is None:
            return value
        if name is not None:
            got = self.dist.fetched(name, value)
            if got is not None:
                return got
        return self.dist.sum_scalar(value)

    def _reduce(self):
        return None if self.dist is None else self.dist.allreduce_grads

    def _labels(self, shape):
        """(real, fake) smoothed labels of one get_gan_labels call: python floats, or device scalars while an
        iteration is captured / replayed as a CUDA graph (mggan/graph.py)."""
        if self._graph is not None:
            return self._graph.next_labels()
        return _label_scalars(shape)

    @staticmethod
    def _phi(phi, d_out, labels, gen_idx=None, counts=None, inv_denom=None):
        """Apply one GAN-objective term (abstract_train: phi_1 / phi_2 / phi_3) to discriminator outputs."""
        fn, which, sign = phi
        label = labels[0] if which == "real" else labels[1]
        d_out = d_out.contiguous()
        if torch.is_tensor(label):
            # device-resident label (graph replay): BCE is affine in the label, L(l) = L(0) + l (L(1) - L(0)), and so is
            # its gradient -- two launches of the same kernel with constant labels
            assert fn is K.bce_scalar_label, "CUDA-graph replay covers the BCE objectives (NS, MM)"
            l0 = fn(d_out, 0.0, gen_idx, counts, sign * inv_denom)
            l1 = fn(d_out, 1.0, gen_idx, counts, sign * inv_denom)
            return l0 + label * (l1 - l0)
        return fn(d_out, label, gen_idx, counts, sign * inv_denom)

    # ------------------------------------------------------------------ G step
    def generator_step(self, in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, train_metrics, loss_mask, img=None):
        cfg = self.config
        b = in_xy.size(1)
        k = cfg.num_samples
        noise = self._noise(sub_batches, k)
        gen_out, _, gen_idxs = self.G(in_xy, in_dxdy, sub_batches, noise=noise, all_gen_out=False, img=img,
                                      mask=loss_mask, num_samples=k)
        n_act = gen_idxs.shape[0]
        counts = self.G.last_selection.totals
        if self.dist is not None:
            counts = self.dist.sum_tensor(counts.clone())
        denom = self._global(n_act, "active") * k
        loss = None
        if cfg.l2_loss_type != "none":
            scenes = K.SceneIndex.get(sub_batches, self.device)
            min_l2 = K.l2_scene_min(gen_out.abs, gt_xy, scenes, 1.0 / self._global(b, "agents"),
                                    squared=cfg.l2_loss_type == "mse")
            train_metrics["train/L2_loss"].append(min_l2.detach())
            loss = cfg.l2_loss_weight * min_l2
        with _frozen(self.D):
            disc_out = self.D(in_xy, in_dxdy, gen_out.abs, gen_out.rel, sub_batches, img=img, mask=loss_mask)
        branch_out = None
        if isinstance(disc_out, tuple):
            disc_out, branch_out = disc_out
        adv_loss = self._phi(self.phi_3, disc_out, self._labels(disc_out.shape), gen_idxs, counts, 1.0 / denom)
        train_metrics["train/gen_loss"].append(adv_loss.detach())
        loss = adv_loss if loss is None else loss + adv_loss
        if self.gan_type == "mgan":
            clf = K.ce_generators(branch_out.flatten(0, 1), gen_idxs.reshape(-1), counts, 1.0 / denom)
            train_metrics["train/info_mgan_loss"].append(clf.detach())
            loss = loss + cfg.clf_loss_weight * clf
        self.G.drop_shared()                # the trunk graph is now owned by `loss`; the weights are about to change
        self.D.zero_grad()
        self.G.zero_grad()
        loss.backward()
        self.optimizerG.step(max_norm=cfg.clipping_threshold_g, reduce_fn=self._reduce())

    # ------------------------------------------------------------------ D step
    def discriminator_step(self, in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, train_metrics, loss_mask, img=None):
        cfg = self.config
        with self.D.share_observed():       # real and fake passes share the observed-input features
            train_loss = self._discriminator_loss(in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, train_metrics,
                                                  loss_mask, img)
        self.D.zero_grad()
        train_loss.backward()
        self.optimizerD.step(max_norm=cfg.clipping_threshold_d, reduce_fn=self._reduce())

    def _discriminator_loss(self, in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, train_metrics, loss_mask, img=None):
        cfg = self.config
        real_result = self.D(in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, img=img, mask=loss_mask)
        if isinstance(real_result, tuple):
            real_result = real_result[0]
        n_act = real_result.shape[0]
        denom = self._global(n_act, "active")
        real_loss = self._phi(self.phi_1, real_result, self._labels(real_result.shape), inv_denom=1.0 / denom)
        noise = self._noise(sub_batches)[None]
        self.G.share_trunk()                # the generator step that follows runs the same weights on these inputs
        with torch.no_grad():
            gen_out, _, gen_labels_gt = self.G(in_xy, in_dxdy, sub_batches, noise=noise, all_gen_out=False, img=img,
                                               num_samples=1, mask=loss_mask)
        disc_out = self.D(in_xy, in_dxdy, gen_out.abs, gen_out.rel, sub_batches, img=img, mask=loss_mask)
        train_loss = None
        if self.gan_type == "mgan":
            disc_out, branch_out = disc_out
            ce_loss = K.ce_generators(branch_out.flatten(0, 1), gen_labels_gt.flatten(), None, 1.0 / denom)
            train_metrics["train/info_mgan_disc_loss"].append(ce_loss.detach())
            train_loss = ce_loss
        fake_loss = self._phi(self.phi_2, disc_out, self._labels(disc_out.shape), inv_denom=1.0 / denom)
        train_loss = real_loss + fake_loss if train_loss is None else train_loss + real_loss + fake_loss
        train_metrics["train/discr_loss"].append((fake_loss + real_loss).detach())
        return train_loss

    # ------------------------------------------------------------------ PM-Network step
    def net_chooser_step(self, in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, metrics, mask, img):
        cfg = self.config
        if cfg.weighting_target == "none":
            return
        self.G.drop_shared()
        gen_out, net_chooser_weights, _ = self.G(in_xy, in_dxdy, sub_batches, noise=None, all_gen_out=True, img=img,
                                                 num_samples=cfg.num_expectation_samples, mask=mask)
        with torch.no_grad():
            probs = torch.softmax(net_chooser_weights, 1).mean(0)
            for i in range(probs.shape[0]):
                metrics[f"probs/Gen {i} probability"].append(probs[i])
        n_act = net_chooser_weights.shape[0]
        inv_n = 1.0 / self._global(n_act, "active")
        if cfg.weighting_target == "ml":
            loss, _ = K.pm_ml_loss(net_chooser_weights, gen_out.abs, gt_xy, cfg.sigma, cfg.pi_net_loss_weight, inv_n)
            logged = loss.detach()
        elif cfg.weighting_target in ("l2", "endpoint"):
            # reference train.py:618-624 / :641-647: the PM-Network is trained to point at the generator whose best sample
            # is closest to the ground truth (mean L2 over time, or the end point); target selection is no-grad glue,
            # the cross-entropy and its gradient are the mggan_ce_generators kernel
            with torch.no_grad():
                if cfg.weighting_target == "l2":
                    dist = torch.norm(gen_out.abs - gt_xy[:, None, None], p=2, dim=-1).mean(0)
                else:
                    dist = torch.norm(gen_out.abs[-1] - gt_xy[-1, None, None], p=2, dim=-1)
                min_idx = torch.argmin(dist.min(0)[0].transpose(0, 1), dim=1)
            logged = K.ce_generators(net_chooser_weights, min_idx, None, inv_n)
            loss = logged * cfg.pi_net_loss_weight
            logged = logged.detach()
        else:
            # "mgan", reference train.py:604-614.  As written there, D's branch logits have shape (b, 1, G) and
            # `torch.softmax(branch_out, 1)` normalises over the singleton sample axis: the "target" is all ones, and its
            # product with out_probs.log() (b, G) broadcasts to (b, b, G); after .sum(1).mean() the loss is
            # -(1/G) sum_{j,g} log p[j,g] (a sum over agents, no 1/b), minus the entropy bonus 0.9^epoch * mean_j H(p_j).
            # Reproduced as executed (the discriminator forward still runs: it advances D's BatchNorm statistics).
            with torch.no_grad():
                self.D(in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, mask=mask, img=img)
            logp = torch.log_softmax(net_chooser_weights, 1)
            loss = -logp.sum() / logp.shape[1]
            reg = (0.9 ** self.epoch) * -(logp.exp() * logp).sum() * inv_n
            loss = loss - reg
            logged = loss.detach()
            loss = loss * cfg.pi_net_loss_weight
        metrics["train/net_chooser_loss"].append(logged)
        self.optimizerG.zero_grad()
        loss.backward()                 # the gradient already carries pi_net_loss_weight
        self.optimizerG.step(reduce_fn=self._reduce())

    # ------------------------------------------------------------------ prediction / evaluation
    def get_predict_func(self, strategy: str):
        """Reference train.py:553-576 (same names, same eps thresholds)."""
        assert strategy in ("uniform_expected", "sampling", "expected", "rejection", "smart_expected",
                            "smart_sampling", "uniform_sampling"), strategy
        if strategy == "expected":
            return self.predict_expected
        if strategy == "rejection":
            return self.predict_rejection
        if strategy == "uniform_expected":
            return self.predict_uniform
        if strategy == "smart_expected":
            return partial(self.predict_uniform, eps=1.0 / self.G.n_gs)
        if strategy == "smart_sampling":
            return partial(self.predict_smart_sampling, eps=1.0 / self.G.n_gs ** 2)
        if strategy == "uniform_sampling":
            return partial(self.predict_smart_sampling, eps=0.0)
        return self.predict

    def get_predictions(self, loader, num_preds=20, strategy="sampling"):
        assert isinstance(loader.sampler, torch.utils.data.SequentialSampler)
        self.D.eval()
        self.G.eval()
        pred_func = self.get_predict_func(strategy)
        all_preds = []
        for batch in loader:
            in_xy, in_dxdy, _, _, sub_batches, img = self._to_device(batch)
            preds, _, _, _ = pred_func(in_dxdy, in_xy, sub_batches, img=img, num=num_preds)
            all_preds.append(to_numpy(preds))
        return np.concatenate(all_preds, 2)

    def check_accuracy(self, loader, vis=False, prefix="", num_k=20, predict_strategy="sampling", debug=False,
                       **kwargs):
        preds = self.get_predictions(loader, num_preds=num_k, strategy=predict_strategy)
        return evaluate_ade_fde(loader.dataset, preds, [num_k])

    def predict(self, in_dxdy, in_xy, sub_batches, img=None, num=20, noise=None, mask=None):
        """-> (abs (pred_len, num, b, 2), rel, probs (b, G) numpy, gen_idxs (b, num) numpy)"""
        self.G.eval()
        with torch.no_grad():
            preds, net_chooser_out, gen_idxs = self.G(in_xy, in_dxdy, sub_batches, noise=noise, all_gen_out=False,
                                                      img=img, num_samples=num, mask=mask)
            probs = torch.softmax(net_chooser_out, 1)
        assert preds.abs.shape[1] == num
        return preds.abs, preds.rel, to_numpy(probs), to_numpy(gen_idxs)

    def _predict_selected(self, index_fn, in_dxdy, in_xy, sub_batches, img, num, noise, mask):
        """Shared body of the index-selecting strategies.  The reference decodes every generator on num (or
        num * G) noise samples and gathers prediction j of agent i from (sample = occurrence rank of its generator
        among the agent's earlier predictions, generator idx[i, j]) (train.py:342-350, 389-403, 453-462); that
        gather is exactly the selected-sequence decode of the training path, so only num * n sequences are run."""
        self.G.eval()
        with torch.no_grad():
            if noise is not None:
                noise = noise[:num]            # occurrence ranks are < num: later noise samples are never gathered
            preds, net_chooser_out, idxs = self.G(
                in_xy, in_dxdy, sub_batches, noise=noise, all_gen_out=False, img=img, num_samples=num, mask=mask,
                gen_idxs=lambda logits: index_fn(torch.softmax(logits, 1)))
            probs = torch.softmax(net_chooser_out, 1)
        assert preds.abs.shape[1] == num
        return preds.abs, preds.rel, to_numpy(probs), to_numpy(idxs)

    def predict_expected(self, in_dxdy, in_xy, sub_batches, img=None, num=20, noise=None, mask=None):
        """Reference train.py:291-351: predictions per generator proportional to the PM-Network probabilities."""
        return self._predict_selected(partial(expected_sample_indices, num=num), in_dxdy, in_xy, sub_batches, img, num,
                                      noise, mask)

    def predict_uniform(self, in_dxdy, in_xy, sub_batches, img=None, num=20, noise=None, eps=0.0, mask=None):
        """Reference train.py:353-412 ('uniform_expected'; 'smart_expected' with eps = 1 / G)."""
        return self._predict_selected(partial(uniform_sample_indices, num=num, eps=eps), in_dxdy, in_xy, sub_batches,
                                      img, num, noise, mask)

    def predict_smart_sampling(self, in_dxdy, in_xy, sub_batches, img=None, num=20, noise=None, eps=0.0, mask=None):
        """Reference train.py:414-465 ('smart_sampling' with eps = 1 / G^2; 'uniform_sampling' with eps = 0)."""
        return self._predict_selected(partial(threshold_sample_indices, num=num, eps=eps), in_dxdy, in_xy, sub_batches,
                                      img, num, noise, mask)

    def predict_rejection(self, in_dxdy, in_xy, sub_batches, img=None, num=20, noise=None, sigma=1e-3, N=10,
                          truncation_ratio=0.7, debug=False, mask=None, eps_noise=None):
        """Reference train.py:467-551 ("no GAN's land" rejection for single-generator models): draw
        num + ceil((1 - truncation_ratio) num) samples, estimate each sample's Jacobian Frobenius norm from N
        perturbed decodes, keep the `num` with the smallest norm.  `eps_noise` (additive): the N perturbations
        (total, b, noise_dim), drawn here when None."""
        from math import ceil
        self.G.eval()
        assert self.config.num_gens == 1, "Only implemented for single generator"
        assert 0.0 < truncation_ratio <= 1.0
        b = in_xy.shape[1]
        total = num + ceil((1 - truncation_ratio) * num)
        if noise is None:
            noise = self._noise(sub_batches, total)
        with torch.no_grad():
            preds, net_chooser_out, gen_idxs = self.G(in_xy, in_dxdy, sub_batches, noise=noise, all_gen_out=True,
                                                      img=img, num_samples=total, mask=mask)
            nb = preds.abs.shape[3]
            pred_vec = preds.abs.permute(3, 1, 2, 0, 4).reshape(nb, total, -1)
            probs = to_numpy(torch.softmax(net_chooser_out, 1))
            jac = torch.zeros(nb, total, device=pred_vec.device)
            for i in range(N):
                eps_i = (torch.randn(total, b, self.config.noise_dim, device=noise.device) * sigma ** 2
                         if eps_noise is None else eps_noise[i].to(noise.device))
                preds_eps, _, _ = self.G(in_xy, in_dxdy, sub_batches, noise=noise + eps_i, all_gen_out=True, img=img,
                                         num_samples=total, mask=mask)
                pv = preds_eps.abs.permute(3, 1, 2, 0, 4).reshape(nb, total, -1)
                jac += 1 / (sigma ** 2) * ((pv - pred_vec) ** 2).sum(-1)
            jac /= N
        _, indices = torch.sort(jac, dim=1)
        ar = torch.arange(nb, device=indices.device)
        if debug:
            gen_idxs[:] = 1
            gen_idxs[ar[None], indices[:, :num]] = 0
            return preds.abs.squeeze(2), preds.rel.squeeze(2), probs, to_numpy(gen_idxs)
        batch_abs = preds.abs[:, indices[:, :num], 0, ar[:, None]].permute(0, 2, 1, 3)
        batch_rel = preds.rel[:, indices[:, :num], 0, ar[:, None]].permute(0, 2, 1, 3)
        gen_idxs = gen_idxs[ar[:, None], indices[:, :num]]
        assert batch_abs.shape[1] == num
        return batch_abs, batch_rel, probs, to_numpy(gen_idxs)

    @staticmethod
    def construct_model(config):
        return construct_model(config)


def main(argv=None):
    args = get_parser().parse_args(argv)
    torch.manual_seed(getattr(args, "seed", 42))
    np.random.seed(getattr(args, "seed", 42))
    if args.checkpoint:
        output_dir = Path(args.checkpoint)
        assert output_dir.is_dir()
        model, config = PiNetMultiGeneratorGAN.load_from_path(output_dir)
        config.gpus = True
        config.val_every = 1
    else:
        output_dir = Path(args.log_dir) / args.experiment
        output_dir.mkdir(exist_ok=True, parents=True)
        print(str(output_dir.resolve()))
        logger = Experiment(output_dir.resolve(), name=args.name, debug=args.debug,
                            version=random.randint(10 ** 10, (10 ** 11) - 1))
        G, D = construct_model(config=args)
        logger.argparse(args)
        model = PiNetMultiGeneratorGAN(G, D, args, logger)
        logger.save()
    model.train()
    return model


if __name__ == "__main__":
    main()
