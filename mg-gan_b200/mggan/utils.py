"""Helpers shared by the hot path (reference: mggan/utils.py).

Same names, arguments and RNG-stream behaviour as the reference for the helpers its trainer
uses (`get_gan_labels` :18-25 draws from numpy, fake first then real; `get_global_noise`
:160-165 one N(0,1) vector per scene); the per-scene Python loop of the latter is replaced by
one draw of (S, dim) and a device gather, which consumes the torch generator identically.
Dead / paper-table helpers of the reference file are not reproduced.
"""
import numpy as np
import torch
from torch import nn


def get_gan_labels(shape, smoothness=0.1):
    """Scalar smoothed labels broadcast to `shape` (reference utils.py:18-25)."""
    label_fake = torch.zeros(shape) + np.random.uniform(0, smoothness)
    label_real = torch.ones(shape) * np.random.uniform(1 - smoothness, 1.0)
    return label_real, label_fake


def draw_gan_label_scalars(smoothness=0.1):
    """The two numpy draws of `get_gan_labels`, in its order, without materialising tensors."""
    fake = float(np.random.uniform(0, smoothness))
    real = float(np.random.uniform(1 - smoothness, 1.0))
    return real, fake


def to_numpy(x):
    return x.detach().cpu().numpy()


def count_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def load_hparams_from_tags_csv(tags_csv):
    import pandas as pd

    tags_df = pd.read_csv(tags_csv)
    return {row["key"]: convert(row["value"]) for row in tags_df.to_dict(orient="records")}


def get_argparse_defaults(parser):
    defaults = {}
    for action in parser._actions:
        if not action.required and action.dest != "help":
            defaults[action.dest] = action.default
    return defaults


def convert(val):
    if type(val) is str:
        if val.lower() == "true":
            return True
        if val.lower() == "false":
            return False
    for c in (int, float, str):
        try:
            return c(val)
        except (ValueError, TypeError):
            pass
    return val


def make_mlp(dim_list, activation="relu", batch_norm=False, dropout=0):
    """Parameter container with the reference's Sequential indices (utils.py:134-149).  A two
    element `dim_list` is a single Linear without activation."""
    layers = []
    if len(dim_list) > 2:
        for dim_in, dim_out in zip(dim_list[:-2], dim_list[1:-1]):
            layers.append(nn.Linear(dim_in, dim_out))
            if batch_norm:
                layers.append(nn.BatchNorm1d(dim_out))
            if activation == "relu":
                layers.append(nn.ReLU())
            elif activation == "leaky_relu":
                layers.append(nn.LeakyReLU())
            if dropout > 0:
                layers.append(nn.Dropout(p=dropout))
    layers.append(nn.Linear(dim_list[-2], dim_list[-1]))
    return nn.Sequential(*layers)


def gan_noise(shape, noise_type, device=None):
    if noise_type == "gaussian":
        return torch.randn(*shape, device=device)
    elif noise_type == "uniform":
        return torch.rand(*shape, device=device).sub_(0.5).mul_(2.0)
    raise ValueError('Unrecognized noise type "%s"' % noise_type)


_SCENE_IDS = {}


def _scene_ids(sub_batches, device):
    """Scene index of every agent (device int64), cached per batch structure: building it needs a host->device copy
    and a size-dependent repeat_interleave, neither of which may run inside a CUDA-graph capture."""
    key = (tuple((int(s), int(e)) for s, e in sub_batches), str(device))
    ids = _SCENE_IDS.get(key)
    if ids is None:
        if len(_SCENE_IDS) > 1024:        # ragged datasets: one small entry per batch structure
            _SCENE_IDS.clear()
        sizes = torch.tensor([e - s for s, e in sub_batches], device=device)
        ids = torch.repeat_interleave(torch.arange(len(sub_batches), device=device), sizes)
        _SCENE_IDS[key] = ids
    return ids


def get_global_noise(dim, sub_batches, noise_type, device=None, num_samples=None):
    """Scene-shared noise: one vector per scene repeated for its agents -> (N, dim), or
    (num_samples, N, dim) drawn sample-major like the reference loop (train.py:38-43)."""
    S = len(sub_batches)
    ids = _scene_ids(sub_batches, device)
    if num_samples is None:
        return gan_noise((S, dim), noise_type, device)[ids]
    return gan_noise((num_samples, S, dim), noise_type, device)[:, ids]


def get_selection_indices(sampled_gen_idxs):
    """Occurrence rank of every draw within its row, e.g. [1, 2, 3, 1] -> [0, 0, 0, 1]
    (reference utils.py:234-248; vectorised, no per-row loop).  The training path does this on
    the device inside `mggan_selection_build`; this host version backs the prediction helpers."""
    idx = sampled_gen_idxs
    same = idx[:, :, None] == idx[:, None, :]
    k = idx.shape[1]
    earlier = torch.tril(torch.ones(k, k, dtype=torch.bool, device=idx.device), -1)
    return (same & earlier[None]).sum(-1).to(idx.dtype)


# ------------------------------------------------------------------ prediction strategies (index builders)
def expected_sample_indices(probs, num):
    """Generator index of each of the `num` predictions of the 'expected' strategy (reference
    mggan/model/train.py:313-340): round(probs * num) predictions per generator, the rounding surplus / deficit
    spread one at a time over the generators in descending order of their expected count, then the generators
    visited round-robin in that order.  probs (b, G) -> (b, num) int64, on probs' device, no per-agent loop."""
    b, G = probs.shape
    dev = probs.device
    expected = torch.round(probs.float() * num).to(torch.int64)
    sort_idxs = torch.argsort(-expected, dim=1, stable=True)                   # numpy's argsort is stable for G <= 16
    missing = num - expected.sum(1)
    rank = torch.arange(G, device=dev)
    per_rank = torch.clamp((missing.abs()[:, None] - rank[None, :] + G - 1) // G, min=0)
    expected = expected + torch.zeros_like(expected).scatter_add_(1, sort_idxs, torch.sign(missing)[:, None] * per_rank)
    assert bool((expected.sum(1) == num).all())
    counts = expected.gather(1, sort_idxs)                                      # in visiting order
    rounds = torch.arange(num, device=dev)
    key = torch.where(rounds[None, None, :] < counts[:, :, None], rounds[None, None, :] * G + rank[None, :, None],
                      torch.full((1, 1, 1), num * G + G, device=dev, dtype=torch.int64))
    first = torch.sort(key.reshape(b, G * num), dim=1).values[:, :num]
    assert bool((first < num * G + G).all()), "fewer than `num` predictions selected"
    return sort_idxs.gather(1, first % G)


def uniform_sample_indices(probs, num, eps=0.0):
    """'uniform_expected' / 'smart_expected' (reference train.py:353-412): the generators whose PM-Network
    probability exceeds `eps` (all of them if none does), cycled in descending order of probability.
    probs (b, G) -> (b, num) int64."""
    b, G = probs.shape
    over = probs > eps
    over = over | (over.sum(1, keepdim=True) < 1)
    order = torch.argsort(torch.where(over, -probs, torch.full_like(probs, float("inf"))), dim=1, stable=True)
    n_sel = over.sum(1, keepdim=True)
    j = torch.arange(num, device=probs.device)[None, :]
    return order.gather(1, j % n_sel)


def threshold_sample_indices(probs, num, eps=0.0):
    """'smart_sampling' / 'uniform_sampling' (reference train.py:414-465): `num` uniform draws among the generators
    whose probability exceeds `eps` (all if none does).  probs (b, G) -> (b, num) int64."""
    over = (probs > eps).float()
    over[over.sum(1) < 1.0] = 1.0
    return torch.multinomial(over, num, replacement=True)
