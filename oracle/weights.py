"""TEST INFRASTRUCTURE ONLY -- deterministic, key-seeded parameter values.

The reference ships no checkpoints, so every parity test fills a ``state_dict`` with
values that depend only on (key name, shape, seed).  Both the reference modules (in
``make_golden.py``) and the B200 modules (in ``tests/``) are loaded with the same
dict, so fixtures need to store inputs' seeds and outputs only -- not weights.

Zero-initialised gates of the reference (``alpha`` network_restormer_guided_arch.py:343,
NAFNet ``beta``/``gamma`` network_nafnet_guided_arch.py:213-214) are randomised on
purpose: with them at zero the guidance path is an identity and parity passes vacuously
("Zero-init gates").
"""
import hashlib
import math

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def seeded_tensor(key: str, shape, seed: int = 0) -> torch.Tensor:
    shape = tuple(shape)
    g = _gen(key, seed)
    leaf = key.rsplit(".", 1)[-1]
    n = 1
    for s in shape:
        n *= s

    def uni(lo, hi):
        return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    if leaf == "temperature":
        return uni(0.5, 1.5)
    if leaf == "alpha":
        return uni(0.2, 1.0)
    if leaf in ("beta", "gamma") and len(shape) == 4:      # NAFBlock gates [1,C,1,1]
        return uni(0.1, 1.0)
    if leaf == "gamma":                                    # LayerScale
        return uni(0.5, 1.0)
    if len(shape) == 1 or n == max(shape):                 # norm weight / bias / conv bias
        if leaf == "weight":
            return 1.0 + 0.2 * (torch.rand(shape, generator=g) - 0.5)
        return 0.1 * (torch.rand(shape, generator=g) - 0.5)
    if leaf in ("cls_token", "pos_embed", "mask_token", "class_embedding"):
        return 0.02 * torch.randn(shape, generator=g)
    # conv / linear weights: uniform with the default-init bound 1/sqrt(fan_in)
    fan_in = n // shape[0]
    bound = 1.0 / math.sqrt(fan_in)
    if "masa_enc." in key:
        # The MASA feature encoder (17-21 ReLU convs, no normalisation) feeds two arg-max searches.  With the default
        # bound its signal shrinks ~6x in variance per conv while the biases do not, so the deepest features of a
        # random-weight encoder are almost position-independent and every candidate of a search is tied within 1e-5 --
        # a property of the fixture, not of any implementation (the oracle itself then resolves the ties by fp32
        # summation noise).  Kaiming-uniform (gain sqrt 2) keeps the features discriminative, as trained weights are.
        bound = math.sqrt(6.0 / fan_in)
    return uni(-bound, bound)


def seeded_state_dict(shapes: dict, seed: int = 0) -> dict:
    """shapes: name -> shape (e.g. ``{k: v.shape for k, v in module.state_dict().items()}``)."""
    return {k: seeded_tensor(k, s, seed) for k, s in shapes.items()}


def load_seeded(module: torch.nn.Module, seed: int = 0) -> dict:
    sd = seeded_state_dict({k: v.shape for k, v in module.state_dict().items()}, seed)
    module.load_state_dict(sd, strict=True)
    return sd


def seeded_image(key: str, shape, seed: int = 0) -> torch.Tensor:
    return torch.rand(tuple(shape), generator=_gen("img:" + key, seed), dtype=torch.float32)
