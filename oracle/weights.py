"""Oracle (test infrastructure): deterministic synthetic weights keyed by reference state_dict names.

No pretrained checkpoints exist offline, so tests, goldens and the benchmark use
random weights of the CosyVoice2-0.5B architecture.  The reference's own initialisers depend on module
construction order and the global RNG; to make the *same* tensors reproducible on the GPU box (where the
reference tree is absent) every tensor is drawn from its own generator seeded by crc32(key) ^ seed, with a
scale rule chosen per key pattern so that activations stay O(1) through the depth of each stage.

``shapes`` maps reference state_dict key -> shape (see ``oracle.hift.param_shapes`` etc.; each is checked
against the reference module's ``state_dict()`` in tests/test_oracle_vs_reference.py).
"""
import math
import zlib
from collections import OrderedDict

import torch


def _gen(key, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def _randn(shape, key, seed):
    return torch.randn(shape, generator=_gen(key, seed), dtype=torch.float32)


def _rand(shape, key, seed):
    return torch.rand(shape, generator=_gen(key, seed), dtype=torch.float32)


def synth_tensor(key, shape, seed, gains=None):
    """One synthetic tensor.  ``gains``: optional {substring: multiplier} applied to matrix-like tensors."""
    shape = tuple(shape)
    gain = 1.0
    for pat, g in (gains or {}).items():
        if pat in key:
            gain *= g
    if key.endswith("parametrizations.weight.original0"):
        # weight-norm magnitude g: ||v|| per leading index, jittered (effective weight = g * v / ||v||)
        vkey = key[:-1] + "1"
        raise RuntimeError("original0 is derived from original1; use synth_state_dict")
    if key.endswith(".alpha"):                      # Snake alpha (activation.py:57-63): keep positive
        return 0.5 + _rand(shape, key, seed)
    if "pos_bias_" in key:
        return 0.1 * _randn(shape, key, seed)
    if len(shape) == 1:
        if key.endswith("bias"):
            return 0.05 * _randn(shape, key, seed) * gain
        return 1.0 + 0.1 * _randn(shape, key, seed)   # norm scales
    if "embed" in key and len(shape) == 2 and shape[0] > 512:   # embedding tables
        return 0.5 * _randn(shape, key, seed) * gain
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return _randn(shape, key, seed) * (gain / math.sqrt(fan_in))


def synth_state_dict(shapes, seed=1986, gains=None):
    sd = OrderedDict()
    for key, shape in shapes.items():
        if key.endswith("parametrizations.weight.original0"):
            continue
        sd[key] = synth_tensor(key, shape, seed, gains)
    for key, shape in shapes.items():
        if key.endswith("parametrizations.weight.original0"):
            v = sd[key[:-1] + "1"]
            norm = v.flatten(1).norm(dim=1).view(shape)
            gain = 1.0
            for pat, g in (gains or {}).items():
                if pat in key:
                    gain *= g
            sd[key] = norm * (1.0 + 0.1 * _randn(shape, key, seed)) * gain
    return OrderedDict((k, sd[k]) for k in shapes)


def weight_norm_effective(g, v):
    """torch.nn.utils.parametrizations.weight_norm (dim=0): w = g * v / ||v||_2 over dims != 0
    (reference: hifigan/generator.py:26-29 wraps every conv)."""
    return g * v / v.flatten(1).norm(dim=1).view(g.shape)
