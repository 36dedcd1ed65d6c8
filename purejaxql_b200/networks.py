"""Q-network parameter blocks: flax-named views over the flat per-seed float32
block that libpqn_b200's kernels consume (``pqn_net_layout``).

Reference modules: ``QNetwork``/``CNN`` purejaxql/pqn_minatar.py:24-69 and MLP
``QNetwork`` purejaxql/pqn_gymnax.py:29-58; parameter tree names per SURVEY
Appendix C (flax auto-naming), e.g. ``params["CNN_0"]["Dense_0"]["kernel"]``.
Only NORM_TYPE="layer_norm", NORM_INPUT=False (the shipped defaults of
pqn_minatar.yaml / pqn_cartpole.yaml) are built.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib

NET_CNN = 0
NET_MLP = 1


class QNetworkSpec:
    def __init__(self, kind: int, in_c: int, num_actions: int, hidden: int = 128, layers: int = 2):
        self.kind, self.in_c, self.num_actions, self.hidden, self.layers = kind, in_c, num_actions, hidden, layers
        self.desc = _lib.NetDesc(kind, in_c, hidden, layers, num_actions)
        lay = _lib.NetLayout()
        _lib.check(_lib.lib().pqn_net_layout(self.desc, lay), "pqn_net_layout")
        self.layout = lay
        self.total = int(lay.total)
        self.entries = self._entries()

    # (flax path, offset, shape, fan_in for init or None, init kind)
    def _entries(self):
        L, A = self.layout, self.num_actions
        e = []
        if self.kind == NET_CNN:
            C = self.in_c
            e += [(("BatchNorm_0", "scale"), L.bn_scale, (C,), "ones"),
                  (("BatchNorm_0", "bias"), L.bn_bias, (C,), "zeros"),
                  (("CNN_0", "Conv_0", "kernel"), L.conv_w, (3, 3, C, 16), "he"),
                  (("CNN_0", "Conv_0", "bias"), L.conv_b, (16,), "zeros"),
                  (("CNN_0", "LayerNorm_0", "scale"), L.ln0_scale, (16,), "ones"),
                  (("CNN_0", "LayerNorm_0", "bias"), L.ln0_bias, (16,), "zeros"),
                  (("CNN_0", "Dense_0", "kernel"), L.d0_w, (1024, 128), "he"),
                  (("CNN_0", "Dense_0", "bias"), L.d0_b, (128,), "zeros"),
                  (("CNN_0", "LayerNorm_1", "scale"), L.ln1_scale, (128,), "ones"),
                  (("CNN_0", "LayerNorm_1", "bias"), L.ln1_bias, (128,), "zeros"),
                  (("Dense_0", "kernel"), L.head_w, (128, A), "lecun"),
                  (("Dense_0", "bias"), L.head_b, (A,), "zeros")]
        else:
            D, H = self.in_c, self.hidden
            e += [(("BatchNorm_0", "scale"), L.bn_scale, (D,), "ones"),
                  (("BatchNorm_0", "bias"), L.bn_bias, (D,), "zeros"),
                  (("Dense_0", "kernel"), L.d0_w, (D, H), "lecun"),
                  (("Dense_0", "bias"), L.d0_b, (H,), "zeros"),
                  (("LayerNorm_0", "scale"), L.ln0_scale, (H,), "ones"),
                  (("LayerNorm_0", "bias"), L.ln0_bias, (H,), "zeros")]
            if self.layers == 2:
                e += [(("Dense_1", "kernel"), L.d1_w, (H, H), "lecun"),
                      (("Dense_1", "bias"), L.d1_b, (H,), "zeros"),
                      (("LayerNorm_1", "scale"), L.ln1_scale, (H,), "ones"),
                      (("LayerNorm_1", "bias"), L.ln1_bias, (H,), "zeros")]
            e += [((f"Dense_{self.layers}", "kernel"), L.head_w, (H, A), "lecun"),
                  ((f"Dense_{self.layers}", "bias"), L.head_b, (A,), "zeros")]
        return e

    # ------------------------------------------------------------------ #
    def unflatten(self, flat: torch.Tensor) -> dict:
        """float32[S, total] -> nested flax-style dict of [S, *shape] views."""
        tree: dict = {}
        for path, off, shape, _ in self.entries:
            n = int(np.prod(shape))
            v = flat[:, off:off + n].reshape((flat.shape[0],) + tuple(shape))
            d = tree
            for k in path[:-1]:
                d = d.setdefault(k, {})
            d[path[-1]] = v
        return tree

    def flatten(self, tree_or_flatdict: dict, S: int | None = None, device="cuda") -> torch.Tensor:
        """Nested dict, or flat dict keyed by "A/B/c" (oracle) or "A,B,c"
        (safetensors), of [S,*shape] (or unbatched [*shape]) arrays -> float32[S,total]."""
        def lookup(path):
            d = tree_or_flatdict
            for sep in ("/", ","):
                k = sep.join(path)
                if k in d:
                    return d[k]
            for k in path:
                d = d[k]
            return d
        first = torch.as_tensor(np.asarray(lookup(self.entries[0][0])))
        batched = first.dim() == 2
        if S is None:
            S = first.shape[0] if batched else 1
        flat = torch.zeros((S, self.total), dtype=torch.float32)
        for path, off, shape, _ in self.entries:
            v = torch.as_tensor(np.asarray(lookup(path)), dtype=torch.float32)
            n = int(np.prod(shape))
            flat[:, off:off + n] = v.reshape(S if batched else 1, n)
        return flat.to(device)

    def init(self, keys, device="cuda") -> torch.Tensor:
        """``network.init`` on the device (``pqn_net_init``): flax-default initialisers, deterministic in the
        per-seed key but not flax's draws (flax folds module paths into the key).
        ``keys``: int32[S,2] CUDA tensor."""
        S = keys.shape[0]
        flat = torch.empty((S, self.total), dtype=torch.float32, device=keys.device)
        _lib.check(_lib.lib().pqn_net_init(self.desc, _lib.p(keys.contiguous()), _lib.p(flat), S, _lib.stream_ptr()),
                   "pqn_net_init")
        return flat

    def init_host(self, keys_u32: np.ndarray, device="cuda") -> torch.Tensor:
        """Host (torch CPU generator) variant of the same initialisers, kept for tests.  flax-default initialisers (he_normal for Conv/Dense_0 of the CNN,
        lecun_normal elsewhere — truncated normal at +-2 sigma, variance-scaled by
        fan_in; zeros biases; ones LayerNorm/BatchNorm scales).  Deterministic in
        the per-seed key but NOT bit-identical to flax's draws (flax folds module
        paths into the key)."""
        S = keys_u32.shape[0]
        flat = torch.zeros((S, self.total), dtype=torch.float32)
        for s in range(S):
            g = torch.Generator().manual_seed(int(keys_u32[s, 0]) << 32 | int(keys_u32[s, 1]))
            for path, off, shape, kind in self.entries:
                n = int(np.prod(shape))
                if kind == "ones":
                    flat[s, off:off + n] = 1.0
                elif kind in ("he", "lecun"):
                    fan_in = int(np.prod(shape[:-1]))
                    var = (2.0 if kind == "he" else 1.0) / fan_in
                    std = math.sqrt(var) / 0.87962566103423978
                    w = torch.empty(n)
                    torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
                    flat[s, off:off + n] = w * std
        return flat.to(device)

    def flat_names(self, sep=","):
        return [sep.join(p) for p, *_ in self.entries]
