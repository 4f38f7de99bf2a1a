"""TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py) — torch-CPU port of
the reference hot path, used as the timed CPU baseline (``cpu_baseline.kind ==
"port"`` and ``bench.py --impl reference``).

The reference (``/root/reference/models/TPNet.py:67-129``) is itself a short
sequence of stock ATen ops executed on the host cores; it cannot travel to the
GPU box, so this file issues the SAME ATen op sequence (index, mul,
scatter_add_, stack/cat, batched matmul, clamp, log, Linear-ReLU-Linear) on
``device='cpu'`` so that its wall time is representative of the reference's CPU
path, thread for thread.  ``tests/golden/make_golden.py`` checks it bit-for-bit
against the imported reference.  Never imported by the product.
"""
from __future__ import annotations

import math
from typing import List

import numpy as np
import torch


class CpuWalkProjection:
    """Minimal stateful holder: dense per-layer [N, d] fp32 matrices on the CPU."""

    def __init__(self, node_num: int, dim: int, num_layer: int, lam: float, beginning_time: float,
                 not_scale: bool = False, with_mlp: bool = True, seed: int = 0, device: str = 'cpu'):
        """``device='cpu'`` is the CPU baseline.  ``device='cuda:0'`` runs the same stock ATen op sequence on
        the GPU (what the unmodified reference does with ``--gpu 0``): the "stock PyTorch on the same B200"
        comparison point of bench.py, never part of the product."""
        g = torch.Generator().manual_seed(seed)
        self.device = torch.device(device)
        self.N, self.d, self.L, self.lam = node_num, dim, num_layer, float(lam)
        self.not_scale = not_scale
        self.layers: List[torch.Tensor] = [(torch.randn(node_num, dim, generator=g) / math.sqrt(dim)).to(self.device)]
        self.layers += [torch.zeros(node_num, dim, device=self.device) for _ in range(num_layer)]
        self.clock = np.float64(beginning_time)
        F = (2 * num_layer + 2) ** 2
        self.head = (torch.nn.Sequential(torch.nn.Linear(F, 4 * F), torch.nn.ReLU(), torch.nn.Linear(4 * F, F))
                     if with_mlp else torch.nn.Identity()).to(self.device)

    @torch.no_grad()
    def update(self, src: np.ndarray, dst: np.ndarray, times: np.ndarray) -> None:
        """Op-for-op with TPNet.py:74-99."""
        s = torch.from_numpy(src).to(self.device)
        t = torch.from_numpy(dst).to(self.device)
        t_last = times[-1]
        tf = torch.from_numpy(times).to(dtype=torch.float).to(self.device)
        w = torch.exp(-self.lam * (t_last - tf))[:, None]
        base = np.exp(-self.lam * (t_last - self.clock))
        for i in range(1, self.L + 1):
            self.layers[i] = self.layers[i] * np.power(base, i)
        wide = (len(src), self.d)
        for i in range(self.L, 0, -1):
            to_s = self.layers[i - 1][t] * w
            to_t = self.layers[i - 1][s] * w
            self.layers[i].scatter_add_(0, s[:, None].expand(*wide), to_s)
            self.layers[i].scatter_add_(0, t[:, None].expand(*wide), to_t)
        self.clock = np.float64(t_last)

    def gram_features(self, a_ids: np.ndarray, b_ids: np.ndarray) -> torch.Tensor:
        """Op-for-op with TPNet.py:119-128 (everything before ``self.mlp``)."""
        if self.device.type != 'cpu':                     # TPNet.py:109: indexing a device tensor with host ids
            a_ids = torch.from_numpy(np.asarray(a_ids)).to(self.device)
            b_ids = torch.from_numpy(np.asarray(b_ids)).to(self.device)
        xa = torch.stack([m[a_ids] for m in self.layers], dim=1)
        xb = torch.stack([m[b_ids] for m in self.layers], dim=1)
        x = torch.cat([xa, xb], dim=1)
        g = torch.matmul(x, x.transpose(1, 2)).reshape(len(a_ids), -1)
        if self.not_scale:
            return g
        g[g < 0] = 0
        return torch.log(g + 1.0)

    def pair_wise(self, a_ids: np.ndarray, b_ids: np.ndarray) -> torch.Tensor:
        """TPNet.py:112-129 including the trainable head."""
        return self.head(self.gram_features(a_ids, b_ids))
