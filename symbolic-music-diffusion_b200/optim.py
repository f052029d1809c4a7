"""flax.optim-shaped Adam over the flat parameter arena (train_ncsn.py:187-190, 287)."""
from __future__ import annotations

import torch

from .nn import Model


class Optimizer:
    def __init__(self, target: Model, learning_rate: float, beta1=0.9, beta2=0.999, eps=1e-8):
        self.target = target
        self.learning_rate = learning_rate
        self.beta1, self.beta2, self.eps = beta1, beta2, eps
        n = target.arena.flat.numel()
        dev = target.arena.flat.device
        self.step = 0
        self.grad_ema = torch.zeros(n, dtype=torch.float32, device=dev)      # flax Adam state names
        self.grad_sq_ema = torch.zeros(n, dtype=torch.float32, device=dev)
        self._scratch = torch.zeros(1024, dtype=torch.float32, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)

    def apply_gradient(self, grad: torch.Tensor, learning_rate=None, max_norm: float = float("inf"), ema=None,
                       mu: float = 0.999):
        """grad: flat fp32 arena of gradients.  Global-norm clipping (jax clip_grads) is fused into the same pass;
        pass max_norm=inf to apply the gradient as is."""
        lr = self.learning_rate if learning_rate is None else float(learning_rate)
        eng = self.target.engine(1)
        eng.lib  # noqa: B018  (library must be loaded; raises otherwise)
        from . import lib as _lib
        mn = 3.0e38 if max_norm == float("inf") else float(max_norm)
        _lib.check(eng.lib.smd_clip_adam(self.target.arena.flat.data_ptr(), grad.data_ptr(), self.grad_ema.data_ptr(),
                                         self.grad_sq_ema.data_ptr(), None if ema is None else ema.data_ptr(), None,
                                         self.target.arena.flat.numel(), lr, self.step, mn, self.beta1, self.beta2,
                                         self.eps, mu, self._scratch.data_ptr(), self.grad_norm.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream))
        self.step += 1
        self.target.arena.bump()
        return self


class Adam:
    def __init__(self, learning_rate=None, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
        if weight_decay:
            raise ValueError("weight_decay is not used by the reference and not implemented")
        self.hyper = dict(learning_rate=learning_rate, beta1=beta1, beta2=beta2, eps=eps)

    def create(self, model: Model) -> Optimizer:
        return Optimizer(model, **self.hyper)
