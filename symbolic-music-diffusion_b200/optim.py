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
                       mu: float = 0.999, engine=None):
        """grad: flat fp32 arena of gradients.  Global-norm clipping (jax clip_grads) is fused into the same pass;
        pass max_norm=inf to apply the gradient as is.

        ``engine``: the training Engine that produced ``grad`` (train_ncsn.train_step passes it).  Its bf16 shadow
        arena is refreshed by the same kernel pass and only the padded out.kernel copy is re-packed -- the path
        bench.py times -- instead of a full re-cast on the next forward.  ``ema`` is the flat EMA arena or an
        EMAHelper (whose ParamArena version is bumped so cached inference engines re-pack)."""
        from . import lib as _lib
        lr = self.learning_rate if learning_rate is None else float(learning_rate)
        lib = _lib.load_library()
        arena = self.target.arena
        ema_flat = ema.params.flat if hasattr(ema, "params") else ema
        shadow = None
        if engine is not None and engine.params is not None and engine.params.data_ptr() == arena.flat.data_ptr():
            shadow = lib.smd_shadow_arena(engine._plan)
        mn = 3.0e38 if max_norm == float("inf") else float(max_norm)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.smd_clip_adam(arena.flat.data_ptr(), grad.data_ptr(), self.grad_ema.data_ptr(),
                                     self.grad_sq_ema.data_ptr(), None if ema_flat is None else ema_flat.data_ptr(),
                                     shadow, arena.flat.numel(), lr, self.step, mn, self.beta1, self.beta2, self.eps, mu,
                                     self._scratch.data_ptr(), self.grad_norm.data_ptr(), st))
        self.step += 1
        arena.bump()
        if shadow is not None:
            _lib.check(lib.smd_pack_weights_after_adam(engine._plan, arena.flat.data_ptr(), st))
            engine._packed_tag = (id(arena.flat), arena.version)      # this engine's operands are already current
        if hasattr(ema, "params"):
            ema.params.bump()
        return self


class Adam:
    def __init__(self, learning_rate=None, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
        if weight_decay:
            raise ValueError("weight_decay is not used by the reference and not implemented")
        self.hyper = dict(learning_rate=learning_rate, beta1=beta1, beta2=beta2, eps=eps)

    def create(self, model: Model) -> Optimizer:
        return Optimizer(model, **self.hyper)
