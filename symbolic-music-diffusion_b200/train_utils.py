"""Training helpers with the reference's names (utils/train_utils.py): EMAHelper, EarlyStopping, log_metrics,
report_model.  Host-side logic only; the EMA arithmetic runs in the fused clip+Adam kernel or smd_ema_update."""
from __future__ import annotations

import logging as _pylogging

import numpy as np
import torch

from . import lib as _lib

logging = _pylogging.getLogger("smd_b200")


class EarlyStopping:
    """utils/train_utils.py:26-59.  ``update(metric)`` returns ``(improved, new_state)``; a non-improving update
    raises ``should_stop`` once ``patience_count`` has already reached ``patience``."""

    def __init__(self, min_delta=0.0, patience=0, best_metric=float("inf"), patience_count=0, should_stop=False):
        self.min_delta = min_delta
        self.patience = patience
        self.best_metric = best_metric
        self.patience_count = patience_count
        self.should_stop = should_stop

    def replace(self, **kw) -> "EarlyStopping":
        d = self.state_dict()
        d.update(kw)
        return EarlyStopping(**d)

    def update(self, metric):
        import math
        if math.isinf(self.best_metric) or self.best_metric - metric > self.min_delta:
            return True, self.replace(best_metric=metric, patience_count=0)
        stop = self.patience_count >= self.patience or self.should_stop
        return False, self.replace(patience_count=self.patience_count + 1, should_stop=stop)

    def state_dict(self):
        return dict(min_delta=float(self.min_delta), patience=int(self.patience), best_metric=float(self.best_metric),
                    patience_count=int(self.patience_count), should_stop=bool(self.should_stop))


class EMAHelper:
    """utils/train_utils.py:62-78: params_ema <- params_ema * mu + params * (1 - mu), treewise."""

    def __init__(self, mu: float, params):
        self.mu = float(mu)
        self.params = params          # a ParamArena

    def update(self, model) -> "EMAHelper":
        lib = _lib.load_library()
        _lib.check(lib.smd_ema_update(self.params.flat.data_ptr(), model.arena.flat.data_ptr(),
                                      self.params.flat.numel(), self.mu, torch.cuda.current_stream().cuda_stream))
        self.params.bump()
        return self


def report_model(model) -> int:
    """utils/train_utils.py:121-131: logs every parameter tensor's shape and the total parameter count."""
    total = 0
    for name, _off, shape in model.arena.layout:
        n = int(np.prod(shape))
        total += n
        logging.info("%s %s = %d", name, tuple(shape), n)
    logging.info("Total parameters: %d", total)
    return total


def log_metrics(metrics: dict, step: int, total_steps: int, epoch=None, summary_writer=None, verbose=True) -> str:
    """utils/train_utils.py:81-118: one line per call, scalars mirrored to TensorBoard when a writer is given."""
    parts = []
    for k, v in metrics.items():
        v = float(v.item() if hasattr(v, "item") else v)
        parts.append(f"{k}={v:.6g}")
        if summary_writer is not None:
            summary_writer.add_scalar(k, v, step)
    head = f"[epoch {epoch}] " if epoch is not None else ""
    line = f"{head}step {step}/{total_steps}: " + " ".join(parts)
    if verbose:
        logging.info(line)
    return line
