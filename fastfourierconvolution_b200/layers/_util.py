"""Helpers shared by the FFC modules: parameter-holder access and activation / norm mapping."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.nn.utils.spectral_norm import SpectralNorm as _SpectralNorm
from torch.nn.utils.weight_norm import WeightNorm as _WeightNorm

from .. import ops


_counted = set()      # id(BatchNorm holder) whose num_batches_tracked was already advanced for the running forward (batched_bn_counters)


class batched_bn_counters:
    """``with batched_bn_counters(model, skip=(".lfu.",)): y = model(x)`` -- torch's ``num_batches_tracked += 1`` of every
    training-mode BatchNorm the forward will run, as ONE multi-tensor launch up front instead of one tiny launch per layer in the
    middle of the forward (10 per generator forward of fgan32).  ``skip``: name fragments of holders the forward never runs
    (the constructed-but-unused ``lfu`` of the reference, spectral_transform.py:65-67 vs :94-105), whose counters must stay."""

    def __init__(self, model: nn.Module, skip=(".lfu.",)):
        self.bns = [m for n, m in model.named_modules()
                    if isinstance(m, nn.modules.batchnorm._BatchNorm) and m.training and m.track_running_stats
                    and m.num_batches_tracked is not None and not any(f in "." + n + "." for f in skip)]

    def __enter__(self):
        if self.bns:
            torch._foreach_add_([m.num_batches_tracked for m in self.bns], 1)
            _counted.update(id(m) for m in self.bns)
        return self

    def __exit__(self, *exc):
        _counted.difference_update(id(m) for m in self.bns)
        return False


def count_batch(bn: nn.Module) -> None:
    """torch _BatchNorm.forward's bookkeeping for one holder (a no-op when batched_bn_counters already did it)."""
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None and id(bn) not in _counted:
        bn.num_batches_tracked.add_(1)


_prefetched = {}      # id(holder module) -> (normalised weight, wait): filled by prefetch_spectral_norm, consumed by effective_weight


def _standard_sn_hook(mod: nn.Module):
    for hook in mod._forward_pre_hooks.values():
        if isinstance(hook, _SpectralNorm) and hook.name == "weight" and hook.n_power_iterations == 1 and \
                (hook.dim == 0 or (hook.dim == 1 and mod.weight_orig.dim() == 4)):
            return hook
    return None


class prefetch_spectral_norm:
    """``with prefetch_spectral_norm(model, device): y = model(x)`` -- the spectral-norm power iterations of every holder
    module of ``model`` (they depend on the parameters alone) are launched up front on side streams (``ops.fork_map``), and
    ``effective_weight`` hands each layer its weight after waiting for that one computation only.  Each holder's iteration
    still runs exactly once per forward; leftovers (holders the forward did not reach) are joined on exit."""

    def __init__(self, model: nn.Module, device):
        self.mods = [m for m in model.modules() if _standard_sn_hook(m) is not None]
        self.device = device

    def __enter__(self):
        if self.device.type == "cuda" and self.mods:
            res = ops.fork_map(self.device, [(lambda m=m: _compute_sn(m)) for m in self.mods])
            for m, r in zip(self.mods, res):
                _prefetched[id(m)] = r
        return self

    def __exit__(self, *exc):
        for m in self.mods:
            r = _prefetched.pop(id(m), None)
            if r is not None:
                r[1]()
        return False


def _compute_sn(mod: nn.Module) -> torch.Tensor:
    hook = _standard_sn_hook(mod)
    return ops.spectral_norm_weight(mod.weight_orig, mod.weight_u, mod.weight_v, mod.training, hook.eps, hook.dim)


def effective_weight(mod: nn.Module) -> torch.Tensor:
    """The weight a holder module would use in its own forward.

    ``torch.nn.utils.spectral_norm`` (layers/snffc/snffc.py:23-33) installs a forward-pre-hook that
    recomputes ``mod.weight = weight_orig / sigma`` (one power iteration in training mode).  The
    B200 path never calls ``mod.forward``, so the hook's computation is done here, exactly once per forward: by
    ``ops.spectral_norm_weight`` (three kernels) for the standard configurations -- one power iteration, ``dim == 0``
    (nn.Conv2d) or ``dim == 1`` (nn.ConvTranspose2d) -- and by running the hook itself otherwise.
    """
    pre = _prefetched.pop(id(mod), None)
    if pre is not None:                               # computed on a side stream by prefetch_spectral_norm
        pre[1]()
        setattr(mod, "weight", pre[0])
        return pre[0]
    for hook in mod._forward_pre_hooks.values():
        if isinstance(hook, _SpectralNorm) and hook.name == "weight" and hook.n_power_iterations == 1 and \
                (hook.dim == 0 or (hook.dim == 1 and mod.weight_orig.dim() == 4)):
            w = ops.spectral_norm_weight(mod.weight_orig, mod.weight_u, mod.weight_v, mod.training, hook.eps, hook.dim)
            setattr(mod, "weight", w)                 # what the hook leaves behind for other readers of mod.weight
            return w
        if isinstance(hook, (_SpectralNorm, _WeightNorm)) and getattr(hook, "name", None) == "weight":
            hook(mod, (None,))       # a non-standard spectral norm / weight norm: let the hook itself recompute mod.weight
        # any other forward-pre-hook (user, profiler, framework) is not ours to fire: mod.forward is never called here
    return mod.weight


def act_code(act: nn.Module):
    """(code, slope) for activations the fused kernel implements, else None."""
    if isinstance(act, nn.Identity):
        return ops.ACT_IDENTITY, 0.0
    if isinstance(act, nn.LeakyReLU):
        return ops.ACT_LEAKY, float(act.negative_slope)
    if isinstance(act, nn.ReLU):
        return ops.ACT_RELU, 0.0
    if isinstance(act, nn.GELU) and getattr(act, "approximate", "none") == "none":
        return ops.ACT_GELU, 0.0
    if isinstance(act, nn.Tanh):
        return ops.ACT_TANH, 0.0
    if isinstance(act, nn.Sigmoid):
        return ops.ACT_SIGMOID, 0.0
    return None


def bn_act(x: torch.Tensor, bn: nn.Module, act_code_slope) -> torch.Tensor:
    """act(bn(x)) through the fused kernel; ``bn`` is an nn.BatchNorm2d holder or nn.Identity."""
    code, slope = act_code_slope
    if isinstance(bn, nn.Identity):
        if code == ops.ACT_IDENTITY:
            return x
        return ops.bn_act(x, norm=False, act=code, slope=slope)
    if bn.momentum is None:
        raise NotImplementedError("BatchNorm2d(momentum=None) (cumulative average) is not supported")
    use_batch_stats = bn.training or bn.running_mean is None
    count_batch(bn)                              # torch _BatchNorm.forward bookkeeping
    gamma, beta = bn.weight, bn.bias
    if gamma is None:                            # affine=False
        gamma = torch.ones(bn.num_features, device=x.device)
        beta = torch.zeros(bn.num_features, device=x.device)
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    return ops.bn_act(x, gamma, beta, rm, rv, norm=True, training=use_batch_stats, eps=bn.eps,
                      momentum=bn.momentum, act=code, slope=slope)
