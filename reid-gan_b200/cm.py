"""Drop-in for clustercontrast/models/cm.py (CM :9-33, cm :36, CM_Hard :40-72, cm_hard :75,
ClusterMemory :110-137).

  cm(inputs, indexes, features, momentum) / cm_hard(...)  -> logits (B, C); `features` is
      updated IN PLACE during backward, exactly like the reference (momentum chain in batch
      order for CM; hardest-positive = first argmin of x.f[label] for CM_Hard).
  ClusterMemory(num_features, num_samples, temp, momentum, use_hard).forward(inputs, targets)
      -> per-sample loss (B,)  (reduction="none", cm.py:135).  The module path is fully fused:
      normalise + centroid GEMM + /temp + log-softmax + NLL in two launches, backward in three,
      update in one -- no host synchronisation anywhere (the reference's CM_Hard.backward does
      256 .cpu() round trips, cm.py:66).

All arithmetic runs in libreid_b200.so on the tensors' CUDA device; CPU tensors are rejected
(the reference's forward does `.cuda()` at cm.py:125 as well).
"""
from abc import ABC

import torch
from torch import nn, autograd

from . import _lib
from ._lib import call, check, ptr, stream_ptr


def _need_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("reid_gan_b200.cm needs CUDA tensors; there is no CPU fallback")


def _momentum_value(momentum):
    return float(momentum.item()) if isinstance(momentum, torch.Tensor) else float(momentum)


def _update(xhat, targets, features, momentum, hard):
    L = _lib.lib()
    B, D = xhat.shape
    if not features.is_contiguous():
        raise ValueError("the centroid buffer must be contiguous (it is updated in place)")
    call("reid_cm_update", ptr(xhat), ptr(targets), ptr(features), B, features.shape[0], D, momentum, int(hard), None,
                           stream_ptr())


class _CMBase(autograd.Function):
    HARD = False

    @staticmethod
    def _fwd(ctx, inputs, targets, features, momentum):
        _need_cuda(inputs, targets, features)
        L = _lib.lib()
        if features.dtype != torch.float32 or not features.is_contiguous():
            raise ValueError("the centroid buffer must be a contiguous float32 tensor (it is read and updated in place)")
        ctx.features = features                      # alias, not a copy (cm.py:13)
        ctx.momentum = _momentum_value(momentum)
        x = inputs.detach().to(torch.float32).contiguous()
        t = targets.to(torch.int64).contiguous()
        ctx.save_for_backward(x, t)
        B, D = x.shape
        C = features.shape[0]
        out = torch.empty((B, C), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            call("reid_cm_logits", ptr(x), ptr(features), B, C, D, ptr(out), stream_ptr())
        return out                                   # inputs.mm(features.t())  (cm.py:16,47)

    @staticmethod
    def _bwd(ctx, grad_outputs, hard):
        L = _lib.lib()
        x, t = ctx.saved_tensors
        f = ctx.features
        B, D = x.shape
        C = f.shape[0]
        grad_inputs = None
        with torch.cuda.device(x.device):
            if ctx.needs_input_grad[0]:
                g = grad_outputs.to(torch.float32).contiguous()
                grad_inputs = torch.empty((B, D), dtype=torch.float32, device=x.device)
                # grad_outputs.mm(features) with the PRE-update centroids (cm.py:26,56)
                call("reid_cm_grad_inputs", ptr(g), ptr(f), B, C, D, ptr(grad_inputs), stream_ptr())
            _update(x, t, f, ctx.momentum, hard)     # cm.py:29-31 / 58-70
        return grad_inputs, None, None, None


class CM(_CMBase):
    @staticmethod
    def forward(ctx, inputs, targets, features, momentum):
        return _CMBase._fwd(ctx, inputs, targets, features, momentum)

    @staticmethod
    def backward(ctx, grad_outputs):
        return _CMBase._bwd(ctx, grad_outputs, False)


class CM_Hard(_CMBase):
    @staticmethod
    def forward(ctx, inputs, targets, features, momentum):
        return _CMBase._fwd(ctx, inputs, targets, features, momentum)

    @staticmethod
    def backward(ctx, grad_outputs):
        return _CMBase._bwd(ctx, grad_outputs, True)


def cm(inputs, indexes, features, momentum=0.5):
    return CM.apply(inputs, indexes, features, torch.Tensor([momentum]).to(inputs.device))


def cm_hard(inputs, indexes, features, momentum=0.5):
    return CM_Hard.apply(inputs, indexes, features, torch.Tensor([momentum]).to(inputs.device))


class _FusedClusterLoss(autograd.Function):
    """normalize -> centroid GEMM -> /temp -> cross_entropy(reduction='none'), with the
    momentum update in backward: ClusterMemory.forward (cm.py:125-135) as one autograd node."""

    @staticmethod
    def forward(ctx, inputs, targets, features, momentum, temp, hard):
        _need_cuda(inputs, targets, features)
        L = _lib.lib()
        if features.dtype != torch.float32 or not features.is_contiguous():
            raise ValueError("the centroid buffer must be a contiguous float32 tensor (it is read and updated in place)")
        x = inputs.detach().to(torch.float32).contiguous()
        t = targets.to(torch.int64).contiguous()
        B, D = x.shape
        C = features.shape[0]
        dev = x.device
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        xhat = torch.empty((B, D), dtype=torch.float32, device=dev)
        inv_norm = torch.empty(B, dtype=torch.float32, device=dev)
        z = torch.empty((B, C), dtype=torch.float32, device=dev)
        scratch = torch.empty(max(1, L.reid_cm_forward_scratch_bytes(B, C, D)), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            call("reid_cm_forward", ptr(x), ptr(t), ptr(features), B, C, D, float(temp), ptr(loss), ptr(xhat),
                                    ptr(inv_norm), ptr(z), ptr(scratch), stream_ptr())
        ctx.features = features
        ctx.momentum = float(momentum)
        ctx.temp = float(temp)
        ctx.hard = bool(hard)
        ctx.save_for_backward(t, xhat, inv_norm, z)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        L = _lib.lib()
        t, xhat, inv_norm, z = ctx.saved_tensors
        f = ctx.features
        B, D = xhat.shape
        C = f.shape[0]
        dev = xhat.device
        grad_inputs = None
        with torch.cuda.device(dev):
            if ctx.needs_input_grad[0]:
                g = grad_loss.to(torch.float32).contiguous()
                gz = torch.empty((B, C), dtype=torch.float32, device=dev)
                grad_inputs = torch.empty((B, D), dtype=torch.float32, device=dev)
                call("reid_cm_backward", ptr(g), ptr(z), ptr(t), ptr(f), ptr(xhat), ptr(inv_norm), B, C, D, ctx.temp,
                                         ptr(gz), ptr(grad_inputs), stream_ptr())
            _update(xhat, t, f, ctx.momentum, ctx.hard)
        return grad_inputs, None, None, None, None, None


class ClusterMemory(nn.Module, ABC):
    def __init__(self, num_features, num_samples, temp=0.05, momentum=0.2, use_hard=False, use_conf=False):
        super(ClusterMemory, self).__init__()
        self.num_features = num_features
        self.num_samples = num_samples

        self.momentum = momentum
        self.temp = temp
        self.use_hard = use_hard

        self.register_buffer('features', torch.zeros(num_samples, num_features))
        self.register_buffer('gan_features', torch.zeros(num_samples, num_features))

    def forward(self, inputs, targets, gan_inputs=None, conf_weight=None):
        # gan_inputs / conf_weight are accepted and unused, as in the reference (cm.py:123-132)
        dev = self.features.device
        if not self.features.is_cuda:
            raise RuntimeError("ClusterMemory must live on a CUDA device (call .cuda(), train_usl.py:189)")
        inputs = inputs.to(dev)                       # cm.py:125 `.cuda()`
        targets = targets.to(dev)
        if self.features.dtype != torch.float32 or not self.features.is_contiguous():
            self.features = self.features.to(torch.float32).contiguous()
        return _FusedClusterLoss.apply(inputs, targets, self.features, self.momentum, self.temp, self.use_hard)
