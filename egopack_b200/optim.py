"""Optimiser of the training loops: ``torch.optim.Adam(params, lr, weight_decay)`` of
``main_temporal.py:265-271`` / ``main_egopack.py:317-324`` as ONE kernel over flat buffers.

``FlatAdam`` is a ``torch.optim.Optimizer`` (LR schedulers, ``zero_grad``, ``state_dict`` work as usual) whose
parameters, ``exp_avg`` and ``exp_avg_sq`` live back to back in flat fp32 buffers -- every ``nn.Parameter`` keeps its
identity and shape, its ``.data`` becomes a view -- and whose step is ``egp_adam_step``: one pass that also writes the
bf16 copy of every updated parameter.  Those copies are what the tensor-core GEMMs of the bf16 compute mode read, so
the per-step fp32 -> bf16 weight casts disappear (``ops.weight_cache`` hands the views out while the parameter has not
been touched by anything else).  Step counter and learning rate are read from device memory, so the step is
CUDA-graph capturable as is.
"""
from __future__ import annotations

import ctypes
from typing import Iterable

import torch

from . import _lib as L
from . import ops

_CHUNK = 4096
_ALIGN = 8


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, shadow_dtype=torch.bfloat16):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.shadow_dtype = shadow_dtype
        self._flat = []
        for group in self.param_groups:
            self._flat.append(self._flatten(group))

    # -- construction --------------------------------------------------------------------------------------------
    def _flatten(self, group):
        ps = group["params"]
        if not ps:
            return None
        dev = ps[0].device
        if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in ps):
            raise ValueError("FlatAdam keeps fp32 CUDA parameters of one device per group (no CPU fallback)")
        offs, total = [0], 0
        for p in ps:
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            offs.append(total)
        flat = torch.zeros(max(total, _ALIGN), dtype=torch.float32, device=dev)
        m, v = torch.zeros_like(flat), torch.zeros_like(flat)
        shadow = torch.zeros(flat.shape, dtype=self.shadow_dtype, device=dev) if self.shadow_dtype is not None else None
        step = torch.zeros(max(len(ps), 1), dtype=torch.int64, device=dev)        # one counter per parameter, as torch keeps
        lr = torch.full((1,), float(group["lr"]), dtype=torch.float32, device=dev)
        ct, cs, cl = [], [], []
        with torch.no_grad():
            for i, p in enumerate(ps):
                n, o = p.numel(), offs[i]
                view = flat[o:o + n].view(p.shape)
                view.copy_(p.data)
                p.data = view                                   # same Parameter object, storage now inside the flat buffer
                for c0 in range(0, n, _CHUNK):
                    ct.append(i)
                    cs.append(o + c0)
                    cl.append(min(_CHUNK, n - c0))
                self.state[p] = {"step": step[i], "exp_avg": m[o:o + n].view(p.shape), "exp_avg_sq": v[o:o + n].view(p.shape)}
            if shadow is not None:
                shadow.copy_(flat)
        fl = dict(params=ps, flat=flat, m=m, v=v, shadow=shadow, step=step, lr=lr, lr_host=float(group["lr"]),
                  seg=torch.tensor(offs, dtype=torch.int64, device=dev),
                  chunk_tensor=torch.tensor(ct, dtype=torch.int32, device=dev),
                  chunk_start=torch.tensor(cs, dtype=torch.int64, device=dev),
                  chunk_len=torch.tensor(cl, dtype=torch.int32, device=dev), offs=offs, nchunks=len(ct))
        self._register_shadows(fl)
        return fl

    def _register_shadows(self, fl):
        if fl["shadow"] is None:
            return
        for i, p in enumerate(fl["params"]):
            o, n = fl["offs"][i], p.numel()
            ops.weight_cache.register_shadow(p, fl["shadow"][o:o + n].view(p.shape))

    # -- step ----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group, fl in zip(self.param_groups, self._flat):
            if fl is None:
                continue
            if float(group["lr"]) != fl["lr_host"]:             # LR scheduler moved it: refresh the device copy
                fl["lr_host"] = float(group["lr"])
                fl["lr"].fill_(fl["lr_host"])
            ps = fl["params"]
            table = (ctypes.c_void_p * len(ps))()
            keep = []
            for i, p in enumerate(ps):
                g = p.grad
                if g is None:
                    table[i] = None
                    continue
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                    keep.append(g)
                if g.is_sparse:
                    raise RuntimeError("FlatAdam does not support sparse gradients")
                table[i] = g.data_ptr()
            b1, b2 = group["betas"]
            L.call("egp_adam_step", L.ptr(fl["flat"]), L.ptr(fl["m"]), L.ptr(fl["v"]), L.ptr(fl["shadow"]), L.ptr(fl["seg"]),
                   L.ptr(fl["chunk_tensor"]), L.ptr(fl["chunk_start"]), L.ptr(fl["chunk_len"]), fl["nchunks"], table, len(ps),
                   L.ptr(fl["step"]), L.ptr(fl["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                   L.stream())
            if fl["shadow"] is not None:
                ops.weight_cache.shadows_synced(ps)
        # the parameters changed through raw pointers (no Python version counter moved): invalidate derived caches
        ops.bump_param_generation()
        return loss

    # -- state dict ------------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)                     # replaces the state tensors by copies: move them back in
        with torch.no_grad():
            for fl in self._flat:
                if fl is None:
                    continue
                for i, p in enumerate(fl["params"]):
                    st = self.state.get(p)
                    if not st:
                        continue
                    o, n = fl["offs"][i], p.numel()
                    mv, vv = fl["m"][o:o + n].view(p.shape), fl["v"][o:o + n].view(p.shape)
                    mv.copy_(st["exp_avg"])
                    vv.copy_(st["exp_avg_sq"])
                    fl["step"][i] = int(torch.as_tensor(st["step"]).reshape(-1)[0].item())
                    self.state[p] = {"step": fl["step"][i], "exp_avg": mv, "exp_avg_sq": vv}
                fl["lr_host"] = None                            # force a refresh of the device learning rate
