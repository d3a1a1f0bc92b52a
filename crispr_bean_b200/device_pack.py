"""Device records of a tensorised screen + the `BeanScreen` C struct that points at them.

The reference keeps counts as `(R, B, G)` with G fastest (data_class.py:220-228) and permutes views per
step (model.py:362, :534).  A kernel thread owns a whole guide and walks its replicates, so the screen is re-tiled
ONCE here to replicate-major rows `x[layer][r][g][b]`: the rows the 32 threads of a warp read for replicate r are
contiguous (one 128-bit load per thread at B = 4, 512 contiguous bytes per warp).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib


def quantile_thresholds(upper_bounds: torch.Tensor, lower_bounds: torch.Tensor):
    """Phi^-1 of the bin bounds in float64; +-inf where the quantile is 1 / 0 (model/utils.py:48-54)."""
    uq, lq = upper_bounds.double().cpu(), lower_bounds.double().cpu()
    tu = torch.where(uq == 1.0, torch.full_like(uq, math.inf), torch.erfinv(2 * uq.clamp(max=1 - 1e-16) - 1) * math.sqrt(2))
    tl = torch.where(lq == 0.0, torch.full_like(lq, -math.inf), torch.erfinv(2 * lq.clamp(min=1e-300) - 1) * math.sqrt(2))
    return tu.numpy().copy(), tl.numpy().copy()


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


_LOG_FACTORIAL = {}  # device -> f64 [4096] table of ln k!
_ROW_CONST = {torch.float32: "bean_row_const_f32", torch.float64: "bean_row_const_f64"}


def row_constants(x: torch.Tensor, with_xlogx: bool):
    """`bean_row_const_*` on a contiguous CUDA tensor of counts (..., B): per row lgamma(N + 1) - sum lgamma(x + 1)
    [+ sum x ln(x / max(N, 1))] and N, both float64 of shape x.shape[:-1].  One pass; asynchronous on the current stream."""
    if not x.is_cuda:
        raise _lib.BeanError("bean_row_const needs a CUDA tensor: there is no CPU fallback")
    x = x.contiguous()
    dev = x.device
    if dev not in _LOG_FACTORIAL:
        _LOG_FACTORIAL[dev] = torch.lgamma(torch.arange(1, 4097, dtype=torch.float64)).to(dev)
    table = _LOG_FACTORIAL[dev]
    rc = torch.empty(x.shape[:-1], dtype=torch.float64, device=dev)
    tot = torch.empty(x.shape[:-1], dtype=torch.float64, device=dev)
    name = _ROW_CONST[x.dtype]
    rc_code = getattr(_lib.lib(), name)(x.data_ptr(), rc.numel(), x.shape[-1], 1 if with_xlogx else 0, table.data_ptr(), table.numel(),
                                        rc.data_ptr(), tot.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc_code, name)
    return rc, tot


class DeviceScreen:
    """Device-resident, replicate-major copy of the per-step data of a *ScreenData object."""

    def __init__(self, data, device="cuda", dtype=torch.float32, use_bcmatch=True, mask_thres=10):
        self.device = torch.device(device)
        self.dtype = dtype
        self.mode = _lib.MODE_SURVIVAL if getattr(data, "is_survival", False) else _lib.MODE_SORTING
        G, R, B = data.n_guides, data.n_reps, data.n_condits
        layers = [(data.X_masked, data.size_factor, data.a0)]
        if use_bcmatch and getattr(data, "X_bcmatch_masked", None) is not None:
            layers.append((data.X_bcmatch_masked, data.size_factor_bcmatch, data.a0_bcmatch))
        self.n_guides, self.n_reps, self.n_bins, self.n_layers = G, R, B, len(layers)
        self.mask_thres = int(mask_thres)
        # (R, B, G) -> (R, G, B), layers stacked in front; uploaded as they are (asynchronously when the host
        # tensors are pinned, `data.pin_memory()`) and re-tiled on the device
        self.x = torch.stack([x.to(self.device, non_blocking=True).permute(0, 2, 1) for x, _, _ in layers]).to(dtype).contiguous()
        self.a0 = torch.stack([torch.as_tensor(a) for _, _, a in layers]).to(device=self.device, dtype=dtype).contiguous()
        self.row_mask = data.repguide_mask.to(self.device, non_blocking=True).to(torch.uint8).contiguous()  # (R, G)
        self._sf = np.ascontiguousarray(torch.stack([torch.as_tensor(s).double() for _, s, _ in layers]).numpy())
        self._smask = np.ascontiguousarray(data.sample_mask.double().numpy())
        if self.mode == _lib.MODE_SORTING:
            self._tu, self._tl = quantile_thresholds(data.upper_bounds, data.lower_bounds)
            self._tp = np.zeros(B)
        else:
            self._tu = self._tl = np.zeros(B)
            self._tp = np.ascontiguousarray(data.timepoints.double().numpy())
        # data-only part of every row's Dirichlet-Multinomial log-pmf, hoisted out of the SVI loop
        # (recomputed every step by the reference: L*(B+1) of its L*(3B+3) lgammas per row)
        self.row_const, n64 = row_constants(self.x, with_xlogx=True)  # (L, R, G) float64, one pass (csrc/bean_row_const.cu)
        self.row_weight = (n64 > self.mask_thres) & (self.row_mask != 0).unsqueeze(0)  # (L, R, G)
        self.ll_const = float((self.row_const * self.row_weight).sum())
        del n64
        s = _lib.BeanScreen()
        s.n_guides, s.n_reps, s.n_bins, s.n_layers = G, R, B, self.n_layers
        s.mode, s.mask_thres = self.mode, self.mask_thres
        s.x, s.a0, s.row_mask = self.x.data_ptr(), self.a0.data_ptr(), self.row_mask.data_ptr()
        s.row_const = self.row_const.data_ptr()
        s.size_factor, s.sample_mask = _dptr(self._sf), _dptr(self._smask)
        s.upper_thres, s.lower_thres, s.timepoints = _dptr(self._tu), _dptr(self._tl), _dptr(self._tp)
        self.c = s

    @property
    def cells(self) -> int:
        """guide x replicate x bin cells (the unit of BASELINE.json's cells/s metric)."""
        return self.n_guides * self.n_reps * self.n_bins


def pi_to_guide_major(pi: torch.Tensor) -> torch.Tensor:
    """reference `pi (R, 1, G, A)` -> kernel layout `(G, R, A)`."""
    return pi[:, 0].permute(1, 0, 2).contiguous()


def pi_from_guide_major(pi_g: torch.Tensor) -> torch.Tensor:
    return pi_g.permute(1, 0, 2).unsqueeze(1)
