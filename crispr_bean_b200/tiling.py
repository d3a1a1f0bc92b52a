"""Allele <- edit contraction of the tiling models as an autograd node over the CSR/CSC kernels
(`bean_allele_gather_*` / `bean_allele_scatter_*`)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_SUF = {torch.float32: "f32", torch.float64: "f64"}


class AlleleMap:
    """Device CSR (slot -> edits) + CSC (edit -> slots) of the (G, A-1, E) 0/1 assignment."""

    def __init__(self, allele_ptr, allele_edit, n_guides: int, n_alleles: int, n_edits: int, device="cuda"):
        ptr = np.asarray(allele_ptr, dtype=np.int64)
        idx = np.asarray(allele_edit, dtype=np.int64)
        assert len(ptr) == n_guides * (n_alleles - 1) + 1 and ptr[-1] == len(idx)
        slot_of = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
        order = np.argsort(idx, kind="stable")  # CSC: slots of each edit, in slot order
        edit_ptr = np.concatenate([[0], np.cumsum(np.bincount(idx, minlength=n_edits))])
        self.device = torch.device(device)
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.int32).to(self.device)
        self.allele_ptr, self.allele_edit = t(ptr), t(idx if len(idx) else np.zeros(1))
        self.edit_ptr, self.edit_slot = t(edit_ptr), t(slot_of[order] if len(idx) else np.zeros(1))
        self.n_guides, self.n_alleles, self.n_edits = n_guides, n_alleles, n_edits
        m = _lib.BeanAlleleMap()
        m.n_guides, m.n_alleles, m.n_edits, m.nnz = n_guides, n_alleles, n_edits, len(idx)
        m.allele_ptr, m.allele_edit = self.allele_ptr.data_ptr(), self.allele_edit.data_ptr()
        m.edit_ptr, m.edit_slot = self.edit_ptr.data_ptr(), self.edit_slot.data_ptr()
        self.c = m


class AlleleGather(torch.autograd.Function):
    """(mu_edits (E,), sd_edits (E,)) -> (mu_alleles (G, A), sd_alleles (G, A)), WT column = (0, 1)."""

    @staticmethod
    def forward(ctx, mu_edit, sd_edit, amap: AlleleMap):
        if not mu_edit.is_cuda:
            raise _lib.BeanError("bean_allele_gather needs CUDA tensors: there is no CPU fallback")
        lib = _lib.lib()
        dtype = mu_edit.dtype
        mu_e, sd_e = mu_edit.detach().contiguous(), sd_edit.detach().contiguous()
        mu_a = torch.empty((amap.n_guides, amap.n_alleles), dtype=dtype, device=mu_e.device)
        sd_a = torch.empty_like(mu_a)
        st = torch.cuda.current_stream(mu_e.device).cuda_stream
        _lib.check(getattr(lib, f"bean_allele_gather_{_SUF[dtype]}")(amap.c, mu_e.data_ptr(), sd_e.data_ptr(), mu_a.data_ptr(),
                                                                    sd_a.data_ptr(), st), "bean_allele_gather")
        ctx.amap = amap
        ctx.save_for_backward(sd_e, sd_a)
        return mu_a, sd_a

    @staticmethod
    def backward(ctx, g_mu, g_sd):
        sd_e, sd_a = ctx.saved_tensors
        amap, lib, dtype = ctx.amap, _lib.lib(), sd_e.dtype
        g_mu, g_sd = g_mu.contiguous(), g_sd.contiguous()
        d_mu, d_sd = torch.empty_like(sd_e), torch.empty_like(sd_e)
        st = torch.cuda.current_stream(sd_e.device).cuda_stream
        _lib.check(getattr(lib, f"bean_allele_scatter_{_SUF[dtype]}")(amap.c, sd_e.data_ptr(), sd_a.data_ptr(), g_mu.data_ptr(),
                                                                     g_sd.data_ptr(), d_mu.data_ptr(), d_sd.data_ptr(), st),
                   "bean_allele_scatter")
        return d_mu, d_sd, None


def allele_gather(mu_edit, sd_edit, amap: AlleleMap):
    return AlleleGather.apply(mu_edit, sd_edit, amap)
