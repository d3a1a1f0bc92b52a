"""Editing-rate sites (`pi` Dirichlet prior, reporter Multinomial, guide Dirichlet) of the tiling and survival programs
as one autograd node over the C-ABI kernel `bean_pi_sites_f32/f64` (include/bean_b200.h)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

_NAME = {torch.float32: "bean_pi_sites_f32", torch.float64: "bean_pi_sites_f64"}


class PiSiteData:
    """Per-engine constants of the sites: control allele counts (R, C, G, A), the replicate x guide mask (R, G), control
    timepoints (survival) and the data-only Multinomial constant sum m (lgamma(N + 1) - sum lgamma(x + 1))."""

    def __init__(self, counts: torch.Tensor, rep_guide_mask: torch.Tensor, control_time: Optional[torch.Tensor] = None,
                 mask_guide_site: bool = True):
        assert counts.dim() == 4 and counts.is_cuda, "allele_counts_control must be a CUDA (R, C, G, A) tensor"
        self.R, self.C, self.G, self.A = counts.shape
        self.counts = {counts.dtype: counts.contiguous()}
        mask = rep_guide_mask.reshape(self.R, self.G).to(counts.device)
        self.mask_u8 = mask.to(torch.uint8).contiguous()
        self.mask_guide_site = bool(mask_guide_site)
        self.control_time = None if control_time is None else [float(t) for t in control_time.detach().cpu().reshape(-1)]
        if self.control_time is not None:
            assert len(self.control_time) == self.C
        x = counts.double()
        const = torch.lgamma(x.sum(-1) + 1) - torch.lgamma(x + 1).sum(-1)  # (R, C, G)
        self.const = torch.where(mask.bool().unsqueeze(1), const, torch.zeros_like(const)).sum()

    def counts_as(self, dtype):
        if dtype not in self.counts:
            self.counts[dtype] = next(iter(self.counts.values())).to(dtype).contiguous()
        return self.counts[dtype]


class _PiSites(torch.autograd.Function):
    @staticmethod
    def forward(ctx, conc_guide, conc_model, pi, growth, data: PiSiteData, work_dtype):
        if not pi.is_cuda:
            raise _lib.BeanError("bean_pi_sites needs CUDA tensors: there is no CPU fallback")
        G, R, A = data.G, data.R, data.A
        dev = pi.device
        cast = lambda t: t.detach().to(work_dtype).contiguous()
        cg, cm, p = cast(conc_guide), cast(conc_model), cast(pi).reshape(R, G, A)
        assert cg.shape == cm.shape == (G, A), (cg.shape, cm.shape, (G, A))
        args = _lib.BeanPiSitesArgs()
        args.n_guides, args.n_reps, args.n_alleles, args.n_controls = G, R, A, data.C
        args.mask_guide_site = int(data.mask_guide_site)
        counts = data.counts_as(work_dtype)
        args.conc_guide, args.conc_model, args.pi, args.counts = cg.data_ptr(), cm.data_ptr(), p.data_ptr(), counts.data_ptr()
        args.rep_guide_mask = data.mask_u8.data_ptr()
        d_cg, d_cm, d_pi = torch.empty_like(cg), torch.empty_like(cm), torch.empty_like(p)
        partial = torch.empty((G,), dtype=torch.float64, device=dev)
        gr = d_gr = None
        if growth is not None:
            gr = cast(growth)
            assert gr.shape == (G, A) and data.control_time is not None
            d_gr = torch.empty_like(gr)
            args.growth, args.d_growth = gr.data_ptr(), d_gr.data_ptr()
            args.control_time = (C.c_double * data.C)(*data.control_time)
        args.prob_eps = float(torch.finfo(work_dtype).eps)
        args.partial, args.d_conc_guide, args.d_conc_model, args.d_pi = partial.data_ptr(), d_cg.data_ptr(), d_cm.data_ptr(), d_pi.data_ptr()
        _lib.check(getattr(_lib.lib(), _NAME[work_dtype])(args, torch.cuda.current_stream(dev).cuda_stream), _NAME[work_dtype])
        ctx.save_for_backward(d_cg, d_cm, d_pi, *([d_gr] if d_gr is not None else []))
        ctx.meta = (conc_guide.dtype, conc_model.dtype, pi.dtype, pi.shape, None if growth is None else growth.dtype)
        return (partial.sum() + data.const).to(work_dtype)

    @staticmethod
    def backward(ctx, grad_out):
        saved = ctx.saved_tensors
        dt_g, dt_m, dt_pi, pi_shape, dt_gr = ctx.meta
        g_gr = (grad_out * saved[3]).to(dt_gr) if dt_gr is not None else None
        return ((grad_out * saved[0]).to(dt_g), (grad_out * saved[1]).to(dt_m), (grad_out * saved[2]).to(dt_pi).reshape(pi_shape),
                g_gr, None, None)


def pi_sites(conc_guide, conc_model, pi, data: PiSiteData, growth=None, work_dtype=None):
    """model `pi` + `control_allele_count` log-probs minus the guide's `pi` log-prob (what these sites add to the ELBO).

    conc_* (G, A); pi (R, 1, G, A) or (R, G, A); growth (G, A) per-allele growth rates (survival) or None."""
    if work_dtype is None:  # the dtype torch's promotion gives the reference's own expression
        work_dtype = torch.promote_types(conc_guide.dtype, pi.dtype)
        if growth is not None:
            work_dtype = torch.promote_types(work_dtype, growth.dtype)
    return _PiSites.apply(conc_guide, conc_model, pi, growth, data, work_dtype)
