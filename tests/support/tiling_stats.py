"""Per-variant summaries `bean run` adds to the tiling result table, from the CSR allele map instead of the dense
(G, A-1, E) `allele_to_edit` tensor (429 MB at the published 7,488-guide example).

Host-side mirrors of bean/preprocessing/utils.py:254-310 (`_obtain_effective_edit_rate`,
`_obtain_n_guides_alleles_per_variant`, `_obtain_n_cooccurring_variants`), called at bean/cli/run.py:176-201.
A "slot" is one (guide, edited allele) pair; `allele_ptr / allele_edit` list the edits of every slot.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch


def _slots(ndata):
    ptr = ndata.allele_ptr.numpy().astype(np.int64)
    edits = ndata.allele_edit.numpy().astype(np.int64)
    k = np.diff(ptr)                                  # edits per slot
    slot_of = np.repeat(np.arange(len(k)), k)         # slot of every (slot, edit) entry
    return ptr, edits, k, slot_of


def _obtain_effective_edit_rate(ndata, count_thres=10) -> Tuple[List[torch.Tensor], List[list], torch.Tensor]:
    """Effective editing rate of each variant: an allele's control-sample rate (mean over replicates and control conditions
    with at least `count_thres` barcode-matched reads) is split evenly over its edits and summed per (guide, edit).

    Returns (guides generating each variant [(n, 1) index tensors], their per-guide rates [lists], total rate (E,))."""
    G, A, E = ndata.n_guides, ndata.n_max_alleles, ndata.n_edits
    bc = ndata.X_bcmatch_control
    rates = ndata.allele_counts_control / bc[:, :, :, None]
    rates = torch.where((bc < count_thres)[:, :, :, None].expand(rates.shape), torch.full_like(rates, float("nan")), rates)
    mean_rate = rates.nanmean(axis=(0, 1))[:, 1:].reshape(-1).numpy()  # (G * (A - 1),) per slot; NaN where never covered
    _, edits, k, slot_of = _slots(ndata)
    share = mean_rate[slot_of] / k[slot_of].astype(mean_rate.dtype)
    share = np.where(np.isnan(share), 0, share).astype(mean_rate.dtype)  # nansum
    guide_of = slot_of // (A - 1)
    key = guide_of * E + edits
    uniq, inv = np.unique(key, return_inverse=True)
    pair_rate = np.zeros(len(uniq), dtype=mean_rate.dtype)
    np.add.at(pair_rate, inv, share)
    pair_guide, pair_edit = uniq // E, uniq % E
    total = np.zeros(E, dtype=mean_rate.dtype)
    np.add.at(total, pair_edit, pair_rate)
    order = np.lexsort((pair_guide, pair_edit))  # by edit, guides ascending
    pair_guide, pair_edit, pair_rate = pair_guide[order], pair_edit[order], pair_rate[order]
    live = pair_rate > 0
    bounds = np.searchsorted(pair_edit, np.arange(E + 1))
    guide_idx, per_guide = [], []
    for e in range(E):
        sl = slice(bounds[e], bounds[e + 1])
        sel = live[sl]
        guide_idx.append(torch.as_tensor(pair_guide[sl][sel]).reshape(-1, 1))
        per_guide.append(pair_rate[sl][sel].tolist())
    return guide_idx, per_guide, torch.as_tensor(total)


def _obtain_n_guides_alleles_per_variant(ndata) -> torch.Tensor:
    """Number of guides that produce each variant in at least one of their alleles, (E,)."""
    _, edits, _, slot_of = _slots(ndata)
    pairs = np.unique((slot_of // (ndata.n_max_alleles - 1)) * ndata.n_edits + edits)
    return torch.as_tensor(np.bincount(pairs % ndata.n_edits, minlength=ndata.n_edits))


def _obtain_n_cooccurring_variants(ndata) -> np.ndarray:
    """Number of OTHER variants seen together with each variant in any allele of any guide, (E,)."""
    E = ndata.n_edits
    ptr, edits, k, slot_of = _slots(ndata)
    # all ordered pairs (e, e') within a slot, the pair (e, e) included: per slot k^2 entries
    first = np.repeat(np.arange(len(edits)), k[slot_of])                     # entry i repeated k(slot(i)) times
    offs = np.arange(len(first)) - np.repeat(np.cumsum(k[slot_of]) - k[slot_of], k[slot_of])
    second = ptr[slot_of[first]] + offs
    pairs = np.unique(edits[first] * E + edits[second])
    n = np.bincount(pairs // E, minlength=E) - 1
    n[np.bincount(edits, minlength=E) == 0] = -1  # an edit in no allele: the reference's empty sum gives 0 - 1
    return n
