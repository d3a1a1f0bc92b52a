"""Guide accessibility from a bigWig ATAC / DNase track (`bean run --scale-by-acc --acc-bw-path`).

Host-side mirror of bean/preprocessing/utils.py:70-146 (`_get_accessibility_single`, `get_accessibility_guides`) on the
pure-Python reader `tests/support/bigwig.py` instead of pyBigWig: the geometric mean of (signal + 1) over the
`half_window_size` bases either side of the guide's `genomic_pos` (NaN bases ignored), NaN guides filled with the median.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from . import bigwig


def _get_accessibility_single(pos, track, chrom: str = "chr19", guide_start_pos=0, half_window_size: int = 100):
    if half_window_size < 0:
        raise ValueError("Window size must be non-negative.")
    if pos == "control" or np.isnan(pos):
        return np.nan
    centre = guide_start_pos + pos
    try:
        window = track.values(chrom, int(centre - half_window_size), int(centre + half_window_size))
        return np.exp(np.nanmean(np.log(np.asarray(window) + 1.0)))
    except Exception as exc:  # out-of-range windows, unknown chromosomes: the guide gets the median (as the reference)
        print(exc)
        return np.nan


def get_accessibility_guides(accessibility_bw_path: str, guide_info: pd.DataFrame, half_window_size: int = 100) -> torch.Tensor:
    """(G,) float64 tensor from `guide_info` columns `genomic_pos` and `chrom` (or `chr`; chr19 when neither exists)."""
    track = bigwig.open(accessibility_bw_path)
    if "chr" in guide_info.columns and "chrom" not in guide_info.columns:
        guide_info = guide_info.rename(columns={"chr": "chrom"})
    chroms = guide_info["chrom"].tolist() if "chrom" in guide_info.columns else ["chr19"] * len(guide_info)
    acc = torch.as_tensor(np.asarray([_get_accessibility_single(pos, track, chrom=c, half_window_size=half_window_size)
                                      for pos, c in zip(guide_info["genomic_pos"].tolist(), chroms)], dtype=np.float64))
    if torch.isnan(acc).all():
        raise ValueError("Cannot retrieve guide accessibility from the bigWig file. Check your inputs.")
    acc[torch.isnan(acc)] = torch.nanmedian(acc)
    return acc
