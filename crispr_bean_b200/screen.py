"""Minimal count-matrix container with the attribute surface `bean run` reads from a ReporterScreen.

The reference's container (`bean/framework/ReporterScreen.py`, an anndata.AnnData subclass) is out of
scope and not importable here (anndata / perturb_tools are absent).  The tensoriser in
`data_class.py` only touches `.X`, `.layers[...]`, `.guides`, `.samples`, `.uns` and 2-D slicing
(SURVEY App. A.10), so this duck-typed stand-in -- or a real ReporterScreen -- can be passed in.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import pandas as pd


def _take_columns(a: np.ndarray, cols: np.ndarray) -> np.ndarray:
    """a[:, cols] as a ROW-MAJOR array, like an AnnData slice (float32 column means depend on the layout: the a0 fit).
    np.take writes row-major directly; `a[:, cols]` would come out column-major and need a second pass."""
    return np.take(a, cols, axis=1)


class MiniScreen:
    """guides x samples count matrix + per-guide / per-sample tables (+ named layers)."""

    def __init__(self, X, guides: pd.DataFrame, samples: pd.DataFrame,
                 layers: Optional[Dict[str, np.ndarray]] = None, uns: Optional[dict] = None):
        self.X = np.asarray(X)
        self.guides = guides
        self.samples = samples
        self.layers = {k: np.asarray(v) for k, v in (layers or {}).items()}
        self.uns = dict(uns or {})
        assert self.X.shape == (len(guides), len(samples)), (self.X.shape, len(guides), len(samples))
        for k, v in self.layers.items():
            assert v.shape == self.X.shape, (k, v.shape)

    # anndata aliases used by the reference
    @property
    def obs(self):
        return self.guides

    @property
    def var(self):
        return self.samples

    @property
    def n_obs(self):
        return self.X.shape[0]

    @property
    def n_vars(self):
        return self.X.shape[1]

    @staticmethod
    def _resolve(idx, index: pd.Index):
        if isinstance(idx, slice):
            return np.arange(len(index))[idx]
        if isinstance(idx, (pd.Series, pd.Index)):
            idx = idx.to_numpy()
        idx = np.asarray(idx)
        if idx.ndim == 0 and idx.dtype.kind in "biu":
            return np.asarray([int(idx)])  # AnnData: a scalar (bool included, an int subclass) selects that one position
        if idx.dtype == bool:
            return np.nonzero(idx)[0]
        if idx.dtype.kind in "OUS":
            return index.get_indexer(idx)
        return idx.astype(np.int64)

    def __getitem__(self, key):
        gi, si = key if isinstance(key, tuple) else (key, slice(None))
        all_guides = isinstance(gi, slice) and gi == slice(None)
        all_samples = isinstance(si, slice) and si == slice(None)
        si = self._resolve(si, self.samples.index)
        if all_guides:  # sample subsets / reorders of a whole library (1M guides): one column gather per matrix
            if all_samples or np.array_equal(si, np.arange(len(self.samples))):
                take = lambda a: a
            else:
                take = lambda a: _take_columns(a, si)
            guides, uns = self.guides.copy(), dict(self.uns)
        else:
            gi = self._resolve(gi, self.guides.index)
            take = lambda a: a[np.ix_(gi, si)]
            guides, uns = self.guides.iloc[gi].copy(), dict(self.uns)
            for k, v in self.uns.items():  # per-guide tables follow the guide subset (repguide_mask)
                if isinstance(v, pd.DataFrame) and len(v) == len(self.guides) and v.index.equals(self.guides.index):
                    uns[k] = v.iloc[gi]
        return MiniScreen(take(self.X), guides, self.samples.iloc[si].copy(), {k: take(v) for k, v in self.layers.items()}, uns)

    def copy(self):
        """Own tables, shared count matrices (they are never written in place)."""
        return self[:, :]

    # ---- the ReporterScreen members `bean run` touches besides the tables (bean/framework/ReporterScreen.py) -----------
    @property
    def tiling(self):
        return self.uns["tiling"]  # ReporterScreen.py:172-174

    @property
    def target_base_changes(self):
        """{"A": "G", ...} from uns["target_base_changes"] ("A>G,C>T"; ReporterScreen.py:164-170)."""
        spec = self.uns["target_base_changes"] if "target_base_changes" in self.uns else self.uns["target_base_change"]
        return {change[0]: change[-1] for change in spec.split(",")}

    def get_guide_edit_rate(self, normalize_by_editable_base=None, edited_bases=None, editable_base_start=3, editable_base_end=8,
                            bcmatch_thres=1, prior_weight=None, return_result=False, count_layer="X_bcmatch", edit_layer="edits",
                            condition_col="condition", unsorted_condition_label=None):
        """Per-guide reporter editing rate in the unsorted samples, written to `guides["edit_rate"]` (and
        `guides["edit_rate_norm"]`, per editable base of the guide's activity window, for tiling screens).

        Restatement of ReporterScreen.get_guide_edit_rate (ReporterScreen.py:448-529; that module needs anndata, so it is
        not executed by tests/refharness): rate = (edits + w/2) / (barcode-matched reads + w/2) over the samples whose
        condition contains `unsorted_condition_label`, w = prior_weight (default 1); NaN below `bcmatch_thres` reads."""
        if normalize_by_editable_base is None:
            normalize_by_editable_base = self.tiling
        if count_layer not in self.layers or edit_layer not in self.layers:
            raise ValueError("edits or barcode matched guide counts not available.")
        n_sites = 1.0
        if normalize_by_editable_base:
            bases = list(self.target_base_changes.keys()) if edited_bases is None else ([edited_bases] if isinstance(edited_bases, str) else edited_bases)
            if any(b not in ("A", "C", "T", "G") for b in bases):
                raise ValueError("Specify the correct edited_base")
            window = self.guides["sequence"].map(lambda seq: seq[editable_base_start:editable_base_end])
            n_sites = sum(window.map(lambda w, b=b: w.count(b)) for b in bases)
        if unsorted_condition_label is not None:
            cols = np.where(self.samples[condition_col].astype(str).map(lambda c: unsorted_condition_label in c))[0]
            if len(cols) == 0:
                raise ValueError(f"'{unsorted_condition_label}' is not found in ReporterScreen.samples['{condition_col}'] that has values "
                                 f"{self.samples[condition_col].unique()}. Check your input.")
        else:
            cols = np.arange(len(self.samples))
        w = 1 if prior_weight is None else prior_weight
        n_edits = self.layers[edit_layer][:, cols].sum(axis=1)
        n_counts = self.layers[count_layer][:, cols].sum(axis=1)
        rate = (n_edits + w / 2) / (n_counts + w / 2)
        rate[n_counts < bcmatch_thres] = np.nan
        if normalize_by_editable_base:
            sites = np.asarray(n_sites, dtype=np.float64)
            rate_norm = (n_edits + w / 2) / (n_counts * sites + w / 2)
            rate_norm[sites == 0] = np.nan
        if return_result:
            return rate
        self.guides["edit_rate"] = rate
        if normalize_by_editable_base:
            self.guides["edit_rate_norm"] = rate_norm
        return None


def read_csvs(guides_csv: str, samples_csv: str, counts_csv: str,
              layer_csvs: Optional[Dict[str, str]] = None) -> MiniScreen:
    """Load the CSV triple `bean create-screen` consumes (framework/read_from_csvs.py:9-23)."""
    guides = pd.read_csv(guides_csv, index_col=0)
    guides = guides.loc[:, ~guides.columns.duplicated()]
    samples = pd.read_csv(samples_csv, index_col=0)
    counts = pd.read_csv(counts_csv, index_col=0)
    counts = counts.loc[guides.index, samples.index]
    layers = {}
    for name, path in (layer_csvs or {}).items():
        layers[name] = pd.read_csv(path, index_col=0).loc[guides.index, samples.index].to_numpy()
    return MiniScreen(counts.to_numpy(), guides, samples, layers)


# ---- .h5ad input without h5py / anndata ----------------------------------------------
def _h5ad_value(node):
    """Decode one element of anndata's on-disk format (encoding-type attributes, anndata >= 0.8)."""
    from . import h5lite

    enc = node.attrs.get("encoding-type", None)
    if isinstance(node, h5lite.Dataset):
        val = node.read()
        if enc in ("string", "numeric-scalar") and isinstance(val, np.ndarray) and val.shape == ():
            val = val[()]
        return val
    if enc == "dataframe":
        return _h5ad_dataframe(node)
    if enc == "categorical":
        codes, cats = node["codes"].read(), node["categories"].read()
        return pd.Categorical.from_codes(codes, categories=pd.Index(cats), ordered=bool(node.attrs.get("ordered", False)))
    if enc in ("csr_matrix", "csc_matrix"):
        from scipy import sparse

        shape = tuple(int(v) for v in node.attrs["shape"])
        cls = sparse.csr_matrix if enc == "csr_matrix" else sparse.csc_matrix
        return cls((node["data"].read(), node["indices"].read(), node["indptr"].read()), shape=shape).toarray()
    if enc in ("nullable-integer", "nullable-boolean"):
        vals, mask = node["values"].read(), node["mask"].read().astype(bool)
        return pd.array(np.where(mask, 0, vals), dtype="Int64" if enc == "nullable-integer" else "boolean").__setitem__(mask, pd.NA) or vals
    return {k: _h5ad_value(node[k]) for k in node.keys()}  # "dict" and anything unknown: a plain mapping


def _h5ad_dataframe(group) -> pd.DataFrame:
    index_key = group.attrs.get("_index", "_index")
    order = [str(c) for c in np.atleast_1d(group.attrs.get("column-order", []))]
    index = pd.Index(np.asarray(group[index_key].read()), name=None if index_key == "_index" else index_key)
    cols = {}
    for c in order:
        cols[c] = _h5ad_value(group[c])  # a "/" in a column name is a nested HDF5 group; the path lookup follows it
    return pd.DataFrame(cols, index=index)


def read_h5ad(path: str) -> MiniScreen:
    """Read a ReporterScreen `.h5ad` (bean/framework/ReporterScreen.py:1009) into a MiniScreen with the pure-Python
    HDF5 reader `h5lite`: X (guides x samples), layers, obs -> .guides, var -> .samples, uns (nested tables, scalars)."""
    from . import h5lite

    f = h5lite.File(path)
    X = _h5ad_value(f["X"])
    guides, samples = _h5ad_dataframe(f["obs"]), _h5ad_dataframe(f["var"])
    layers = {k: np.asarray(_h5ad_value(f["layers"][k])) for k in f["layers"].keys()} if "layers" in f.root else {}
    uns = _h5ad_value(f["uns"]) if "uns" in f.root else {}
    return MiniScreen(np.asarray(X), guides, samples, layers, uns)
