"""Helpers shared by tests/golden/make_reference_golden.py (writer) and tests/test_reference_golden.py (reader).

A golden case = one screen + the tensors the REFERENCE's data class made of it + one (or several) draws of
reparameterisation noise + what the REFERENCE's model/guide programs computed from them (loss, per-site
log-prob sums, gradients w.r.t. the unconstrained parameters), in float64 ("f64": default dtype switched to
double, data cast to double) and in the reference's own mixed precision ("native": float32 default dtype).
"""
from __future__ import annotations

import copy
import os

import numpy as np
import pandas as pd
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden")


# ---- screen <-> flat dict of arrays ----------------------------------------------------------------
def screen_to_arrays(scr, prefix="screen/"):
    out = {prefix + "X": np.asarray(scr.X)}
    for k, v in scr.layers.items():
        out[f"{prefix}layer/{k}"] = np.asarray(v)
    for tag, df in (("guides", scr.guides), ("samples", scr.samples)):
        out[f"{prefix}{tag}/index"] = df.index.to_numpy().astype(str)
        for c in df.columns:
            col = df[c].to_numpy()
            out[f"{prefix}{tag}/col/{c}"] = col.astype(str) if col.dtype.kind in "OUS" else col
    for key, tbl in scr.uns.items():  # per-allele count tables of tiling screens (values may be Allele objects -> str)
        if isinstance(tbl, (list, tuple)) and all(isinstance(v, str) for v in tbl):
            out[f"{prefix}uns_list/{key}"] = np.asarray(list(tbl)).astype(str)
        if isinstance(tbl, pd.DataFrame):
            out[f"{prefix}uns/{key}/columns"] = np.asarray(list(tbl.columns)).astype(str)
            for c in tbl.columns:
                col = tbl[c].to_numpy()
                out[f"{prefix}uns/{key}/col/{c}"] = np.asarray([str(v) for v in col]) if col.dtype.kind in "OUS" else col
    return out


def screen_from_arrays(z, prefix="screen/"):
    from crispr_bean_b200.screen import MiniScreen

    def frame(tag):
        idx = pd.Index(z[f"{prefix}{tag}/index"].astype(str), name="name")
        cols = {k.split("/col/", 1)[1]: z[k] for k in z.files if k.startswith(f"{prefix}{tag}/col/")}
        return pd.DataFrame({c: (v.astype(str) if v.dtype.kind in "US" else v) for c, v in cols.items()}, index=idx)

    layers = {k.split("layer/", 1)[1]: z[k] for k in z.files if k.startswith(prefix + "layer/")}
    uns = {}
    for k in z.files:
        if k.startswith(prefix + "uns/") and k.endswith("/columns"):
            key = k[len(prefix) + 4:-len("/columns")]
            uns[key] = pd.DataFrame({c: (z[f"{prefix}uns/{key}/col/{c}"].astype(str) if z[f"{prefix}uns/{key}/col/{c}"].dtype.kind in "US"
                                         else z[f"{prefix}uns/{key}/col/{c}"]) for c in z[k].astype(str)})
    for k in z.files:
        if k.startswith(prefix + "uns_list/"):
            uns[k[len(prefix) + 9:]] = [str(v) for v in z[k]]
    return MiniScreen(z[prefix + "X"], frame("guides"), frame("samples"), layers, uns)


def data_tensors(data):
    """Every tensor / scalar attribute of a *ScreenData object, as numpy."""
    out = {}
    for k, v in vars(data).items():
        if torch.is_tensor(v):
            out[k] = v.detach().cpu().numpy()
        elif isinstance(v, (bool, int, float, np.integer, np.floating)):
            out[k] = np.asarray(v)
    return out


def cast_floats(data, dtype):
    nd = copy.copy(data)
    for k, v in vars(data).items():
        if torch.is_tensor(v) and v.is_floating_point():
            setattr(nd, k, v.to(dtype))
    return nd


# ---- running the reference programs ---------------------------------------------------------------
def noise_from_guide_trace(pyro, gt):
    """Standardised reparameterisation noise behind the guide's draws (what the oracle / kernels inject)."""
    st = pyro.get_param_store()
    val = lambda k: gt.nodes[k]["value"].detach()
    sampled = {k for k, site in gt.nodes.items() if site["type"] == "sample"}  # (param sites share the trace)
    noise = {}
    if "mu_targets" in gt:
        noise["eps_mu"] = (val("mu_targets") - st["mu_loc"].detach()) / st["mu_scale"].detach()
    if "sd_targets" in gt:
        noise["eps_sd"] = (val("sd_targets").log() - st["sd_loc"].detach()) / st["sd_scale"].detach()
    if "mu_cov" in sampled:
        noise["eps_cov"] = (val("mu_cov") - st["mu_cov_loc"].detach()) / st["mu_cov_scale"].detach()
    if "pi" in gt:
        noise["pi"] = val("pi")
    if "logit_pi_noise" in gt:
        if "noise_loc" in st:
            noise["eps_noise"] = (val("logit_pi_noise") - st["noise_loc"].detach()) / st["noise_scale"].detach()
        else:
            noise["eps_noise"] = val("logit_pi_noise") / 0.655
    for k in ("initial_guide_abundance", "initial_abundance"):
        if k in sampled:
            noise["q0"] = val(k)
    return noise


def reference_loss_and_grads(pyro, model, guide, data, seed, dtype):
    """One Trace_ELBO evaluation of the reference programs at their initial parameters."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        d = cast_floats(data, dtype) if dtype == torch.float64 else data
        pyro.clear_param_store()
        torch.manual_seed(seed)
        elbo = pyro.infer.Trace_ELBO()
        loss = elbo.differentiable_loss(model, guide, d)
        mt, gt = elbo.last_traces
        loss.backward()
        st = pyro.get_param_store()
        out = {"loss": np.asarray(float(loss.detach()))}
        for k in st.keys():
            g = st.unconstrained(k).grad
            out[f"grad/{k}"] = (g if g is not None else torch.zeros_like(st.unconstrained(k))).detach().double().numpy()
            out[f"param0/{k}"] = st[k].detach().double().numpy()
        for tag, tr in (("model", mt), ("guide", gt)):
            for name, site in tr.nodes.items():
                if site["type"] == "sample":
                    out[f"site/{tag}/{name}"] = np.asarray(float(site["log_prob_sum"].detach()))
        noise = {f"noise/{k}": v.numpy() for k, v in noise_from_guide_trace(pyro, gt).items()}  # in the run's own dtypes
        for name, site in mt.nodes.items():  # model-only latents (drawn from the prior in the model)
            if site["type"] == "sample" and not site["is_observed"] and name not in gt:
                noise[f"noise/model_only/{name}"] = site["value"].detach().double().numpy()
                if name == "mu_negctrl":  # Normal(m0, s0) prior draw -> its standard-normal noise
                    fn = site["fn"]
                    noise["noise/eps_negctrl"] = ((site["value"] - fn.loc) / fn.scale).detach().double().numpy()
        return out, noise
    finally:
        torch.set_default_dtype(old)
        torch.autograd.set_detect_anomaly(False)  # the reference switches it on as a side effect (App. B6)
