import sys; sys.path.insert(0, '.')
import ast, numpy as np, torch
from tests.test_reference_golden import load_case, group, oracle_eval
from tests.test_gpu_golden import make_engine
for name, dt in (("mixture_acc_fitnoise", torch.float64), ("mixture_ragged_lowdepth", torch.float32), ("mixture_ragged_lowdepth", torch.float64)):
    z, data = load_case(name)
    eng = make_engine(z, data, "cuda", dt, 4)
    noise = {k: torch.as_tensor(v) for k, v in group(z, "f64/noise/").items() if "/" not in k}
    got = eng.gradients(noise)
    loss, grads, _ = oracle_eval(z, data, "f64", torch.float64)
    print(name, dt, "loss gpu", got["loss"].item(), "oracle", loss, "ref", float(z["f64/loss"]))
    for k, g in group(z, "f64/grad/").items():
        a = got[k].double().cpu().numpy().reshape(-1); b = g.reshape(-1); c = grads[k].reshape(-1)
        i = np.abs(a-b).argmax()
        print(f"  {k:12s} max|gpu-ref| {np.abs(a-b).max():.3e} at {i}: gpu {a[i]:.12g} ref {b[i]:.12g} oracle {c[i]:.12g}  mean|ref| {np.abs(b).mean():.3e}")
    if dt == torch.float32 and name.startswith("mixture_ragged"):
        k = "alpha_pi"
        a = got[k].double().cpu().numpy(); b = z["f64/grad/alpha_pi"]
        bad = np.argwhere(np.abs(a-b) > 1e-4*np.abs(b).mean())
        print("bad alpha entries", bad[:10].tolist())
        for g_, j in bad[:5]:
            print("  guide", g_, "pi_a0", float(data.pi_a0[g_]), "pi", noise["pi"][:, 0, g_].tolist(), "counts", data.allele_counts_control[:, 0, g_].tolist(), "gpu", a[g_], "ref", b[g_])
