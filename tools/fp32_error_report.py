"""fp32 kernels vs (a) the float64 truth on the reference's float32 draws and (b) the reference's own float32 results,
element-wise relative error (tests/fp32_floor.py), for every golden case.  GPU; writes one JSON object.

    python tools/fp32_error_report.py > gpurun_out/fp32_errors.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tests.fp32_floor import elem_rel, reference_fp32_floor  # noqa: E402
from tests.test_gpu_golden import FUSED, make_engine  # noqa: E402
from tests.test_reference_golden import edit_perm, group, load_case, to_ours  # noqa: E402

dev = torch.device("cuda:0")
out = {}
for name in FUSED:
    z, data = load_case(name)
    truth, floor = reference_fp32_floor(name)
    eng = make_engine(z, data, dev, torch.float32, 4)
    perm = edit_perm(z, data)
    noise = {k: torch.as_tensor(to_ours(v, perm, k)) for k, v in group(z, "native/noise/").items() if "/" not in k}
    got = eng.gradients(noise)
    row = {"loss": {"floor": floor["loss"], "vs_truth": abs(got["loss"].item() - truth["loss"]) / abs(truth["loss"]),
                    "vs_native": abs(got["loss"].item() - float(z["native/loss"])) / abs(float(z["native/loss"]))}}
    for k, g in group(z, "native/grad/").items():
        mine = got[k].detach().double().cpu().numpy().reshape(-1)
        row[k] = {"floor": floor[k], "vs_truth": elem_rel(mine, truth["grads"][k].reshape(-1)),
                  "vs_native": elem_rel(mine, to_ours(g, perm, k).reshape(-1))}
    out[name] = row
    print(name, {k: (float(f"{v['floor']:.2g}"), float(f"{v['vs_truth']:.2g}"), float(f"{v['vs_native']:.2g}")) for k, v in row.items()}, file=sys.stderr)
print(json.dumps(out))
