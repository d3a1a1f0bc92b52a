"""N steps of the fused survival MixtureNormal step at c4 scale (1M guides x 3 replicates x 3 timepoints): the target of the
ncu capture of `surv_guide_kernel` (tools/gpu_session_e.sh) and a per-100-step timing of a run.

    python profiles/survival_steady.py [n_steps] [n_variants]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import build_data  # noqa: E402
from crispr_bean_b200.survival_fused import SurvivalFusedEngine  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
data = build_data("c4_survival", 101)
eng = SurvivalFusedEngine(data, "cuda", num_steps=N)
for blk in range(N // 100):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.run(100)
    b.record()
    torch.cuda.synchronize()
    q0 = eng.q0_u.exp()
    print(f"steps {blk * 100:5d}-{blk * 100 + 99:5d}: {a.elapsed_time(b) / 100:.3f} ms/step  loss {eng.loss[eng.step - 1].item():.6g}  "
          f"q0 sum {q0.sum().item():.4g} max {q0.max().item():.3g}  x at clamp: {(eng.gamma[0] <= 2e-38).float().mean().item():.4f}", flush=True)
