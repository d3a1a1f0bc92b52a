"""N steps of the fused tiling step (MultiMixtureNormal) at BASELINE config c3 (800 guides x <= 16 alleles, 4 replicates): the
target of the ncu launch list / capture of `tiling_guide_kernel`, and a per-100-step timing of a run.

    python profiles/tiling_steady.py [n_steps] [n_guides]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from crispr_bean_b200.data_class import TilingSortingReporterScreenData  # noqa: E402
from crispr_bean_b200.synth import make_tiling_screen  # noqa: E402
from crispr_bean_b200.tiling_fused import TilingFusedEngine  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
G = int(sys.argv[2]) if len(sys.argv) > 2 else 800
scr = make_tiling_screen(n_guides=G, max_alleles=16, n_reps=4, seed=3)
data = TilingSortingReporterScreenData(scr, control_can_be_selected=True, allele_df_key="allele_counts")
eng = TilingFusedEngine(data, "cuda", num_steps=N)
for blk in range(N // 100):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.run(100)
    b.record()
    torch.cuda.synchronize()
    print(f"steps {blk * 100:5d}-{blk * 100 + 99:5d}: {a.elapsed_time(b) / 100:.4f} ms/step  loss {eng.loss[eng.step - 1].item():.6g}", flush=True)
