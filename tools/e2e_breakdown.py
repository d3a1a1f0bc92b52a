"""Where the end-to-end time of bench.py's e2e leg goes (c5, one GPU): upload, re-tiling, data-only constants, steps, read-back.
Wall clock with a device synchronise after every stage (diagnostic; the bench itself does not synchronise in between)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import build_data  # noqa: E402
from crispr_bean_b200.device_pack import DeviceScreen  # noqa: E402
from crispr_bean_b200.svi import SviEngine  # noqa: E402

dev = torch.device("cuda:0")
data = build_data("c5_genome_scale", 101)
data.pin_memory()
torch.cuda.synchronize()
out = {}


def lap(name, t0):
    torch.cuda.synchronize()
    out[name] = round((time.perf_counter() - t0) * 1e3, 3)
    return time.perf_counter()


for trial in range(2):  # second trial: allocator warm
    t = time.perf_counter()
    x = [data.X_masked.to(dev, non_blocking=True), data.X_bcmatch_masked.to(dev, non_blocking=True)]
    t = lap(f"t{trial}_h2d_counts_256MB", t)
    scr = DeviceScreen(data, dev, dtype=torch.float32)
    t = lap(f"t{trial}_device_screen_total", t)
    eng = SviEngine(data, "MixtureNormal", dev, dtype=torch.float32, num_steps=20, seed=7, screen=scr)
    t = lap(f"t{trial}_engine_rest", t)
    eng.run(20)
    t = lap(f"t{trial}_20_steps", t)
    p = {k: v.cpu() for k, v in eng.params().items()}
    t = lap(f"t{trial}_params_to_host", t)
    del x, scr, eng, p
print(json.dumps(out))
