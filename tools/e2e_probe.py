"""Why bench.py's e2e loop (run(1) + one pinned D2H copy per step) can take 1.5 ms per step at c5: variants of the loop, timed on
the device (CUDA events) and on the host (perf_counter).  GPU diagnostic."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import build_data  # noqa: E402
from crispr_bean_b200.svi import SviEngine  # noqa: E402

dev = torch.device("cuda:0")
data = build_data("c5_genome_scale", 101)
eng = SviEngine(data, "MixtureNormal", dev, dtype=torch.float32, num_steps=2000, seed=101)
eng.run(400)
torch.cuda.synchronize()
data.pin_memory()
del eng
out = {}
for label in ("copy_each_step", "no_copy", "one_call", "copy_each_step_again", "steps_2000_schedule"):
    N = 100
    loss_host = torch.zeros(N, dtype=torch.float64).pin_memory()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.perf_counter()
    e0.record()
    eng2 = SviEngine(data, "MixtureNormal", dev, dtype=torch.float32, num_steps=2000 if label == "steps_2000_schedule" else N, seed=7)
    e1.record()
    t1 = time.perf_counter()
    if label == "one_call":
        eng2.run(N)
    else:
        for t in range(N):
            eng2.run(1)
            if label != "no_copy":
                loss_host[t].copy_(eng2.loss[t], non_blocking=True)
    t2 = time.perf_counter()
    e2.record()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    out[label] = {"device_setup_ms": e0.elapsed_time(e1), "device_steps_ms": e1.elapsed_time(e2), "host_setup_ms": (t1 - t0) * 1e3,
                  "host_loop_issue_ms": (t2 - t1) * 1e3, "host_until_done_ms": (t3 - t1) * 1e3}
    del eng2
print(json.dumps(out))
