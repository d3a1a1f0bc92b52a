"""Static SASS statistics of one kernel of libbean_b200.so (no GPU needed).

    python tools/sass_stats.py <substring of the mangled kernel name> [--dump]

Prints instruction count, opcode histogram, spill instructions (STL/LDL), vector loads (LDG.E.128) and calls.
"""
import collections
import re
import subprocess
import sys

so = __import__('os').environ.get('SO', 'crispr_bean_b200/libbean_b200.so')
key = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for blk in blocks[1:]:
    name = blk.split("\n", 1)[0].strip()
    if key not in name:
        continue
    ins = re.findall(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", blk)
    ops = collections.Counter(op.split(".")[0] for _, op in ins)
    full = collections.Counter(op for _, op in ins)
    print(f"{name}: {len(ins)} instructions ({len(ins) * 16 / 1024:.1f} KB)")
    print("  top:", ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
    print("  MUFU:", {k: v for k, v in full.items() if k.startswith("MUFU")})
    print("  LDG:", {k: v for k, v in full.items() if k.startswith("LDG")}, " STG:", {k: v for k, v in full.items() if k.startswith("STG")})
    print("  spills: STL", ops.get("STL", 0), "LDL", ops.get("LDL", 0), " CALL", ops.get("CALL", 0), " BRA", ops.get("BRA", 0))
    if "--dump" in sys.argv:
        open("/tmp/kernel.sass", "w").write(blk)
        print("  dumped to /tmp/kernel.sass")
