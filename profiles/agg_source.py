"""Aggregate `ncu --page source --csv --print-source cuda,sass` output per CUDA source line.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python profiles/agg_source.py src.csv [top_n]
"""
import collections
import csv
import sys


def num(s):
    try:
        return int(s)
    except (ValueError, TypeError):
        return 0


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur, ie = None, None
    agg = collections.OrderedDict()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            ie = r.index("Instructions Executed")
        elif r[0].isdigit() and ie is not None and len(r) > ie and r[2] == "-":
            agg[(cur, int(r[0]), r[1].strip()[:100])] = (num(r[ie]), num(r[4]))
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    byfile = collections.Counter()
    for k, v in agg.items():
        byfile[k[0]] += v[0]
    print("total warp instructions", tot, "stall samples", tots)
    print({k: f"{v / tot * 100:.1f}%" for k, v in byfile.items()})
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0] / tot * 100:5.1f}% inst {v[1] / tots * 100:5.1f}% stall  {k[0]}:{k[1]}  {k[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
