"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares.

    python profiles/analyze_launches.py gpurun_out/launches.csv [--skip N] [--top K] > profiles/rNN_launches_summary.txt

Per-launch times under ncu are cold-cache and serialised, so compare SHARES, not absolutes (B200_PROFILING.md).
"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mv, mn, mu, gs, bs = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit", "Grid Size", "Block Size"))
    out = []
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", ""))
        u = r[mu]
        if u == "ns":
            v /= 1e3
        elif u == "ms":
            v *= 1e3
        elif u == "s":
            v *= 1e6
        out.append((int(r[0]), r[kn], r[gs], r[bs], v))
    return out


def short(name):
    name = name.replace("void ", "").replace("ssg::", "").replace("__nv_bfloat16", "bf16")
    i = name.find("(")
    return name[:i] if i > 0 else name


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    L = [x for x in load(path) if x[0] >= skip]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for _, k, _, _, v in L:
        a = agg[short(k)]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("launches %d   total %.1f us (serialised, cold-cache: compare shares)" % (len(L), tot))
    ours = sum(v[1] for k, v in agg.items() if not k.startswith("at::") and "nccl" not in k.lower())
    print("own kernels: %.1f%% of device time; torch (autograd grad accumulation / fills / copies): %.1f%%" % (100 * ours / tot, 100 * (tot - ours) / tot))
    print("%12s %7s %6s  kernel" % ("us", "share", "n"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%12.1f %6.1f%% %6d  %s" % (v[1], 100 * v[1] / tot, v[0], k[:110]))


if __name__ == "__main__":
    main()
