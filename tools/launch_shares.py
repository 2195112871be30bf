"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel launches / average / total / share."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, out):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rd:
        if len(r) <= iv:
            continue
        name = re.sub(r"\(.*", "", r[ik]).strip()
        v = float(r[iv].replace(",", ""))
        us = v / 1e3 if r[iu] in ("ns", "nsecond") else (v if r[iu] in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("kernel,launches,avg_us,total_us,share\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.1f,%.1f,%.3f\n" % (k, n, t / n, t, t / tot))
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
