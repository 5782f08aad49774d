"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python scripts/summarize_launches.py gpurun_out/launches_c4.csv > profiles/r01_launches_c4.txt
"""
import collections
import csv
import sys


def main(path):
    hdr = None
    agg = collections.OrderedDict()
    for r in csv.reader(open(path)):
        if len(r) < 6:
            continue
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = d["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(d["Metric Value"].replace(",", ""))
        scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}[d["Metric Unit"]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    import re
    pat = re.compile(r"^(osfm::)?(ba_|bf_|pcg_|tc_|ord_|bsr_|pad_rows)")
    ours = {pat.sub(lambda m: m.group(2), k): v for k, v in agg.items() if pat.match(k)}
    tot = sum(v[1] for v in ours.values())
    print("# %s" % path)
    print("# kernels of libopensfm_b200.so only (cub / torch scene-generation kernels of the same process left out)")
    print("# ncu serialises launches and runs them cold: compare SHARES with bench.py's event timers, not absolute times")
    print("%-44s %8s %12s %12s %7s" % ("kernel", "launches", "total ms", "avg ms", "share"))
    for k, (n, ms) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %8d %12.3f %12.4f %6.1f%%" % (k[:44], n, ms, ms / n, 100.0 * ms / tot))
    print("%-44s %8s %12.3f" % ("total", "", tot))
    other = sum(v[1] for k, v in agg.items() if not pat.match(k))
    print("%-44s %8s %12.3f" % ("(other kernels in the process)", "", other))


if __name__ == "__main__":
    main(sys.argv[1])
