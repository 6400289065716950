"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel launches, total time and share of ONE
batched call (the last `per_call` launches of the log).

    python tools/launch_shares.py gpurun_out/launches.csv 71 > profiles/r02/launches_one_batched_call.csv
"""
import collections
import csv
import re
import sys


def main():
    path, per_call = sys.argv[1], int(sys.argv[2])
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if len(r) > 10]
    hdr, rows = rows[0], rows[1:]
    ix = {h: i for i, h in enumerate(hdr)}
    rows = rows[-per_call:]
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"^(void )?gf::", "", r[ix["Kernel Name"]])
        name = re.sub(r"\(.*$", "", name)
        us = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        us *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0, r[ix["Block Size"]], r[ix["Grid Size"]]])
        a[0] += 1
        a[1] += us
        a[3] = r[ix["Grid Size"]]
    tot = sum(a[1] for a in agg.values())
    w = csv.writer(sys.stdout)
    print("# ncu --metrics gpu__time_duration.sum --clock-control none: the launches of ONE batched call (%d kernels) of" % per_call)
    print("# python bench.py --steps 20 --warmup 5 --repeats 1 --no-graph (serialised, cold cache: SHARES, not absolute times)")
    w.writerow(["kernel", "launches", "total_us", "share_pct", "block", "grid_of_last"])
    for name, a in agg.items():
        w.writerow([name, a[0], round(a[1], 1), round(100 * a[1] / tot, 1), a[2], a[3]])


if __name__ == "__main__":
    main()
