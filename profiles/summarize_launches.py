"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel for ONE step of bench.py.
Usage: python profiles/summarize_launches.py gpurun_out/launches.csv [marker-kernel-substring]"""
import collections
import csv
import re
import sys


def main(path, marker="frame_u8_to_f32"):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"])
            for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    starts = [i for i, r in enumerate(rows) if marker in r[0]]
    if len(starts) < 2:
        print("need two step markers, found", len(starts)); return
    a, b = starts[-2], starts[-1]
    step = rows[a:b]
    tot = sum(r[1] for r in step)
    print(f"# {path}: one step = {len(step)} launches, {tot / 1e6:.3f} ms summed (cold-cache, serialised; compare shares)")
    agg = collections.OrderedDict()
    for name, t, grid, block in step:
        key = re.sub(r"\(.*", "", name).replace("onr::", "")
        agg.setdefault(key, [0, 0.0, 0.0])
        agg[key][0] += 1
        agg[key][1] += t
        agg[key][2] = max(agg[key][2], t)
    print(f"{'us total':>10} {'n':>4} {'share':>6} {'max us':>9}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / 1e3:10.1f} {v[0]:4d} {100 * v[1] / tot:5.1f}% {v[2] / 1e3:9.1f}  {k[:80]}")
    print("\n# launches in order")
    for name, t, grid, block in step:
        print(f"{t / 1e3:9.1f} us  grid {grid:>16} block {block:>12}  {re.sub(r'\(.*', '', name).replace('onr::', '')[:60]}")


if __name__ == "__main__":
    main(*sys.argv[1:])
