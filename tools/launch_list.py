"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one learner iteration in launch order, or totals."""
import collections, csv, re, sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void r2d2::<unnamed>::", "").replace("r2d2::<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        seq.append((k, row["Grid Size"], v))
    return seq


if __name__ == "__main__":
    seq = load(sys.argv[1])
    mode = sys.argv[2] if len(sys.argv) > 2 else "iter"
    idx = [i for i, s in enumerate(seq) if s[0].startswith("tree_sample")]
    a, b = idx[1], idx[2]
    if mode == "iter":
        for s in seq[a:b]:
            print(f"{s[0][:40]:42s} {s[1]:>16s} {s[2]:8.1f}")
    agg = collections.OrderedDict()
    for s in seq[a:b]:
        e = agg.setdefault(s[0], [0, 0.0]); e[0] += 1; e[1] += s[2]
    tot = sum(s[2] for s in seq[a:b])
    print(f"--- one iteration: {b - a} launches, {tot:.1f} us")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k[:48]:50s} {n:3d} {t:8.1f} {100 * t / tot:5.1f}%")
