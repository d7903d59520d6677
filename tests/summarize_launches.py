"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list (development aid)."""
import collections
import csv
import re
import sys


def summarize(path, out=sys.stdout):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt, mx, seq = collections.defaultdict(float), collections.Counter(), collections.defaultdict(float), []
    for row in csv.DictReader(lines):
        v, unit = float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        short = re.sub(r"\(.*", "", row["Kernel Name"])
        short = re.sub(r"^void |cilqr::|<unnamed>::", "", short)
        tot[short] += v
        cnt[short] += 1
        mx[short] = max(mx[short], v)
        seq.append((short, v))
    T = sum(tot.values())
    print("total kernel time %.2f ms over %d launches" % (T / 1000, len(seq)), file=out)
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("%-40s n=%5d  total %10.1f us (%5.1f%%)  mean %8.1f  max %8.1f" % (k[:40], cnt[k], v, 100 * v / T, v / cnt[k], mx[k]), file=out)
    return seq


if __name__ == "__main__":
    seq = summarize(sys.argv[1])
    idx = [j for j, (n, _) in enumerate(seq) if n.startswith("k_derivs")]
    for st in idx[:2] + idx[len(idx) // 2: len(idx) // 2 + 1] + idx[-3:-2]:
        print([(n.split("<")[0], round(v, 1)) for n, v in seq[st:st + 6]])
