"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown on stdout).

    python profiles/summarize.py profiles/r1_launches_final.csv [first_launch last_launch]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))[lo:hi]
    agg, order, tot = {}, [], 0.0
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("efgh::<unnamed>::", "").replace("<unnamed>::", "")
        us = float(r["Metric Value"]) / (1000.0 if r["Metric Unit"] in ("ns", "nsecond") else 1.0)
        if name not in agg:
            agg[name] = [0, 0.0]
            order.append(name)
        agg[name][0] += 1
        agg[name][1] += us
        tot += us
    print("| kernel | launches | total us | share |")
    print("|---|---|---|---|")
    for name in sorted(order, key=lambda n: -agg[n][1]):
        c, us = agg[name]
        print("| `%s` | %d | %.1f | %.1f %% |" % (name, c, us, 100 * us / tot))
    print("| **all** | %d | %.1f | 100 %% |" % (len(rows), tot))


if __name__ == "__main__":
    main()
