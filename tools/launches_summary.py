"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel shares + compact gz copy.
usage: python tools/launches_summary.py gpurun_out/X_launches.csv profiles/rNN_name "<command that was profiled>" """
import collections, csv, gzip, re, sys


def main(src, dst, cmd=""):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    with gzip.open(dst + "_launches.csv.gz", "wt") as gz:
        gz.write("id,kernel,duration_us\n")
        for row in csv.DictReader(lines):
            if row.get("Metric Name") != "gpu__time_duration.sum":
                continue
            k = row["Kernel Name"].split("(")[0]
            v = float(row["Metric Value"].replace(",", ""))
            u = row["Metric Unit"]
            v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
            agg[k][0] += 1
            agg[k][1] += v
            n += 1
            gz.write(f"{row['ID']},{k.replace(',', ';')},{v:.3f}\n")
    tot = sum(v[1] for v in agg.values())
    with open(dst + "_launches_summary.txt", "w") as f:
        f.write(f"# {cmd}\n# {n} launches, {tot / 1e3:.1f} ms total; ncu times are cold-cache and serialised: compare SHARES\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1] / tot * 100:6.2f}%  {v[1] / 1e3:9.2f} ms  n={v[0]:6d}  avg={v[1] / v[0]:8.1f} us  {k}\n")


if __name__ == "__main__":
    main(*sys.argv[1:4])
