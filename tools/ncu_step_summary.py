"""Summarise an `ncu --csv` launch list (metrics per kernel launch) into per-kernel-family totals.
usage: python tools/ncu_step_summary.py file.csv [max rows to print]"""
import csv, collections, io, sys


def load(path):
    txt = open(path).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    per = collections.OrderedDict()
    for r in rows:
        key = (int(r["ID"]), r["Kernel Name"], r["Grid Size"])
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        v *= {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "ms": 1e6, "us": 1e3, "msecond": 1e6, "usecond": 1e3, "second": 1e9}.get(u, 1)
        per.setdefault(key, {})[r["Metric Name"]] = v
    return per


def main():
    per = load(sys.argv[1]); show = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    tot = collections.defaultdict(lambda: collections.defaultdict(float))
    for n, (k, m) in enumerate(per.items()):
        name = k[1].split("(")[0].replace("void ", "")
        for a, b in m.items():
            if not a.endswith(".pct") and "pct_of_peak" not in a:
                tot[name][a] += b
        tot[name]["launches"] += 1
        if n < show:
            print(k[0], name[:40], k[2], {a: round(b / (1e6 if "bytes" in a else 1), 2) for a, b in m.items()})
    for name, m in tot.items():
        line = "%-44s launches %3d" % (name[:44], m["launches"])
        if "gpu__time_duration.sum" in m: line += "  time %.1f us" % (m["gpu__time_duration.sum"] / 1e3)
        if "dram__bytes_read.sum" in m: line += "  dram read %.1f MB write %.1f MB" % (m["dram__bytes_read.sum"] / 1e6, m["dram__bytes_write.sum"] / 1e6)
        if "lts__t_sectors_srcunit_tex_op_read.sum" in m:
            line += "  L2<->SM read %.1f MB write %.1f MB" % (m["lts__t_sectors_srcunit_tex_op_read.sum"] * 32 / 1e6, m["lts__t_sectors_srcunit_tex_op_write.sum"] * 32 / 1e6)
        if "smsp__inst_executed.sum" in m: line += "  warp-inst %.2f M" % (m["smsp__inst_executed.sum"] / 1e6)
        print(line)
    if any("dram__bytes_read.sum" in m for m in tot.values()):
        print("total dram: %.1f MB" % (sum(m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0) for m in tot.values()) / 1e6))


if __name__ == "__main__":
    main()
