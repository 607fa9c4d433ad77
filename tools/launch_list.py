"""Summarise the ncu launch list of `bench.py --steps 2 ...` (the --metrics gpu__time_duration.sum,
dram__bytes_read.sum,dram__bytes_write.sum CSV): the kernels of the FIRST TIMED device step, i.e. the
launches between the first 256 MB L2-flush fill and the second one (warm-up steps are not flushed; the
second timed step is followed by the unflushed warm-ups of the e2e measurement).
Usage: python tools/launch_list.py launches.csv > profiles/rNN_bench_cfg2_launch_list.txt"""
import csv
import re
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ix = {k: i for i, k in enumerate(h)}
launches = OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) != len(h):
        continue
    d = launches.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]]})
    try:
        d[r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
    except ValueError:
        pass


def us(d):
    v, u = d.get("gpu__time_duration.sum", (0.0, "us"))
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)


def mb(d):
    t = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        v, u = d.get(k, (0.0, "byte"))
        t += v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    return t


seq = list(launches.values())
flush = [i for i, d in enumerate(seq) if "FillFunctor<float>" in d["name"] and mb(d) > 100.0]
if len(flush) < 2:
    raise SystemExit("fewer than two L2-flush fills in the launch list")
window = seq[flush[0] + 1:flush[1]]
agg = OrderedDict()
for d in window:
    name = re.sub(r"^void ", "", d["name"])
    name = re.sub(r"^tlod::", "", name)
    name = name.split("(")[0]
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += us(d)
    a[2] += mb(d)
total = sum(a[1] for a in agg.values())
mine = sum(a[0] for n, a in agg.items() if not n.startswith("at::"))
print("kernel                                                              count   total us   share DRAM MB/launch")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-66s %6d %10.1f %6.1f%% %14.2f" % (name[:66], a[0], a[1], 100 * a[1] / total, a[2] / a[0]))
n = sum(a[0] for a in agg.values())
print()
print("kernel time of the step, serialised and cold (without the flush): %.1f us; %d kernel launches" % (total, n))
print("launches of this library: %d; torch's own (cat / fill / add / mul of the autograd glue and the candidate "
      "assembly): %d" % (mine, n - mine))
