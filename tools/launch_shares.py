"""Summarise an ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_*]) into per-kernel shares of one
decode step and one prefill pass.  usage: python tools/launch_shares.py launches.csv [title] > profiles/xxx.txt"""
import collections
import csv
import io
import re
import sys

txt = open(sys.argv[1]).read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID","Process ID"'):])))
title = sys.argv[2] if len(sys.argv) > 2 else ""


def short(n):
    return re.sub(r"\(.*", "", n).replace("void mq::", "").replace("mq::", "")


# one record per launch id, possibly several metrics
launch = collections.OrderedDict()
for r in rows:
    e = launch.setdefault(r["ID"], {"k": short(r["Kernel Name"]), "g": r["Grid Size"], "t": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if r["Metric Name"] == "gpu__time_duration.sum":
        e["t"] = v / 1000.0 if u == "ns" else v
    elif r["Metric Name"] == "dram__bytes_read.sum":
        e["rd"] = v * scale
    elif r["Metric Name"] == "dram__bytes_write.sum":
        e["wr"] = v * scale
seqs, cur = [], None
for e in launch.values():
    if e["k"].startswith("embed_kernel"):
        cur = {"grid": e["g"], "k": []}
        seqs.append(cur)
    if cur is not None:
        cur["k"].append(e)
print("#", title)
print("# passes (grid of embed, kernels, total ms):",
      [(s["grid"], len(s["k"]), round(sum(x["t"] for x in s["k"]) / 1000, 2)) for s in seqs])


def summarize(s, name):
    agg = collections.OrderedDict()
    for e in s["k"]:
        a = agg.setdefault((e["k"], e["g"]), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += e["t"]
        a[2] += e["rd"] + e["wr"]
    tot = sum(a[1] for a in agg.values())
    byt = sum(a[2] for a in agg.values())
    print("== %s: total %.1f us over %d kernels, dram traffic %.3f GB (ncu, cold-cache serialized: compare SHARES)" %
          (name, tot, len(s["k"]), byt / 1e9))
    for (k, g), (n, t, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("  %-46s grid %-16s n=%3d  sum %9.1f us  avg %7.2f us  %5.1f%%  dram %8.1f MB" %
              (k, g, n, t, t / n, 100 * t / tot, b / 1e6))


dec = [s for s in seqs if s["grid"] == "(64, 1, 1)"]
pre = [s for s in seqs if s["grid"] != "(64, 1, 1)"]
if dec:
    summarize(dec[-1], "decode step (B=64, ctx~516)")
if pre:
    summarize(pre[0], "prefill pass " + pre[0]["grid"])
