"""profiles/r02_decode_traffic.json from an ncu launch list (the `roofline.traffic` of bench.py): the DRAM bytes of ONE
decode step = sum of dram__bytes_read + dram__bytes_write over the kernels of the last decode step in the list.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/launches.csv python tools/profile_step.py 64 3
    python tools/ncu_traffic.py gpurun_out/launches.csv <git-rev> > profiles/r02_decode_traffic.json
"""
import collections
import csv
import io
import json
import re
import sys

txt = open(sys.argv[1]).read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID","Process ID"'):])))
launch = collections.OrderedDict()
for r in rows:
    e = launch.setdefault(r["ID"], {"k": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void mq::", ""), "g": r["Grid Size"], "b": 0.0, "t": 0.0})
    v = float(r["Metric Value"].replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
    if r["Metric Name"].startswith("dram__bytes"):
        e["b"] += v * scale
    elif r["Metric Name"] == "gpu__time_duration.sum":
        e["t"] += v / 1000.0 if r["Metric Unit"] == "ns" else v
steps, cur = [], None
for e in launch.values():
    if e["k"].startswith("embed_kernel"):
        cur = {"grid": e["g"], "k": []}
        steps.append(cur)
    if cur is not None:
        cur["k"].append(e)
dec = [s for s in steps if s["grid"] == "(64, 1, 1)"]
last = dec[-1]
# algorithmic bytes of that step: 2 P_mm + sum ctx KV + B KV, ctx = 512 prompt + tokens decoded so far
L, H, I, V, qkv, kvb = 32, 4096, 14336, 128256, 6144, 131072
p_mm = L * (qkv * H + H * H + 3 * I * H) + V * H
n_dec = len(dec)
ctx = 512 + n_dec  # position of the token being decoded in the last profiled step (prefill produced token 1)
alg = 2.0 * p_mm + 64 * ctx * kvb + 64 * kvb
print(json.dumps({
    "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum launch list of tools/profile_step.py (B=64, Llama-3-8B); "
              "last decode step of the list, %d kernels" % len(last["k"]),
    "captured_on": sys.argv[2] if len(sys.argv) > 2 else None,
    "decode_step_dram_bytes": sum(e["b"] for e in last["k"]),
    "decode_step_algorithmic_bytes_same_step": alg,
    "decode_step_kernels": len(last["k"]),
    "decode_step_ncu_time_us_serialised": sum(e["t"] for e in last["k"]),
}, indent=1))
