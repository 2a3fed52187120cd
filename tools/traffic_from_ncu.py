"""Per-kernel DRAM traffic and duration of ONE forward pass from an ncu metric table (the `--metrics ...` CSV that
tools/ncu_final.sh writes for `tools/prof_targets.py all 1024`): profiles/<tag>_forward_all_metrics.csv →
profiles/<tag>_traffic.json.  bench.py reads the newest *_traffic.json to fill `roofline.traffic`.
Usage: python tools/traffic_from_ncu.py profiles/r01_d_forward_all_metrics.csv [batch]"""
import collections
import csv
import json
import re
import sys

path = sys.argv[1]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
rows = list(csv.reader(open(path)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]
col = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Grid Size")}
launch = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    d = launch.setdefault(int(r[col["ID"]]), dict(kernel=re.sub(r"^void |<unnamed>::|unnamed>::|\(.*$|<.*$", "", r[col["Kernel Name"]])))
    d[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
# prof_targets.py "all": frontend launches, then embedding forwards; a forward runs from one stem launch to the next
ids = list(launch)
fe = [i for i in ids if launch[i]["kernel"].startswith("frontend")]
emb = [i for i in ids if not launch[i]["kernel"].startswith("frontend")]
stems = [n for n, i in enumerate(emb) if launch[i]["kernel"].startswith("stem")]
assert len(stems) >= 2, "need one complete forward pass (two stem launches) in the capture"
keep = fe[-1:] + emb[stems[0]:stems[1]]
out = collections.OrderedDict()
for i in keep:
    d = launch[i]
    k = out.setdefault(d["kernel"], dict(launches=0, dram_bytes=0.0, l2_bytes=0.0, duration_us=0.0))
    k["launches"] += 1
    k["dram_bytes"] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    k["l2_bytes"] += d.get("lts__t_bytes.sum", 0)
    k["duration_us"] += d.get("gpu__time_duration.sum", 0) / 1e3
for k in out.values():
    k["dram_bytes_per_launch"] = k["dram_bytes"] / k["launches"]
    k["duration_us"] = round(k["duration_us"], 2)
res = dict(source=path, batch=batch, note="ncu counters of one forward pass at this batch, plain launches, warm L2 between "
           "consecutive kernels as in the real schedule; durations are serialised ncu replays (shares only)", kernels=out)
dst = path.replace("_forward_all_metrics.csv", "_traffic.json")
json.dump(res, open(dst, "w"), indent=1)
print(json.dumps(res, indent=1))
