"""Top SASS lines by stall samples from an `ncu --page source --csv --print-source sass` export.
Usage: python tools/ncu_src.py file.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr_i]
col = {c: i for i, c in enumerate(h)}
body = [r for r in rows[hdr_i + 1:] if len(r) == len(h) and r[0] != "Address"]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
inst = sum(int(r[col["Instructions Executed"]] or 0) for r in body)
print("kernel:", rows[0][1][:110]); print("total samples", tot, "warp instr", inst, "sass lines", len(body))
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
agg = {s: sum(int(r[col[s]] or 0) for r in body) for s in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    st = {s[6:]: int(r[col[s]] or 0) for s in stalls if int(r[col[s]] or 0)}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{i:5d} {int(r[col['# Samples']]):6d} {int(r[col['Instructions Executed']] or 0):9d}  {r[col['Source']][:90]:90s} {st}")
