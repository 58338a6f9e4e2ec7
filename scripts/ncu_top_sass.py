"""Top stalled SASS instructions of the n-th capture of a kernel in an `ncu --page source --csv` dump.
usage: ncu_top_sass.py <source.csv> <kernel substring> [occurrence=0] [top=25]"""
import csv, sys
path, pat = sys.argv[1], sys.argv[2]
occ = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
csv.field_size_limit(1 << 30)
rows = list(csv.reader(open(path)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sel = [i for i in starts if pat in rows[i][1]]
s = sel[occ]
e = min([i for i in starts if i > s] + [len(rows)])
hdr = rows[s + 1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[s + 2:e]
tot = sum(int(r[ix["# Samples"]]) for r in body if len(r) > ix["# Samples"])
print("kernel:", rows[s][1][:90], " instructions:", len(body), " samples:", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in body) for h in stall_cols}
print("stall totals:", ", ".join("%s=%d" % (k[6:], v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[:top]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[ix[h]]), h[6:]) for h in stall_cols), reverse=True)[:2]
    print("%5d %6s %5.1f%%  %-70s %s" % (i, r[ix["# Samples"]], 100.0 * int(r[ix["# Samples"]]) / max(tot, 1), r[ix["Source"]].strip()[:70],
                                       " ".join("%s=%d" % (n, c) for c, n in st if c)))
