mkdir -p gpurun_out; rm -rf gpurun_out/*
for wl in fibinet xdeepfm deepfm; do
CTR_PROFILE_REGION=1 timeout -s KILL 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-secondary > gpurun_out/bench_ncu_$wl.log 2>&1; echo "ncu $wl exit $?"
done
ls -la gpurun_out
python - <<'PY'
import csv, collections
for wl in ("fibinet", "xdeepfm", "deepfm"):
    rows = list(csv.reader(open("gpurun_out/launches_%s.csv" % wl)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr): continue
        v = float(r[idx["Metric Value"]]); u = r[idx["Metric Unit"]]
        v = v / 1000.0 if u in ("ns", "nsecond") else (v * 1000.0 if u in ("ms", "msecond") else v)
        k = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("==", wl, "total us", round(tot), "launches", sum(v[0] for v in agg.values()))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print("  %-60s %5d %10.1f %5.1f%%" % (k[:60], v[0], v[1], 100 * v[1] / tot))
PY
