"""Turn ncu artefacts from gpurun_out/ into committed summaries under profiles/.

    python scripts/summarize_profile.py <tag> <launches.csv> <full.ncu-rep>

writes profiles/<tag>_launches.md (per-kernel launch count / total time / share of the step, from the
`--metrics gpu__time_duration.sum` pass) and profiles/<tag>_kernels.md (per captured launch: duration,
DRAM bytes, DRAM / tensor-pipe / issue utilisation, registers, top stall reason from the source page).
"""
import collections
import csv
import os
import subprocess
import sys

tag, launches_csv, rep = sys.argv[1], sys.argv[2], sys.argv[3]
out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
os.makedirs(out_dir, exist_ok=True)


def short(name):
    return name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")


if os.path.exists(launches_csv):
    rows = list(csv.reader(open(launches_csv)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        val = float(r[idx["Metric Value"]])
        unit = r[idx["Metric Unit"]]
        val = val / 1000.0 if unit in ("ns", "nsecond") else (val * 1000.0 if unit in ("ms", "msecond") else val)
        a = agg.setdefault(short(r[idx["Kernel Name"]]), [0, 0.0])
        a[0] += 1
        a[1] += val
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, tag + "_launches.md"), "w") as f:
        f.write("# %s — launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, "
                "serialised: compare SHARES)\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n" % tag)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (k[:70], v[0], v[1], 100 * v[1] / tot))
        f.write("\ntotal %.1f us over %d launches\n" % (tot, sum(v[0] for v in agg.values())))

if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
            ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %")]
    with open(os.path.join(out_dir, tag + "_kernels.md"), "w") as f:
        f.write("# %s — ncu --set full --clock-control none (one row per captured launch)\n\n" % tag)
        f.write("| kernel | " + " | ".join(c[1] for c in cols) + " |\n|---|" + "---:|" * len(cols) + "\n")
        for r in rows[2:]:
            vals = []
            for m, _ in cols:
                if m in idx and r[idx[m]] not in ("", "n/a"):
                    v = float(r[idx[m]].replace(",", ""))
                    vals.append(("%.1f %s" % (v, units[idx[m]])) if units[idx[m]] not in ("", "%") else "%.1f" % v)
                else:
                    vals.append("-")
            f.write("| `%s` | %s |\n" % (short(r[idx["Kernel Name"]])[:60], " | ".join(vals)))
print("wrote", out_dir)
