"""profiles/ summaries of the memory-bound kernels from gpurun_out/ ncu outputs (run here, no GPU needed).

    python tools/summarize_mem.py <tag>        # reads gpurun_out/mem_<tag>.csv and gpurun_out/full_<tag>_*.ncu-rep

Writes profiles/<tag>_mem_kernels.md: per kernel of one eager step -- launches, device time, DRAM bytes read + written
(ncu dram__bytes_*.sum), achieved DRAM GB/s and the fraction of the measured copy bandwidth (MEASURED_PEAKS.json),
L2 bytes -- and profiles/<tag>_full_<kernel>.md for every --set full capture.
"""
import collections
import csv
import glob
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}.get(unit, 1e-3)


def mem_table(path, out, tag):
    with open(path) as f:
        rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
    per = collections.OrderedDict()     # (kernel, launch id) -> metrics
    for r in rows:
        key = (r["ID"], re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))
        per.setdefault(key, {})[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
    # one step = from one stem_im2col launch to the next
    ids = list(per.keys())
    stems = [i for i, k in enumerate(ids) if "stem_im2col" in k[1]]
    seq = ids[stems[0]:stems[1]] if len(stems) > 1 else ids
    agg = collections.OrderedDict()
    for k in seq:
        m = per[k]
        a = agg.setdefault(k[1], dict(n=0, us=0.0, rd=0.0, wr=0.0, l2=0.0))
        a["n"] += 1
        a["us"] += to_us(*m["gpu__time_duration.sum"])
        a["rd"] += to_bytes(*m["dram__bytes_read.sum"])
        a["wr"] += to_bytes(*m["dram__bytes_write.sum"])
        if "lts__t_bytes.sum" in m:
            a["l2"] += to_bytes(*m["lts__t_bytes.sum"])
    pk = peak_gbs()
    with open(out, "w") as f:
        f.write("# non-conv kernels of ONE eager bench step (%s): ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                "dram__bytes_write.sum,lts__t_bytes.sum --clock-control none\n" % tag)
        f.write("# cold-cache, serialised launches.  DRAM GB/s = (read + written bytes) / duration; frac = of the measured "
                "copy bandwidth %.0f GB/s (MEASURED_PEAKS.json)\n\n" % pk)
        f.write("| kernel | launches | us | DRAM read MB | DRAM write MB | DRAM GB/s | frac of HBM peak | L2 MB |\n|---|---:|---:|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            gbs = (a["rd"] + a["wr"]) / max(a["us"], 1e-9) / 1e3
            f.write("| %s | %d | %.1f | %.2f | %.2f | %.0f | %.3f | %.1f |\n" % (k, a["n"], a["us"], a["rd"] / 1e6, a["wr"] / 1e6, gbs, gbs / pk, a["l2"] / 1e6))
        f.write("\ntotal %.1f us over %d launches\n" % (sum(a["us"] for a in agg.values()), sum(a["n"] for a in agg.values())))


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        return False
    hdr, units = rows[0], rows[1]
    pk = peak_gbs()
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, one launch (%s)\n\n" % os.path.basename(rep))
        for r in rows[2:]:
            f.write("## %s\n\n" % re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]))
            vals = {}
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    vals[w] = (r[i], units[i])
                    f.write("- %s: %s %s\n" % (w, r[i], units[i]))
            try:
                us = to_us(*vals["gpu__time_duration.sum"])
                by = to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"])
                f.write("- **DRAM traffic %.2f MB in %.1f us = %.0f GB/s = %.3f of the measured %.0f GB/s**\n" % (by / 1e6, us, by / us / 1e3, by / us / 1e3 / pk, pk))
            except Exception:
                pass
            f.write("\n")
    return True


if __name__ == "__main__":
    tag = sys.argv[1]
    mem = os.path.join(ROOT, "gpurun_out", "mem_%s.csv" % tag)
    if os.path.exists(mem):
        mem_table(mem, os.path.join(ROOT, "profiles", "%s_mem_kernels.md" % tag), tag)
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "full_%s_*.ncu-rep" % tag))):
        name = os.path.basename(rep)[len("full_%s_" % tag):-len(".ncu-rep")]
        full(rep, os.path.join(ROOT, "profiles", "%s_full_%s.md" % (tag, name)))
