"""Turn gpurun_out/ ncu outputs into the committed summaries under profiles/.

    python tools/summarize_profiles.py <launches.csv> <full.ncu-rep> <tag>
"""
import collections
import csv
import re
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(path, out):
    with open(path) as f:
        rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
    stem = [i for i, r in enumerate(rows) if "stem_im2col" in r["Kernel Name"] or "stem_kernel" in r["Kernel Name"]]
    seq = rows[stem[0]:stem[1]] if len(stem) > 1 else rows
    agg = collections.OrderedDict()
    tot = 0.0
    for r in seq:
        n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        if n.startswith("at::"):
            n = "torch::" + n[4:44]
        v = float(r["Metric Value"].replace(",", "")) / 1e3
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(out, "w") as f:
        f.write("# kernel launches of ONE bench step (ncu --metrics gpu__time_duration.sum --clock-control none;\n"
                "# cold-cache, serialised: compare SHARES, not absolutes)\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.1f%% |\n" % (k, n, v, 100 * v / tot))
        f.write("\ntotal %.1f us over %d launches\n" % (tot, len(seq)))


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on (one row per captured launch)\n\n")
        for r in rows[2:]:
            f.write("## %s grid %s\n\n" % (re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]), r[hdr.index("Grid Size")]))
            for w, i in idx[2:]:
                f.write("- %s: %s %s\n" % (w, r[i], units[i]))
            f.write("\n")


if __name__ == "__main__":
    tag = sys.argv[3]
    launches(sys.argv[1], "profiles/%s_launches.md" % tag)
    full(sys.argv[2], "profiles/%s_conv_tc_full.md" % tag)
