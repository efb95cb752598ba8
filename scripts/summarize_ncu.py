"""Turns an .ncu-rep into the compact text summary kept under profiles/ (run on the CPU box).

    python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_fused_c2.md "title"
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
]


def main():
    rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none --import-source on)", ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines += [f"## {name[:150]}", "", "| metric | value | unit |", "|---|---|---|"]
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"| {w} | {r[i]} | {units[i]} |")
        lines.append("")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = [r for r in csv.reader(io.StringIO(src))]
    # one or more kernels: tables start with a "Kernel Name" row followed by a header row
    i = 0
    while i < len(srows):
        if srows[i] and srows[i][0] == "Kernel Name":
            kname = srows[i][1]
            h = srows[i + 1]
            j = i + 2
            data = []
            while j < len(srows) and not (srows[j] and srows[j][0] == "Kernel Name"):
                if len(srows[j]) == len(h):
                    data.append(srows[j])
                j += 1
            seen, uniq = set(), []
            ia = h.index("Address")
            for r in data:
                if r[ia] not in seen:
                    seen.add(r[ia])
                    uniq.append(r)
            isamp = h.index("# Samples")
            stall = [k for k, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
            tot = sum(int(float(r[isamp] or 0)) for r in uniq) or 1
            agg = sorted(((sum(int(float(r[k] or 0)) for r in uniq), h[k]) for k in stall), reverse=True)[:8]
            lines += [f"### warp-state samples: {kname[:120]}", "", "| stall reason | share |", "|---|---|"]
            lines += [f"| {n} | {100.0 * v / tot:.1f} % |" for v, n in agg]
            isrc = h.index("Source")
            top = sorted(uniq, key=lambda r: -int(float(r[isamp] or 0)))[:8]
            lines += ["", "| hottest SASS | samples |", "|---|---|"]
            lines += [f"| `{r[isrc][:70]}` | {r[isamp]} |" for r in top]
            lines.append("")
            i = j
        else:
            i += 1
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
