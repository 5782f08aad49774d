"""Summarise an .ncu-rep (read here, no GPU needed) into the text kept under profiles/.

    python scripts/summarize_ncu.py gpurun_out/prof_tc.ncu-rep > profiles/r01_ncu_tc.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "smsp__inst_executed.sum",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# %s" % path)
    for r in rows[2:]:
        print("\n## kernel: %s" % r[hdr.index("Kernel Name")][:110])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-70s %s %s" % (k, r[i], units[i]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    # one table per kernel, separated by "Kernel Name" rows
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1]
            hdr = rows[i + 1]
            j = i + 2
            data = []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                if len(rows[j]) == len(hdr):
                    data.append(rows[j])
                j += 1
            isamp, isrc, iex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
            stall = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            tot = sum(int(r[isamp] or 0) for r in data)
            print("\n## source hot spots: %s  (%d SASS instructions, %d samples)" % (name[:80], len(data), tot))
            agg = {}
            for r in data:
                for k in stall:
                    agg[hdr[k]] = agg.get(hdr[k], 0) + int(r[k] or 0)
            print("  stall totals: " + ", ".join("%s=%d" % kv for kv in sorted(agg.items(), key=lambda x: -x[1])[:8]))
            for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:12]:
                st = sorted(((hdr[k], int(r[k] or 0)) for k in stall), key=lambda x: -x[1])[:2]
                print("  %6s samples  %10s exec  %-60s %s" % (r[isamp], r[iex], r[isrc][:60], st))
            i = j
        else:
            i += 1


if __name__ == "__main__":
    main(sys.argv[1])
