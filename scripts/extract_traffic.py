"""DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the bench's dominant kernels,
from `ncu --set full` captures of the bench command itself -> profiles/r01_traffic.json (read by bench.py).

    python scripts/extract_traffic.py gpurun_out/full_ba.ncu-rep gpurun_out/full_tc.ncu-rep
"""
import csv
import json
import subprocess
import sys

OUT = "profiles/r02_traffic.json"
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(paths):
    out = {}
    for path in paths:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("osfm::", "")
            b = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
            e = out.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "source": path.split("/")[-1],
                                      "ncu_time_unit": units[it], "ncu_time": 0.0})
            e["launches"] += 1
            e["dram_bytes"] += b
            e["ncu_time"] += float(r[it])
    for e in out.values():
        e["dram_bytes_per_launch"] = e["dram_bytes"] / e["launches"]
        e["ncu_time_per_launch"] = e.pop("ncu_time") / e["launches"]
        del e["dram_bytes"]
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main(sys.argv[1:])
