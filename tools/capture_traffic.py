#!/usr/bin/env python
"""profiles/k1_traffic.json from the ncu captures of tools/capture_profiles.sh: DRAM bytes of
sweep_tile_kernel per (map, source) pair, stamped with the hash of the kernel sources they were
captured from (bench.py reports roofline.traffic only while that hash matches).

    python tools/capture_traffic.py gpurun_out
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402

PAIRS = {"c2": 1184, "c2s": 1184, "c2d": 1184, "c4": 32768}


def main():
    d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
    out = {"_doc": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of sweep_tile_kernel<float> from one "
                   "`ncu --set full --clock-control none` capture per workload (tools/capture_profiles.sh), stored per "
                   "(map, source) pair; bench.py scales by the pairs of its launch and reports the figure only while "
                   "kernel_source_hash matches the sources it runs",
           "kernel_source_hash": kernel_source_hash()}
    for w, pairs in PAIRS.items():
        rep = os.path.join(d, f"k1_{w}.ncu-rep")
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        def get(name):
            i = hdr.index(name)
            v = float(vals[i].replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        tot = get("dram__bytes_read.sum") + get("dram__bytes_write.sum")
        out[w] = {"pairs_captured": pairs, "dram_bytes": int(tot), "dram_bytes_per_pair": tot / pairs,
                  "dram_bytes_read": int(get("dram__bytes_read.sum"))}
    json.dump(out, open(os.path.join(ROOT, "profiles", "k1_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
