#!/usr/bin/env python
"""profiles/r2_traffic.json from the `ncu --set full` reports of tools/gpu_evidence_r2.sh (read here, on the CPU box).

    python tools/traffic_json.py <commit> c2=gpurun_out/r2_scan_c2_full.ncu-rep:1000000000:profiles/r2_scan_c2_full_ncu.txt ...

Per configuration: dram__bytes_read.sum / dram__bytes_write.sum summed over the captured launches of ONE step (one launch
for the scan kernels, the two radix passes for c4s), the kernel names and the rows the step scanned."""
import csv
import json
import subprocess
import sys

UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def launches(rep):
    if rep.endswith(".csv"):     # `ncu -i <rep> --page raw --csv` made on the GPU box (the report itself stayed there)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    out = []
    for vals in r[2:]:
        m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
        rd = float(m["dram__bytes_read.sum"][0].replace(",", "")) * UNIT[m["dram__bytes_read.sum"][1]]
        wr = float(m["dram__bytes_write.sum"][0].replace(",", "")) * UNIT[m["dram__bytes_write.sum"][1]]
        out.append((m["Kernel Name"][0], rd, wr))
    return out


def main():
    commit = sys.argv[1]
    doc = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel(s) of ONE step from one `ncu --set full "
                       "--clock-control none` capture per configuration (tools/gpu_evidence_r2.sh); bench.py reports it as "
                       "roofline.traffic scaled to the rows of the run"}
    for spec in sys.argv[2:]:
        cfg, rest = spec.split("=", 1)
        rep, rows, src = rest.split(":")
        ls = launches(rep)
        doc[cfg] = {"rows": int(rows), "dram_read_bytes": sum(l[1] for l in ls), "dram_write_bytes": sum(l[2] for l in ls),
                    "kernel": " + ".join(l[0].split("(")[0] for l in ls), "source": src, "commit": commit}
    json.dump(doc, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
