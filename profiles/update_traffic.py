#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report of the bench command (profiles/capture.sh): DRAM
bytes per launch of the scan kernel, tied to the commit the capture was made from.
    python profiles/update_traffic.py gpurun_out/<tag>.ncu-rep <db_descriptors> <query_descriptors>"""
import csv
import io
import json
import os
import subprocess
import sys

rep, n_db, n_q = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"],
                                                   capture_output=True, text=True).stdout)))
h, units = rows[0], rows[1]
c = {n: i for i, n in enumerate(h)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = []
for r in rows[2:]:
    if "imi_scan_kernel" not in r[c["Kernel Name"]]:
        continue
    rd = float(r[c["dram__bytes_read.sum"]]) * scale[units[c["dram__bytes_read.sum"]]]
    wr = float(r[c["dram__bytes_write.sum"]]) * scale[units[c["dram__bytes_write.sum"]]]
    out.append({"db_descriptors": n_db, "query_descriptors": n_q, "dram_bytes_per_launch": int(rd + wr),
                "dram_bytes_read": rd, "dram_bytes_write": wr,
                "gpu_time_us": float(r[c["gpu__time_duration.sum"]]),
                "source": f"ncu --set full --clock-control none of `python bench.py` ({os.path.basename(rep)})",
                "commit": subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True,
                                         text=True).stdout.strip(),
                "note": "entries are stored padded to 48 B (44 B algorithmic)"})
json.dump({"imi_scan_kernel": out[:1]}, open(os.path.join(os.path.dirname(__file__), "traffic.json"), "w"), indent=1)
print(out[:1])
