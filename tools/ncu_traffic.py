"""Per-kernel DRAM traffic of an `ncu --set full` capture as JSON (what bench.py reads for roofline.traffic).
usage: python tools/ncu_traffic.py gpurun_out/prof.ncu-rep > profiles/rN_traffic.json"""
import csv
import hashlib
import json
import os
import re
import subprocess
import sys


def csrc_digest():
    """Same digest as bench.py: sha1 over csrc/*.cu, *.cuh - run this tool on the tree the capture was taken from."""
    h = hashlib.sha1()
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kuzushiji_vision_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
out = {"_csrc_digest": csrc_digest()}
for r in rows[2:]:
    name = re.sub(r"^(void )?(kzd::)?", "", r[col["Kernel Name"]].split("(")[0]).replace("kzd::", "")
    m = re.match(r"decode_infer_kernel<(\d+), (\w+), (\d), (\d), (\d)>", name)
    if m:  # ncu prints bool template arguments as 0/1
        tf = lambda g: "true" if m.group(g) == "1" else "false"
        name = "decode_infer_kernel<%s, %s, %s, %s, %s>" % (m.group(1), m.group(2), tf(3), tf(4), tf(5))
    rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * scale[units[col["dram__bytes_read.sum"]]]
    wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * scale[units[col["dram__bytes_write.sum"]]]
    us = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * scale[units[col["gpu__time_duration.sum"]]]
    if name not in out or not isinstance(out.get(name), dict) or us > out[name]["duration_us"]:  # several launches (NMS phases): keep the longest
        out[name] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr, "duration_us": us}
print(json.dumps(out, indent=1))
