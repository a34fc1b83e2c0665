"""Turn the ncu CSV of tools/ncu_target_big.py into profiles/r02_solve_traffic.json (what bench.py reports as roofline.traffic):
  python tools/ncu_traffic.py <ncu.csv> <target stdout log> [profiles/r02_solve_traffic.json]"""
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
csv_path, log_path = sys.argv[1], sys.argv[2]
out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "r02_solve_traffic.json")
m = re.search(r"PRODUCTS (\d+) CAMERAS (\d+) SOLVE_MS ([0-9.]+)", open(log_path).read())
products, cameras, solve_ms = int(m.group(1)), int(m.group(2)), float(m.group(3))
rows = [r for r in csv.reader(l for l in open(csv_path) if l.startswith('"'))]
hdr = rows[0]
iname, ival, iunit, ikern = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Kernel Name")
vals = {}
for r in rows[1:]:
    if "xm_solve_kernel" not in r[ikern]:
        continue
    v = float(r[ival].replace(",", ""))
    unit = r[iunit].lower()
    scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(unit, 1.0)
    vals[r[iname]] = v * scale
    kern = r[ikern]
total = vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
sys.path.insert(0, ROOT)
import bench  # noqa: E402
rec = {"cameras": cameras, "src_sha16": bench.source_hash(), "kernel": kern[:120], "products": products,
       "dram_bytes_read": vals["dram__bytes_read.sum"], "dram_bytes_write": vals["dram__bytes_write.sum"], "dram_bytes_per_product": total / products,
       "algorithmic_bytes_per_product": 72.0 * cameras ** 2 + 48.0 * cameras * 3, "kernel_ms_under_ncu": vals.get("gpu__time_duration.sum", 0) / 1e6,
       "source": os.path.basename(csv_path)}
rec["traffic_over_algorithmic"] = rec["dram_bytes_per_product"] / rec["algorithmic_bytes_per_product"]
try:
    doc = json.load(open(out_path))
except Exception:
    doc = {"captures": []}
doc["captures"] = [c for c in doc["captures"] if not (c["cameras"] == cameras and c.get("src_sha16") == rec["src_sha16"])] + [rec]
json.dump(doc, open(out_path, "w"), indent=1)
print(json.dumps(rec))
