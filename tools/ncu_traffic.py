"""profiles/traffic.json from `ncu --set full` captures: dram__bytes_read.sum + dram__bytes_write.sum per launch of every
captured kernel (bench.py reads it into roofline.traffic).  python tools/ncu_traffic.py raw1.csv [raw2.csv ...]
(each csv = `ncu -i x.ncu-rep --page raw --csv`)."""
import csv
import json
import os
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
units = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, unit = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[2:]:
        raw = r[col["Kernel Name"]]
        name = re.sub(r"<.*", "", re.sub(r"^void ", "", raw)).split("(")[0].split("::")[-1]
        if "link_heads" in name and re.search(r"<\s*(\(int\))?\d+\s*,\s*(\(bool\))?(1|true)\s*>", raw):
            name += "_zb"           # the per-row-offset launch over the non-empty links: not the main launch
        tot = sum(float(r[col[k]]) * units[unit[col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        per.setdefault(name, []).append((tot, float(r[col["gpu__time_duration.sum"]])))
    for name, v in per.items():
        out[name] = int(sum(t for t, _ in v) / len(v))
        print("%-40s %d launch(es): %.1f MB DRAM per launch, %.1f us" % (name, len(v), out[name] / 1e6, sum(d for _, d in v) / len(v)))
path = os.path.join(REPO, "profiles", "traffic.json")
old = json.load(open(path)) if os.path.exists(path) else {}
old.update(out)
json.dump(old, open(path, "w"), indent=1, sort_keys=True)
