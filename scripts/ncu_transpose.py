"""`ncu --page raw --csv` of ONE launch (wide: one column per metric) -> tall `metric,value,unit` csv kept under
profiles/.  usage: ncu_transpose.py <raw.csv> <out.csv>"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "value", "unit"])
    for h, u, v in zip(hdr, units, vals):
        if h in ("ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device", "CC"):
            continue
        w.writerow([h, v, u])
