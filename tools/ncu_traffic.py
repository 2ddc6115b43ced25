"""profiles/ncu_traffic.json from an `ncu --set full` report of ONE remesh: dram__bytes_read.sum / dram__bytes_write.sum per kernel,
summed over that kernel's launches.  usage: python tools/ncu_traffic.py <report.ncu-rep> <workload name> [doc string]"""
import collections, csv, io, json, pathlib, subprocess, sys

rep, workload = sys.argv[1], sys.argv[2]
doc = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = collections.OrderedDict()
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0]
    e = acc.setdefault(name, {"launches": 0, "dram_bytes_read": 0.0, "dram_bytes_write": 0.0})
    e["launches"] += 1
    for key, col in (("dram_bytes_read", "dram__bytes_read.sum"), ("dram_bytes_write", "dram__bytes_write.sum")):
        e[key] += float(d[col].replace(",", "")) * scale[u[col]]
path = pathlib.Path(__file__).resolve().parent.parent / "profiles" / "ncu_traffic.json"
cur = json.loads(path.read_text()) if path.exists() else {}
if doc:
    cur["_doc"] = doc
cur[workload] = acc
path.write_text(json.dumps(cur, indent=1))
print(json.dumps(acc, indent=1))
