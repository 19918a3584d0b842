"""Collect DRAM bytes per launch of the Gabor kernel from ncu CSV logs (one per BASELINE config) into
profiles/r02_k1_traffic.json, stamped with the hash of the kernel sources bench.py checks before quoting them.
    python tools/traffic_json.py gpurun_out/r02_traffic_c{1,2,3,4,5}.csv"""
import csv, io, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import k1_source_sha16

out = {}
for path in sys.argv[1:]:
    cfg = path.split("_c")[-1].split(".")[0]
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = {}
    for r in rows:
        if "k1_tc_kernel" not in r["Kernel Name"]:
            continue
        per.setdefault(r["ID"], {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    tot = []
    for d in per.values():
        b = 0.0
        for v, unit in d.values():
            b += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        tot.append(b)
    if tot:
        tot.sort()
        out[cfg] = {"dram_bytes_per_launch": tot[len(tot) // 2], "launches_seen": len(tot), "k1_source_sha16": k1_source_sha16(),
                    "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on bench.py --config {cfg} (median over the "
                              f"k1_tc_kernel launches; {os.path.basename(path)})"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_k1_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
