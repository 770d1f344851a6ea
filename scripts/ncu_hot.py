"""Top stall sites of a kernel from an .ncu-rep (SASS view of `ncu --page source --csv`): python scripts/ncu_hot.py rep [kernel_index] [top]"""
import csv, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and r:
        cur["rows"].append(r)
k = kernels[which]
h = k["hdr"]
ix = {n: i for i, n in enumerate(h)}
print(k["name"][:100])
tot = sum(int(r[ix["# Samples"]] or 0) for r in k["rows"])
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(int(r[ix[n]] or 0) for r in k["rows"]) for n in stalls}
print("samples", tot, " by reason:", ", ".join("%s %.1f%%" % (n[6:], 100.0 * v / tot) for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = k["rows"][i]
    s = int(r[ix["# Samples"]] or 0)
    why = sorted(((int(r[ix[n]] or 0), n[6:]) for n in stalls), reverse=True)[:2]
    print("%5d %5.1f%%  #%-5d %-60s %s" % (s, 100.0 * s / tot, i, r[ix["Source"]].strip()[:60], " ".join("%s:%d" % (n, v) for v, n in why if v)))
