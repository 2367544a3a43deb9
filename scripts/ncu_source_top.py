"""Top source lines of an ncu report by warp-stall samples:  python scripts/ncu_source_top.py X.ncu-rep [launch-id] [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
lid = sys.argv[2] if len(sys.argv) > 2 else None
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if lid is not None:
    cmd += ["--launch-skip", lid, "--launch-count", "1"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
data = []
for r in rows:
    if hdr is None:
        if "Source" in r and any("Sampl" in c for c in r):
            hdr = r
        continue
    if len(r) == len(hdr):
        data.append(r)
if hdr is None:
    print(out[:2000]); sys.exit(0)
si = hdr.index("Source")
cols = [i for i, c in enumerate(hdr) if c.startswith("# Warp Stall Sampling (All")] or [i for i, c in enumerate(hdr) if "Sampl" in c]
ci = cols[0]
def val(x):
    try: return float(x.replace(",", ""))
    except ValueError: return 0.0
tot = sum(val(r[ci]) for r in data)
print("# %s  column: %s  total samples %.0f" % (rep, hdr[ci], tot))
for r in sorted(data, key=lambda r: -val(r[ci]))[:n]:
    print("%7.0f %5.1f%%  %s" % (val(r[ci]), 100 * val(r[ci]) / max(tot, 1), r[si].strip()[:150]))
