"""Top stall sites (SASS instructions) of one launch of an .ncu-rep, with the two instructions before each for context.
   python tools/ncu_hot.py <rep> <launch id> [n]"""
import csv
import io
import subprocess
import sys

rep, lid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", lid, "--launch-count", "1"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
lines = raw.splitlines()
print(lines[0][:200])
rd = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rd[0]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_")]
rows = []
tot = 0
for k, r in enumerate([r for r in rd[1:] if len(r) == len(hdr)]):
    try:
        s = int(r[ix["Warp Stall Sampling (All Samples)"]])
    except Exception:
        continue
    tot += s
    rows.append((s, k))
body = [r for r in rd[1:] if len(r) == len(hdr)]
inst_total = sum(int(r[ix["Instructions Executed"]]) for r in body if r[ix["Instructions Executed"]].isdigit())
print("total samples", tot, "warp instructions", inst_total)
for s, k in sorted(rows, key=lambda x: -x[0])[:n]:
    r = body[k]
    why = sorted(((int(r[ix[c]]), c[6:]) for c in stall_cols if r[ix[c]].isdigit() and int(r[ix[c]]) > 0), reverse=True)[:2]
    ctx = " <- ".join(body[j][ix["Source"]].strip()[:60] for j in range(k, max(k - 3, -1), -1))
    print("%5.1f%% exec=%-9s %-28s %s" % (100.0 * s / max(tot, 1), r[ix["Instructions Executed"]], ",".join("%s:%d" % (w, c) for c, w in why), ctx))
