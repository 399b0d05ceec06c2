"""roll an .ncu-rep source page up per CUDA source line: share of instructions executed and of stall samples"""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
i = next(i for i, r in enumerate(rows) if len(r) > 10 and 'Instructions Executed' in r)
hdr = rows[i]
li, ie, ss = hdr.index('Line No'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
agg = {}
for r in rows[i + 1:]:
    if len(r) <= ie: continue
    try:
        a = agg.setdefault(int(r[li]), [0, 0, r[1]])
        a[0] += int(r[ie] or 0); a[1] += int(r[ss] or 0)
    except ValueError:
        pass
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print("total warp instructions %d, stall samples %d" % (ti, ts))
for k, a in sorted(agg.items(), key=lambda kv: -(kv[1][0] / ti + kv[1][1] / ts))[:top]:
    print('%5d %6.2f%% inst %6.2f%% stall | %s' % (k, 100 * a[0] / ti, 100 * a[1] / ts, a[2][:110]))
