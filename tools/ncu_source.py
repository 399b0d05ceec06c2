"""per-source-line roll-up of an .ncu-rep source page (instructions executed / stall samples)"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
for i, r in enumerate(rows):
    if 'Source' in r and 'Instructions Executed' in r:
        hdr, start = r, i + 1
        break
print(hdr)
