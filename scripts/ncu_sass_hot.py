"""Hot SASS regions from an `ncu --page source --csv` export: prints instruction index, executed count, samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, ismp, iex, ithr = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
data = rows[2:]
tot = sum(int(r[iex]) for r in data); tots = sum(int(r[ismp]) for r in data)
print("total inst executed", tot, "samples", tots, "n sass", len(data))
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
for i, r in enumerate(data):
    ex = int(r[iex])
    if ex / tot >= thresh:
        print(f"{i:5d} ex={ex/tot*100:5.2f}% smp={int(r[ismp])/max(tots,1)*100:5.2f}% thr={r[ithr]:>5s}  {r[isrc].strip()[:90]}")
