"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file = None
data = []
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] in ("Function Name",): continue
    if r[0] != "" and r[0].isdigit():
        try:
            s = float(r[4]); n = float(r[7])
        except ValueError:
            continue
        data.append((s, n, cur_file, int(r[0]), r[1].strip()))
tot = sum(d[0] for d in data)
print("total samples", tot)
for s, n, f, ln, src in sorted(data, reverse=True)[:top]:
    print("%6.2f%% %12d  %s:%d | %s" % (100 * s / tot, n, f, ln, src[:120]))
