"""Summarise an `ncu --page source --print-source cuda,sass --csv` export: per CUDA source line,
instructions executed and stall samples.  usage: ncu_hotlines.py file.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; data = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "": continue
    try:
        i_inst = hdr.index("Instructions Executed"); i_s = hdr.index("# Samples"); i_thr = hdr.index("Thread Instructions Executed")
        data.append((int(r[i_inst]), int(r[i_s]), int(r[i_thr]), cur, r[0], r[1].strip()[:100]))
    except ValueError:
        pass
ti = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print("total warp-inst %d  samples %d" % (ti, ts))
for d in sorted(data, key=lambda d: -d[1])[:top]:
    print("%5.1f%% smp %5.1f%% inst  thr/inst %4.1f  %s:%s  %s" % (100.0 * d[1] / ts, 100.0 * d[0] / ti, d[2] / max(d[0], 1), d[3], d[4], d[5]))
