"""Per CUDA source line: instructions executed + stall samples, from an ncu report.
usage: ncu_hotlines.py <file.ncu-rep> <kernel regex> <launch-skip among matches> [top]"""
import csv, io, subprocess, sys
rep, kre, skip = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:" + kre, "-s", skip, "-c", "1"],
                     capture_output=True, text=True).stdout
cur = None; hdr = None; data = []
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] in ("File Path", "File Name"): cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name": print("kernel:", r[1]); continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "": continue
    try:
        i_inst = hdr.index("Instructions Executed"); i_s = hdr.index("# Samples"); i_thr = hdr.index("Thread Instructions Executed")
        data.append((int(r[i_inst]), int(r[i_s]), int(r[i_thr]), cur, r[0], r[1].strip()[:110]))
    except ValueError:
        pass
ti = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print("total warp-inst %d  samples %d" % (ti, ts))
for d in sorted(data, key=lambda d: -d[0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  thr/inst %4.1f  %s:%s  %s" % (100.0 * d[0] / max(ti, 1), 100.0 * d[1] / max(ts, 1), d[2] / max(d[0], 1), d[3], d[4], d[5]))
