"""Per-CUDA-source-line instruction and stall-sample shares from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys, subprocess, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; out = []; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1]; continue
    if r and r[0] == "Line No": hdr = r; ci = {h: i for i, h in enumerate(hdr)}; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        def g(k):
            try: return int(float(r[ci[k]]))
            except Exception: return 0
        out.append((cur.split('/')[-1], int(r[0]), r[1].strip()[:100], g("# Samples"), g("Instructions Executed")))
tot = sum(o[4] for o in out) or 1; ts = sum(o[3] for o in out) or 1
print("total warp-inst", tot, "samples", ts)
for o in sorted(out, key=lambda o: -o[4])[:top]:
    print("%-14s %4d  inst %5.1f%%  samp %5.1f%%  %s" % (o[0], o[1], 100 * o[4] / tot, 100 * o[3] / ts, o[2]))
