"""Warp-stall samples per CUDA source line from an .ncu-rep captured with --set full --import-source on (code built with -lineinfo).
usage: python scripts/ncu_hotspots.py report.ncu-rep "header line" [top=40] > profiles/<name>_source_hotspots.txt"""
import csv, io, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows, fname = [], "?"
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0].isdigit() and r[2] == "-":
        try:
            rows.append((int(r[6]), int(r[7]), fname, int(r[0]), r[1].strip()))
        except ValueError:
            pass
total = sum(x[0] for x in rows)
print(title)
print(f"warp-stall samples per source line (top {top} of {total} samples)")
for s, inst, f, ln, src in sorted(rows, reverse=True)[:top]:
    print(f"{s:6d} {100.0 * s / max(1, total):5.1f}% inst={inst:9d} {f}:{ln}: {src[:150]}")
