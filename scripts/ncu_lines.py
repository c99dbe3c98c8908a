"""Aggregate warp-stall samples of an ncu report per CUDA source line.
usage: python scripts/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id",
                      f"::regex:{kern}:1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg, tot, fname = [], 0.0, ""
for r in rows:
    if r and r[0] == "File Name":
        fname = r[1].split("/")[-1]
    if len(r) > 5 and r[0].isdigit():
        try:
            v = float(r[4])
        except ValueError:
            continue
        tot += v
        agg.append((v, fname, r[0], r[1].strip()[:110]))
agg.sort(reverse=True)
print("total samples", tot)
for v, f, l, s in agg[:top]:
    print(f"{v / tot:6.1%} {f}:{l:>4} {s}")
