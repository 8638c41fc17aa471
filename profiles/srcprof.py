"""Summarise an ncu report per CUDA source line:  python profiles/srcprof.py rep.ncu-rep k_rad [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
data, fname, hdr = [], "", None
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); ns = hdr.index("# Samples"); continue
    if hdr and r and r[0] not in ("", "Function Name") and len(r) > ie:
        try: data.append((int(r[ie]), int(r[ns]), fname, r[0], r[1].strip()[:100]))
        except ValueError: pass
ti = sum(d[0] for d in data) or 1; ts = sum(d[1] for d in data) or 1
print(f"total warp-instructions {ti}  samples {ts}")
for d in sorted(data, reverse=True)[:top]:
    print(f"{d[0]/ti*100:5.1f}% inst {d[1]/ts*100:5.1f}% smp  {d[2]}:{d[3]:>4}  {d[4]}")
# aggregate by file
agg = {}
for d in data:
    key = d[2]
    agg.setdefault(key, [0, 0]); agg[key][0] += d[0]; agg[key][1] += d[1]
print("by file:", {k: (round(v[0]/ti*100,1), round(v[1]/ts*100,1)) for k, v in agg.items()})
if len(sys.argv) > 4:
    bounds = [int(x) for x in sys.argv[4].split(",")]
    reg = {}
    for d in data:
        if d[2] != "we_kernels.cuh": continue
        ln = int(d[3]); k = max([b for b in bounds if b <= ln], default=0)
        reg.setdefault(k, [0, 0]); reg[k][0] += d[0]; reg[k][1] += d[1]
    for k in sorted(reg): print(f"  we_kernels.cuh from line {k}: {reg[k][0]/ti*100:5.1f}% inst {reg[k][1]/ts*100:5.1f}% smp")
