"""profiles/*_fused_kernel_hot_lines.txt from an ncu source-page CSV (sass): python scratch/hot_lines.py src.csv out.txt NQP 'title'"""
import csv, sys
path, out, nqp, title = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
rows = list(csv.reader(open(path)))
kernels = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; kernels.append(cur); continue
    if r and r[0] == "Address": cur["header"] = r; continue
    if cur is not None and r: cur["rows"].append(r)
k = max(kernels, key=lambda k: sum(int(r[k["header"].index("# Samples")] or 0) for r in k["rows"]))
H = k["header"]; R = k["rows"]
def I(name): return H.index(name)
def num(r, i):
    try: return float(r[i] or 0)
    except ValueError: return 0.0
iS, iE, iW, iWi, iSrc = I("# Samples"), I("Instructions Executed"), I("L1 Wavefronts Shared"), I("L1 Wavefronts Shared Ideal"), I("Source")
stall = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
nlaunch = sum(1 for kk in kernels if kk["name"] == k["name"])
tS = sum(num(r, iS) for r in R); tE = sum(num(r, iE) for r in R); tW = sum(num(r, iW) for r in R); tWi = sum(num(r, iWi) for r in R)
with open(out, "w") as f:
    f.write(f"# {k['name']}  ({title})\n")
    f.write(f"# SASS instructions {len(R)}; per QP: {tE / nqp:.0f} warp instructions, {tW / nqp:.0f} shared-memory wavefronts "
            f"({100 * (tW - tWi) / max(tW, 1):.1f} % bank-conflict replays)\n")
    f.write("# stall reasons over the kernel (share of warp-stall samples):\n")
    tot = {H[i]: sum(num(r, i) for r in R) for i in stall}
    ts = sum(tot.values())
    for h, v in sorted(tot.items()):
        if v / ts >= 0.01: f.write(f"#   {h:28s} {100 * v / ts:5.1f}%\n")
    f.write("# regions of 100 SASS instructions: start index, shared wavefronts %, replays % of region, executed instructions %, samples %\n")
    for a in range(0, len(R), 100):
        seg = R[a:a + 100]
        s = sum(num(r, iS) for r in seg); e = sum(num(r, iE) for r in seg); w = sum(num(r, iW) for r in seg); wi = sum(num(r, iWi) for r in seg)
        if s / tS >= 0.005 or e / tE >= 0.005:
            f.write(f"  {a:6d}  {100 * w / max(tW, 1):5.1f}%  {100 * (w - wi) / max(w, 1):5.1f}%  {100 * e / tE:5.1f}%  {100 * s / tS:5.1f}%\n")
    f.write("# single instructions with >= 1 % of the samples: index, samples %, executed, SASS, dominant stall\n")
    for i, r in enumerate(R):
        if num(r, iS) / tS >= 0.01:
            dom = max(stall, key=lambda j: num(r, j))
            f.write(f"  {i:6d}  {100 * num(r, iS) / tS:5.1f}%  {int(num(r, iE)):10d}  {r[iSrc].strip()[:60]:60s}  {H[dom]}\n")
print(open(out).read())
