import csv, sys
path = sys.argv[1]
rows = list(csv.reader(open(path)))
# split into kernels
kernels = []; cur = None; header = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; kernels.append(cur); continue
    if r and r[0] == "Address": header = r; cur["header"] = r; continue
    if cur is not None and r: cur["rows"].append(r)
def col(k, name): return k["header"].index(name)
best = max(kernels, key=lambda k: sum(int(r[col(k, "# Samples")] or 0) for r in k["rows"]))
k = best
iS, iE, iSrc = col(k, "# Samples"), col(k, "Instructions Executed"), col(k, "Source")
tot = sum(int(r[iS] or 0) for r in k["rows"]); totE = sum(int(r[iE] or 0) for r in k["rows"])
print("kernels", len(kernels), "rows", len(k["rows"]), "samples", tot, "inst", totE)
stall_cols = [i for i, h in enumerate(k["header"]) if h.startswith("stall_") or "Stall" in h]
print([k["header"][i] for i in range(len(k["header"]))][:80])
mode = sys.argv[2] if len(sys.argv) > 2 else "top"
if mode == "top":
    idx = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][iS] or 0))[:70]
    for i in sorted(idx):
        r = k["rows"][i]
        print(i, r[iS], r[iE], r[iSrc].strip()[:70])
elif mode == "regions":
    step = int(sys.argv[3]) if len(sys.argv) > 3 else 250
    for a in range(0, len(k["rows"]), step):
        s = sum(int(r[iS] or 0) for r in k["rows"][a:a + step]); e = sum(int(r[iE] or 0) for r in k["rows"][a:a + step])
        print(f"{a:6d} samples {100 * s / tot:5.1f}%  inst {100 * e / totE:5.1f}%")
elif mode == "range":
    a, b = int(sys.argv[3]), int(sys.argv[4])
    hdr = k["header"]
    want = [i for i, h in enumerate(hdr) if h in ("stall_barrier", "stall_short_sb", "stall_wait", "stall_long_sb", "stall_math", "stall_branch_resolving", "stall_dispatch", "stall_mio", "stall_no_inst", "stall_selected", "stall_not_selected", "stall_lg")]
    print([hdr[i] for i in want])
    for i in range(a, b):
        r = k["rows"][i]
        print(i, r[iS], r[iE], r[iSrc].strip()[:60], [r[j] for j in want])
