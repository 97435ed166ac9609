"""Aggregates the `ncu --page source --csv --print-source cuda,sass` dump of ONE kernel by CUDA source line: samples and dominant stall reasons.
usage: python tools/ncu_src.py dump.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
# the dump holds a CUDA-C table (one row per source line) first; find the header
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
out = []
for h in hi:
    hdr = rows[h]
    fname = rows[h - 2][1].split("/")[-1] if h >= 2 else "?"
    stall_cols = [(i, n) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
    ci_samples = hdr.index("# Samples")
    ci_inst = hdr.index("Instructions Executed")
    for r in rows[h + 1:]:
        if r and r[0] == "File Path":
            break
        if not r or not r[0].isdigit():
            continue
        try:
            smp = float(r[ci_samples] or 0)
        except ValueError:
            continue
        st = sorted(((float(r[i] or 0), n) for i, n in stall_cols), reverse=True)[:3]
        out.append((smp, fname + ":" + r[0], r[1].strip()[:100], float(r[ci_inst] or 0), st))
tot = sum(o[0] for o in out)
print("total samples", tot)
for smp, ln, src, inst, st in sorted(out, reverse=True)[:top]:
    print("%5.1f%% %-20s inst %9.0f | %-100s | %s" % (100 * smp / max(tot, 1), ln, inst, src, " ".join("%s=%.0f" % (n[6:], v) for v, n in st if v > 0)))
