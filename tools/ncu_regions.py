"""Offline: executed warp-instructions and stall samples per coarse region of a kernel (split at SYNCS/BAR/EXIT landmarks), merged into big chunks."""
import csv, re, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iSrc, iE = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
data = [r for r in rows[2:] if len(r) > iS and r[iS].isdigit()]
tot_s = sum(int(r[iS]) for r in data); tot_e = sum(int(r[iE] or 0) for r in data)
print("samples", tot_s, "executed warp-instr", tot_e)
land = re.compile(r'SYNCS\.PHASECHK|EXIT|UBLKCP')
cur = [0, 0, 0, 'START', 0]
for i, r in enumerate(data):
    if land.search(r[iSrc]) and cur[2] > 40:
        if cur[1] > 0.004 * tot_e or cur[0] > 0.01 * tot_s:
            print(f"{cur[4]:6d}-{i:6d} n={cur[2]:5d} samples={cur[0]:6d} ({100*cur[0]/tot_s:4.1f}%) exec={cur[1]:9d} ({100*cur[1]/tot_e:4.1f}%) after {cur[3]}")
        cur = [0, 0, 0, r[iSrc].strip()[:50], i]
    cur[0] += int(r[iS]); cur[1] += int(r[iE] or 0); cur[2] += 1
print(f"{cur[4]:6d}-{len(data):6d} n={cur[2]:5d} samples={cur[0]:6d} exec={cur[1]:9d} after {cur[3]}")
