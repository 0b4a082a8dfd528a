"""Offline: bucket ncu warp-stall samples of a kernel between landmark SASS instructions."""
import csv, re, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iSrc, iE = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
data = [r for r in rows[2:] if len(r) > iS and r[iS].isdigit()]
tot = sum(int(r[iS]) for r in data)
land = re.compile(r'LDTM|STTM|UTCHMMA|UTCBAR|BAR\.|SYNCS|UBLKCP|EXIT|FENCE|UTCATOM')
segs, seg, last, lastidx, ninstr = [], 0, 'START', 0, 0
for i, r in enumerate(data):
    if land.search(r[iSrc]):
        segs.append((lastidx, i, last, seg)); seg = 0; last = r[iSrc].strip()[:60]; lastidx = i
    seg += int(r[iS])
segs.append((lastidx, len(data), last, seg))
print("total samples", tot, "instructions", len(data))
for a, b, l, s in segs:
    if s >= thr:
        print(f"{a:6d}-{b:6d} {s:6d} {100*s/tot:5.1f}%  after: {l}")
