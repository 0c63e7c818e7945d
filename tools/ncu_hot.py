"""Hottest SASS instructions (stall samples) of one kernel of an .ncu-rep captured with --set full --import-source on.
usage: python tools/ncu_hot.py prof.ncu-rep <kernel name substring> [top N]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name",')
for blk in blocks[1:2]:
    lines = blk.splitlines()
    print("kernel:", lines[0][:100])
    rows = list(csv.reader(lines[1:]))
    hdr = rows[0]
    ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[1:] if len(r) == len(hdr)]
    tot = sum(int(r[isamp] or 0) for r in data)
    agg = {}
    for r in data:
        for i in stalls:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
    print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for n, r in sorted(enumerate(data), key=lambda nr: -int(nr[1][isamp] or 0))[:top]:
        st = sorted(((int(r[i] or 0), hdr[i]) for i in stalls), reverse=True)[:2]
        print(f"{n:5d} {int(r[isamp]):6d} {100*int(r[isamp])/max(tot,1):5.1f}% ex={r[iex]:>8} {r[isrc][:70]:70s} {st}")
