"""Executed warp-instructions and stall samples of one kernel aggregated by CUDA source line: joins the SASS rows of an
.ncu-rep (--set full --import-source on) with nvdisasm's line table of the same libkzd.so (instruction order is
identical). usage: python tools/ncu_lines.py prof.ncu-rep <mangled-name substring> <kernel regex for ncu> [top N]"""
import collections, csv, glob, os, re, subprocess, sys, tempfile
rep, mangled, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "kuzushiji_vision_b200", "libkzd.so")], cwd=tmp, capture_output=True)
lines = []
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    cur, line, on = None, 0, False
    for l in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            on = mangled in m.group(1) and not lines
            continue
        if l.startswith("//-----") or ".section" in l:
            if on and lines:
                on = False
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(line)
    if lines:
        break
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
blk = raw.split('"Kernel Name",')[1].splitlines()
rows = list(csv.reader(blk[1:]))
hdr = rows[0]
data = [r for r in rows[1:] if len(r) == len(hdr)]
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
print(f"{len(lines)} SASS instructions in nvdisasm, {len(data)} rows in the report")
agg = collections.defaultdict(lambda: [0, 0, 0])
for ln, r in zip(lines, data):
    a = agg[ln]
    a[0] += int(r[iex] or 0); a[1] += int(r[isamp] or 0); a[2] += 1
tex, tsm = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
src = {}
print(f"total executed warp-instructions {tex}, samples {tsm}")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, n = ln if ln else ("?", 0)
    if f not in src:
        try:
            src[f] = open(os.path.join(root, "kuzushiji_vision_b200", "csrc", f)).read().splitlines()
        except Exception:
            src[f] = []
    text = src[f][n - 1].strip()[:90] if 0 < n <= len(src[f]) else ""
    print(f"{f}:{n:4d} ex {a[0]:9d} {100*a[0]/max(tex,1):5.1f}%  samples {100*a[1]/max(tsm,1):5.1f}%  sass {a[2]:4d} | {text}")
