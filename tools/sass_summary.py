"""Per-kernel SASS mnemonic counts of libkzd.so (cuobjdump -sass): which kernels carry bulk-copy (UBLKCP), mbarrier
(SYNCS), cluster barrier (UCGABAR), 128-bit global accesses, REDUX, MUFU and atomics.
usage: python tools/sass_summary.py [libkzd.so] > profiles/rN_sass_summary.md"""
import collections, os, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kuzushiji_vision_b200", "libkzd.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = collections.OrderedDict([("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("UCGABAR", r"\bUCGABAR"), ("LDG.128", r"\bLDG\.E\.\S*128"),
                                ("STG.128", r"\bSTG\.E\.\S*128"), ("LDS.128", r"\bLDS\.128"), ("REDUX", r"\bC?REDUX"), ("MUFU", r"\bMUFU"),
                                ("ATOM/RED", r"\b(ATOMG|ATOMS|ATOM|RED)\b"), ("BAR", r"\bBAR\.")])
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0].replace("kzd::", "")
cur, counts, total = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1))
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        total[cur] += 1
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
print("SASS of `kuzushiji_vision_b200/libkzd.so` (sm_100a, `cuobjdump -sass`), instruction counts per kernel\n")
print("| kernel | instructions | " + " | ".join(pats) + " |")
print("|---|---|" + "---|" * len(pats))
for k, c in counts.items():
    print(f"| `{k[-70:]}` | {total[k]} | " + " | ".join(str(c[p]) for p in pats) + " |")
