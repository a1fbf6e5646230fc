#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show tcgen05 / TMEM / TMA / cluster use in libdfmir_b200.so
(cuobjdump -sass; CPU-only).    python tools/sass_evidence.py > profiles/r1_sass_evidence.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "dfmir_b200", "libdfmir_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
pats = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "SYNCS", "UCGABAR", "RED.E", "ATOM"]
cur, counts, two_cta = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); two_cta[cur] = 0
        continue
    if cur is None:
        continue
    for p in pats:
        if re.search(r"\b" + re.escape(p), line):
            counts[cur][p] += 1
    if "UTCHMMA.2CTA" in line or "UTCBAR.2CTA" in line:
        two_cta[cur] += 1
print("# cuobjdump -sass dfmir_b200/libdfmir_b200.so (sm_100a): instruction counts per kernel; only kernels that use")
print("# tcgen05 (UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld) or TMA (UTMALDG = cp.async.bulk.tensor) are listed")
for fn, c in counts.items():
    if not (c["UTCHMMA"] or c["UTMALDG"]):
        continue
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", demangle(fn)).replace("(int)", "")
    name = re.sub(r"\(.*$", "", name).replace("void ", "")
    print(f"{name[:70]:70s} " + " ".join(f"{p}={c[p]}" for p in pats if c[p]) + (f" 2CTA-forms={two_cta[fn]}" if two_cta[fn] else ""))
