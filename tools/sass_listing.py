"""Opcode histogram of the built library's sm_100a SASS, per kernel (evidence that the tensor-core / TMA / TMEM paths are
hand-written):  tcgen05.mma -> UTCHMMA(.2CTA) | cp.async.bulk.tensor -> UTMALDG / UTMASTG | tcgen05.ld / st -> LDTM / STTM |
tcgen05.commit -> UTCBAR | mbarrier -> SYNCS.*        usage: python tools/sass_listing.py > profiles/sass_libvtoonify_b200.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "vtoonify_b200/lib/libvtoonify_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEEP = re.compile(r"^(UTC|UTMA|LDTM|STTM|UBLKCP|SYNCS|UCGABAR|HMMA|FENCE\.VIEW|ELECT|MEMBAR)")
per = collections.OrderedDict()
fn = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        per[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn and KEEP.match(m.group(1)):
        per[fn][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {so}: tensor-core / TMA / TMEM / mbarrier opcodes per kernel")
tot = collections.Counter()
for (fn, c), name in zip(per.items(), names):
    if not c:
        continue
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    print(f"\n{name.split('(')[0]}")
    for op, n in sorted(c.items(), key=lambda kv: (-kv[1], kv[0])):
        print(f"    {n:6d}  {op}")
    tot.update(c)
print("\nTOTAL")
for op, n in sorted(tot.items(), key=lambda kv: (-kv[1], kv[0])):
    print(f"    {n:6d}  {op}")
