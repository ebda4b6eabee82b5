"""Blackwell-specific SASS instruction census of the shipped library ("commit the grep"):

    python tools/sass_census.py > profiles/r2_sass_census.txt

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA load (cp.async.bulk.tensor), UTCBAR =
tcgen05.commit, SYNCS = mbarrier operations, FFMA2 = packed fp32 FMA, HMMA would be the legacy mma.sync path."""
import collections
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "nano_vs_slam_b200", "lib", "libnanovs.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|HMMA|HGMMA|FFMA2|SYNCS)\b([.\w]*)")
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    for op, mods in pat.findall(line):
        cnt[cur][op + (".MULTICAST" if "MULTICAST" in mods else "")] += 1
names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass nano_vs_slam_b200/lib/libnanovs.so | census of tcgen05 / TMEM / TMA / mbarrier / FFMA2 per kernel")
tot = collections.Counter()
for mangled, name in zip(cnt, names):
    c = cnt[mangled]
    if not any(k.startswith(("UTC", "LDTM", "STTM", "UTMA")) for k in c):
        continue
    print(re.sub(r"\(.*", "", name), dict(sorted(c.items())))
    tot.update(c)
print("TOTAL (tensor-core kernels)", dict(sorted(tot.items())))
ff = {re.sub(r"\(.*", "", n): cnt[m]["FFMA2"] for m, n in zip(cnt, names) if cnt[m].get("FFMA2") and not any(
    k.startswith(("UTC", "LDTM")) for k in cnt[m])}
print("FFMA2 (packed fp32 FMA) in CUDA-core kernels:", ff)
print("HMMA / HGMMA (legacy tensor paths):", sum(c.get("HMMA", 0) + c.get("HGMMA", 0) for c in cnt.values()))
