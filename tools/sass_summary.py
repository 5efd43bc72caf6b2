"""Per-kernel SASS opcode summary of libuse_b200.so: the in-tree evidence that the hot path is tcgen05 / TMEM / TMA code.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma kind::f16/tf32 (.2CTA = cta_group::2), UTMALDG = cp.async.bulk.tensor
(TMA) loads (.MULTICAST = cluster multicast), LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit -> mbarrier,
SYNCS = mbarrier try_wait / arrive, UTMAPF = prefetch.tensormap, HMMA / IMMA = legacy mma.sync (must be absent).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "universal-speech-enhancement_b200", "libuse_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.MULTICAST", "LDTM", "UTCBAR", "SYNCS", "UTMAPF", "UTCATOMSWS", "HMMA", "IMMA",
       "FFMA", "LDS", "STS", "LDG", "STG", "ATOMG", "RED", "SHFL", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    funcs, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        funcs[cur]["_total"] += 1
        base = op.split(".")[0]
        funcs[cur][base] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            funcs[cur]["UTCHMMA.2CTA"] += 1
        if base == "UTMALDG" and ".MULTICAST" in op:
            funcs[cur]["UTMALDG.MULTICAST"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(funcs)} kernels, arch {arch}")
    print("# kernel | instructions | " + " ".join(OPS))
    tot = collections.Counter()
    for (name, c), dn in zip(funcs.items(), demangle):
        short = re.sub(r"\(.*", "", dn)
        short = re.sub(r"^void use::", "", short)
        print(f"{short} | {c['_total']} | " + " ".join(f"{c[o]}" for o in OPS))
        tot.update(c)
    print("# TOTAL | %d | " % tot["_total"] + " ".join(f"{o}={tot[o]}" for o in OPS))
    if tot["HMMA"] or tot["IMMA"]:
        print("# WARNING: legacy mma.sync instructions present", file=sys.stderr)


if __name__ == "__main__":
    main()
