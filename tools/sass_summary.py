"""Per-kernel count of the Blackwell-only SASS instructions in declip_b200/_C.so (no GPU needed):
`python tools/sass_summary.py > profiles/r02_sass_blackwell.md`.

UTCHMMA = tcgen05.mma (bf16 kind::f16), .2CTA = cta_group::2; LDTM / STTM = tcgen05.ld / st (TMEM); UTMALDG / UTMASTG =
cp.async.bulk.tensor load / store (TMA), UBLKCP = cp.async.bulk (non-tensor); UTCBAR = tcgen05.commit; UTCATOM* / UTCCP
would be TMEM alloc / copy helpers; SYNCS = mbarrier ops; REDG / RED = red.global."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "declip_b200", "_C.so")
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCIMMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UTMAREDG|UBLKCP|UTCBAR|UTCATOMSWS|UTCCP|HMMA|MUFU)((?:\.[A-Z0-9_]+)*)")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], check=True, capture_output=True, text=True).stdout
    filt = subprocess.run(["c++filt"], input=txt, capture_output=True, text=True).stdout or txt
    per = collections.OrderedDict()
    cur = None
    arch = set(re.findall(r"arch = (sm_\w+)", filt))
    for line in filt.splitlines():
        m = re.search(r"Function : (.*)$", line)
        if m:
            cur = m.group(1).strip()
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = PAT.search(line)
        if m:
            op, mods = m.group(1), m.group(2)
            keep = [x for x in mods.split(".") if x in ("2CTA", "MULTICAST", "2D", "3D", "4D", "5D", "IM2COL", "32x32b", "x32",
                                                        "x16", "x64", "TANH", "EX2", "RCP", "RSQ", "LG2")]
            per[cur][op + "".join("." + k for k in keep)] += 1
    print("# Blackwell instructions per kernel in `declip_b200/_C.so` (`cuobjdump -sass`, %s)\n" % ", ".join(sorted(arch)))
    print("`python tools/sass_summary.py` (CPU only).  UTCHMMA = `tcgen05.mma.kind::f16` (`.2CTA` = `cta_group::2`), LDTM = `tcgen05.ld`,")
    print("UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = `cp.async.bulk`, UTCBAR = `tcgen05.commit`; HMMA = legacy `mma.sync`")
    print("(only the round-1 fallback attention core `attn_fwd/bwd_kernel` may contain it).\n")
    print("| kernel | tcgen05 / TMEM / TMA instructions (count) |")
    print("|---|---|")
    tot = collections.Counter()
    for k, c in per.items():
        c2 = collections.Counter({a: b for a, b in c.items() if not a.startswith("MUFU")})
        if not c2:
            continue
        tot.update(c2)
        name = re.sub(r"\(.*$", "", k)
        print("| `%s` | %s |" % (name, ", ".join("%s x%d" % (a, b) for a, b in sorted(c2.items()))))
    print("\nTotals: " + ", ".join("%s x%d" % (a, b) for a, b in sorted(tot.items())))
    no_bw = [re.sub(r"\(.*$", "", k) for k, c in per.items() if not any(not a.startswith("MUFU") for a in c)]
    print("\nKernels without any of these (%d; row / element-wise HBM-bound kernels): %s" % (len(no_bw), ", ".join("`%s`" % n for n in no_bw)))


if __name__ == "__main__":
    sys.exit(main())
