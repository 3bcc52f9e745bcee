#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of liblmrl_b200.so, the counts of the Blackwell-specific mnemonics (tcgen05 MMA = UTCHMMA,
TMEM load = LDTM, TMA load / store / reduce = UTMALDG / UTMASTG / UTMAREDG, mbarrier = SYNCS, legacy tensor core = HMMA) and a short
excerpt around the first tcgen05 MMA of the GEMM kernel.  Runs in the build container (cuobjdump, no GPU).

    python scripts/sass_evidence.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lmrl_gym_b200", "csrc", "liblmrl_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "SYNCS", "UTCBAR", "HMMA", "LDSM", "LDGSTS", "REDG", "ATOMG", "SHFL", "MUFU"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)
    names = demangle(list(funcs))
    print("# cuobjdump -sass lmrl_gym_b200/csrc/liblmrl_b200.so (sm_100a), summarised by scripts/sass_evidence.py")
    print("# per kernel: instruction count and the counts of the mnemonics that identify the hardware path it uses")
    print("#   UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld (TMEM -> registers), UTMALDG / UTMASTG / UTMAREDG = TMA tile load / store /")
    print("#   reduce-add (cp.async.bulk.tensor / cp.reduce.async.bulk.tensor), SYNCS = mbarrier ops, UTCBAR = tcgen05.commit,")
    print("#   HMMA = mma.sync (legacy tensor-core path), LDSM = ldmatrix, LDGSTS = cp.async, SHFL = warp shuffle, MUFU = SFU (exp / log / rsqrt)\n")
    first_excerpt = None
    for f, lines in funcs.items():
        ins = [l for l in lines if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        if not ins:
            continue
        cnt = collections.Counter()
        for l in ins:
            for mn in MNEMONICS:
                if re.search(r"\b" + mn + r"\b|\b" + mn + r"\.", l):
                    cnt[mn] += 1
        name = names.get(f, f)
        short = re.sub(r"\(.*", "", name.replace("lmrl::", ""))[:110]
        tags = "  ".join(f"{k}={v}" for k, v in cnt.items() if v)
        print(f"{short:<112} {len(ins):>6} instr   {tags}")
        if first_excerpt is None and "gemm_bf16_tcgen05_kernel<0, 256>" in name:
            idx = next((i for i, l in enumerate(ins) if "UTCHMMA" in l), None)
            if idx is not None:
                first_excerpt = (short, ins[max(0, idx - 6):idx + 10])
    if first_excerpt:
        print(f"\n# excerpt: the MMA issue loop of {first_excerpt[0]} (mbarrier wait -> 4 x UTCHMMA per 64-wide k-block -> commit)")
        for l in first_excerpt[1]:
            print("   ", l.strip()[:150])


if __name__ == "__main__":
    main()
