#!/usr/bin/env python
"""Per-object SASS opcode summary of libgbm_b200.so's CUDA units (cuobjdump -sass on csrc/*.o): the mnemonics that
prove which hardware paths the kernels use -- TMA (UTMALDG / UBLKCP), mbarrier (SYNCS), FP64 tensor pipe (DMMA),
tcgen05 (UTCIMMA = kind::i8 MMA, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit), REDUX, dp4a (IDP.4A), and the
system-scope 16-byte stores / loads of the peer-memory all-reduce (peer_sum_kernel: STG / LDG.E.128.STRONG.SYS).

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "genomicbreedingmodels.jl_b200", "csrc")
KEYS = ["UTMALDG", "UBLKCP", "SYNCS", "DMMA", "UTCIMMA", "UTCBAR", "LDTM", "REDUX", "IDP.4A", "I2F.F64", "DFMA", "RED.E.ADD.F64",
        "ATOMG", "LDS.128", "LDS.64", "HMMA", "IMMA", "STG.E.128.STRONG.SYS", "LDG.E.128.STRONG.SYS"]


def main():
    print("# SASS opcode summary (round 2)\n")
    print("`cuobjdump -sass` of every CUDA object of `libgbm_b200.so` (sm_100a), counts of the mnemonics that identify the")
    print("hardware path.  Regenerate with `python tools/sass_summary.py`.\n")
    print("| object | kernels | " + " | ".join(KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for obj in sorted(glob.glob(os.path.join(CSRC, "*.o"))):
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        if "Function" not in out:
            continue
        kernels = len(re.findall(r"^\s*Function :", out, flags=re.M))
        cnt = collections.Counter()
        for line in out.splitlines():
            m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    cnt[k] += 1
        print(f"| `{os.path.basename(obj)}` | {kernels} | " + " | ".join(str(cnt[k]) if cnt[k] else "" for k in KEYS) + " |")
    arch = subprocess.run(["cuobjdump", "-lelf", os.path.join(os.path.dirname(CSRC), "libgbm_b200.so")], capture_output=True, text=True).stdout
    archs = sorted(set(re.findall(r"sm_\d+a?", arch)))
    print(f"\nELF images in `libgbm_b200.so`: {', '.join(archs)} only (no PTX for other targets, no multi-arch fatbin).")


if __name__ == "__main__":
    main()
