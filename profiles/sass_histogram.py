#!/usr/bin/env python
"""Static SASS evidence from the built objects (no GPU needed): per kernel, the opcode histogram of `cuobjdump -sass` plus the
mnemonics that prove the Blackwell-specific paths -- UBLKCP / SYNCS (cp.async.bulk + mbarrier: the TMA staging of the lane
tables), LDGSTS (cp.async of the plan records), FADD2 (packed f32x2 butterflies), IMAD.HI (multiply-high floor lines), REDUX.
usage: python profiles/sass_histogram.py [kernel-name-substring ...]  > profiles/r02_sass_histograms.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "nvorbis_b200", "csrc", "build")
want = sys.argv[1:] or ["k_imdct_fused_tILb0", "k_spectrum_wfILi2ELi1ELb0", "k_unpack", "k_pcm_s16", "k_imdct_generic"]
MARK = ("UBLKCP", "SYNCS", "LDGSTS", "FADD2", "REDUX", "UTMA", "BAR", "WARPSYNC", "SHFL", "LDS", "STS", "LDG", "STG", "FFMA", "FMUL", "IMAD")
for obj in ("nvb_fused.o", "nvb_kernels.o", "nvb_unpack.o"):
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
    for m in re.finditer(r"Function : (\S+)\n(.*?)(?=\n\s*Function : |\Z)", txt, re.S):
        name, body = m.group(1), m.group(2)
        if not any(w in name for w in want):
            continue
        ops = collections.Counter()
        for line in body.splitlines():
            mm = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if mm:
                ops[mm.group(1).split(".")[0]] += 1
        total = sum(ops.values())
        print(f"== {name} ({obj}): {total} SASS instructions")
        print("   top: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
        print("   markers: " + ", ".join(f"{k} {sum(v for o, v in ops.items() if o.startswith(k))}" for k in MARK if any(o.startswith(k) for o in ops)))
