"""Per-kernel SASS instruction counts of the shipped library (cuobjdump -sass, no GPU needed): how data moves and which
arithmetic forms the kernels use.  Usage: python tools/sass_evidence.py > profiles/r02_sass_evidence.csv"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gan-based-video-style-transfer_b200", "csrc", "libtcl_b200.so")
COLS = [("UTMALDG", r"UTMALDG"), ("UTMAPF", r"UTMAPF"), ("UBLKCP", r"UBLKCP|UBLKPF"), ("SYNCS", r"SYNCS"), ("REDUX", r"C?REDUX"), ("LDS", r"LDS"),
        ("STS", r"STS"), ("LDG", r"LDG"), ("STG", r"STG"), ("RED+ATOMG", r"RED|ATOMG"), ("FFMA", r"FFMA"), ("FFMA2", r"FFMA2"), ("FADD2", r"FADD2"),
        ("FMUL2", r"FMUL2"), ("MUFU", r"MUFU"), ("NANOSLEEP", r"NANOSLEEP"), ("HMMA+UTCMMA", r"HMMA|UTC\w*MMA")]

if __name__ == "__main__":
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if cur and m:
            op = m.group(1)
            counts[cur]["instructions"] += 1
            for name, pat in COLS:
                if re.fullmatch(pat, op):
                    counts[cur][name] += 1
    demangled = dict(zip(order, subprocess.run(["c++filt"] + order, capture_output=True, text=True).stdout.splitlines()))
    print("# SASS evidence (cuobjdump -sass libtcl_b200.so, sm_100a): per kernel, the instructions that show how data moves and how it is computed")
    print("# UTMALDG = TMA tensor load (cp.async.bulk.tensor), UTMAPF = TMA L2 prefetch, UBLKCP = 1-D bulk copy / prefetch (cp.async.bulk), SYNCS = mbarrier ops,")
    print("# REDUX = redux.sync, LDS/STS = shared memory, LDG/STG = global loads / stores, RED+ATOMG = global atomics, FFMA2/FADD2/FMUL2 = packed fp32 arithmetic,")
    print("# NANOSLEEP = back-off of waiting helper warps; HMMA+UTCMMA = tensor-core instructions (none: nothing on this path is a contraction)")
    print("kernel,instructions," + ",".join(n for n, _ in COLS))
    for k in order:
        name = re.sub(r"\s+", " ", demangled.get(k, k)).replace(",", ";")
        print(name[:200] + "," + str(counts[k]["instructions"]) + "," + ",".join(str(counts[k][n]) for n, _ in COLS))
    tot = collections.Counter()
    for k in order:
        tot.update(counts[k])
    print("TOTAL," + str(tot["instructions"]) + "," + ",".join(str(tot[n]) for n, _ in COLS))
