"""Blackwell-native evidence from the built library: per-kernel counts of the SASS mnemonics that matter
(UTMALDG = TMA tensor loads, SYNCS = mbarrier operations, FFMA2 / FADD2 / FMUL2 = packed f32x2 arithmetic, IDP = dp4a,
PRMT byte splices, UTCxMMA / TMEM = none: the path has no dense contraction) and registers / spills from ptxas -v.
usage: python scripts/sass_counts.py > profiles/r2_sass_counts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "libstacker.rs_b200", "libstacker_cuda.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
keys = ["UTMALDG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "IDP", "PRMT", "LDG", "LDS", "DFMA", "UTC", "TMEM"]
print("arch:", sorted(set(re.findall(r"arch = (sm_\w+)", sass))))
print(f"{'kernel':100s} " + " ".join(f"{k:>7s}" for k in keys) + "   instr")
cur, counts, total = None, collections.Counter(), 0
rows = []
idx = 0
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, counts, total))
        cur, counts, total = names[idx].split("(")[0][:100], collections.Counter(), 0
        idx += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if m:
        total += 1
        op = m.group(1)
        for k in keys:
            if op.startswith(k):
                counts[k] += 1
if cur:
    rows.append((cur, counts, total))
default = ("ecc_iter_v2_kernel<3, false, stk::EccCfg<256, 16, 8, 2, 2, 3, 1>", "warp_accumulate_v2_kernel<3, true, true>", "prep_stream_kernel<3, 2>",
           "tenengrad_stream_kernel<16>", "peer_reduce_scale_kernel<8>", "sharpness_stream_kernel", "resize_area", "lane_sum", "seed_acc")
for name, c, t in rows:
    if any(d in name for d in default):
        print(f"{name:100s} " + " ".join(f"{c[k]:7d}" for k in keys) + f" {t:7d}")
tot = collections.Counter()
for _, c, _ in rows:
    tot.update(c)
print(f"{'all ' + str(len(rows)) + ' kernels of the library':100s} " + " ".join(f"{tot[k]:7d}" for k in keys))
log = os.path.join(ROOT, "libstacker.rs_b200", "csrc", "ptxas.log")
if os.path.exists(log):
    txt = open(log).read()
    print("\nptxas -v (default kernels):")
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", txt):
        dn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        if any(d in dn for d in default):
            print(f"  {dn[:100]:100s} {m.group(5):>4s} registers, spill stores/loads {m.group(3)}/{m.group(4)} bytes")
