"""Counts the SASS mnemonics that prove which hardware paths the kernels use (DMMA = FP64 tensor
core, UBLKCP/SYNCS = TMA bulk copy + mbarrier, REDUX = hardware warp reduce, MATCH/VOTE = warp
aggregation).  Runs where the library was built (cuobjdump, no GPU).
usage: python scripts/sass_summary.py > profiles/r01_sass_mnemonics.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parents[1] / "lineqgpr_b200" / "liblineqgpr_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
names = {}
cnt = collections.defaultdict(collections.Counter)
fn = None
WANT = ("DMMA", "UBLKCP", "SYNCS", "REDUX", "CREDUX", "MATCH", "VOTE", "DFMA", "DSETP", "MUFU", "BAR")
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        op = m.group(1)
        cnt[fn]["_all"] += 1
        if op in WANT:
            cnt[fn][op] += 1
dem = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS mnemonic counts per kernel of {lib.name} (cuobjdump -sass, sm_100a)")
for mangled, name in sorted(zip(cnt, dem), key=lambda t: t[1]):
    c = cnt[mangled]
    short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", ""))
    keys = " ".join(f"{k}={c[k]}" for k in WANT if c[k])
    print(f"{short:70s} instr={c['_all']:6d}  {keys}")
