"""cuobjdump -sass of libaccmsm.so, split per kernel: instruction histogram (and optionally an excerpt) of the
kernels whose mangled name contains a pattern.  Usage: python tools/sass_stats.py k_accumulateILi0 [--excerpt N]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("ACCMSM_SO", os.path.join(ROOT, "accumulation_b200", "libaccmsm.so"))
pat = sys.argv[1] if len(sys.argv) > 1 else "k_accumulateILi0"
excerpt = int(sys.argv[sys.argv.index("--excerpt") + 1]) if "--excerpt" in sys.argv else 0
text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, funcs = None, collections.OrderedDict()
for line in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur:
        funcs[cur].append(line)
for name, lines in funcs.items():
    if pat not in name:
        continue
    ops = collections.Counter()
    body = []
    for l in lines:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops[m.group(3)] += 1
            body.append(l.rstrip())
    total = sum(ops.values())
    fam = collections.Counter()
    for k, v in ops.items():
        fam[k.split(".")[0] + (".WIDE" if ".WIDE" in k else "")] += v
    print(f"{name}: {total} instructions")
    print("  " + ", ".join(f"{k} {v}" for k, v in fam.most_common(14)))
    if excerpt:
        # the first run of IMAD.WIDE carry chains: the Montgomery product
        start = next((i for i, l in enumerate(body) if "IMAD.WIDE.U32.X" in l), 0)
        print("\n".join(body[max(0, start - 4):start + excerpt]))
