#!/usr/bin/env python
"""SASS listing of one kernel of the built library for profiles/: `cuobjdump -sass` of the object file, the function whose
mangled name matches, encodings stripped, a static opcode histogram on top.

    python scripts/sass_listing.py kc_resize.o 'kc_resize_v_tma_kernelILb0' profiles/sass_resize_v_tma_fast_r02.txt "title"
"""
import collections
import os
import re
import subprocess
import sys

obj, pat, out, title = sys.argv[1:5]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(root, "kanter_core_b200", "build", obj)
txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
keep, on = [], False
for line in txt.splitlines():
    if "Function :" in line:
        on = re.search(pat, line) is not None
    if on and not re.match(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", line):
        keep.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line))
if not keep:
    sys.exit("no function matches %r in %s" % (pat, path))
ops = collections.Counter()
for line in keep:
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m:
        ops[m.group(1)] += 1
with open(out, "w") as f:
    f.write(title + "\n")
    f.write("cuobjdump -sass of kanter_core_b200/build/%s (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -lineinfo); encodings stripped\n\n" % obj)
    f.write("opcode histogram (static): " + "  ".join("%s:%d" % kv for kv in ops.most_common()) + "\n\n")
    f.write("\n".join(keep) + "\n")
print(out, len(keep), "lines;", ", ".join("%s:%d" % (k, ops[k]) for k in ("UTMALDG.2D", "UTMASTG.2D", "UBLKCP.S.G", "FFMA2", "LDS.128") if ops[k]))
