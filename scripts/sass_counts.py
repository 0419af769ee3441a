#!/usr/bin/env python
"""Static SASS instruction counts per kernel of diffdart_b200/lib/libddb.so (cuobjdump), as a markdown table:
  python scripts/sass_counts.py 'k_forwardILi5ELi2' 'k_backwardILi5ELi2' ...   (substrings of the mangled names)"""
import collections, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("DDB_LIB", os.path.join(ROOT, "diffdart_b200", "lib", "libddb.so"))
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        funcs[cur][m.group(1)] += 1
print("| kernel | all SASS | DFMA | DMUL | DADD | FP64 pipe slots | flop (DFMA = 2) | LDG | STG | LDGSTS / LDS | LDL / STL | LDC + LDCU | FSEL |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for pat in sys.argv[1:]:
    for name, c in funcs.items():
        if pat in name:
            short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.split("(")[0].replace("void ddb::", "")
            f, m, a = c["DFMA"], c["DMUL"], c["DADD"]
            print("| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d / %d | %d / %d | %d | %d |" % (short, sum(c.values()), f, m, a, f + m + a, 2 * f + m + a,
                  c["LDG"], c["STG"], c["LDGSTS"], c["LDS"], c["LDL"], c["STL"], c["LDC"] + c["LDCU"], c["FSEL"]))
