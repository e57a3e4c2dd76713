#!/usr/bin/env python
"""SASS of one kernel of a library (cuobjdump -sass), opcode histogram: tools/sass_fn.py lib.so <substring of the mangled name> [--dump]"""
import subprocess, sys, re, collections
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
cur, fns = None, {}
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m: cur = m.group(1); fns[cur] = []; continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l): fns[cur].append(l)
for name, lines in fns.items():
    if sys.argv[2] in name:
        c = collections.Counter()
        for l in lines:
            s = re.sub(r"/\*[0-9a-f]+\*/", "", l).strip()
            s = re.sub(r"^@!?U?P\d+\s+", "", s)
            c[s.split()[0].split(".")[0]] += 1
        print(name, len(lines), "instructions;", " ".join("%s:%d" % kv for kv in c.most_common(14)))
        if "--dump" in sys.argv:
            print("\n".join(lines))
