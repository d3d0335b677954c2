#!/usr/bin/env python
"""List kernels of libnfm_sm100a.so that use local memory (STACK > 0): a sign
that a register array fell out of registers."""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "nitorch_fastmath_b200/libnfm_sm100a.so"
txt = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout.splitlines()
out = []
for i, l in enumerate(txt):
    m = re.match(r"\s*Function (\S+):", l)
    if m and i + 1 < len(txt):
        r = txt[i + 1]
        st, reg = re.search(r"STACK:(\d+)", r), re.search(r"REG:(\d+)", r)
        if st and int(st.group(1)) > int(sys.argv[2]) if len(sys.argv) > 2 else int(st.group(1)) > 0:
            out.append((int(st.group(1)), int(reg.group(1)), m.group(1)))
names = subprocess.run(["c++filt"], input="\n".join(o[2] for o in out), capture_output=True, text=True).stdout.splitlines()
for (st, reg, _), n in sorted(zip(out, names), key=lambda t: -t[0][0]):
    print(f"stack {st:5d}  regs {reg:3d}  {n.replace('nfm::', '')[:130]}")
print(len(out), "kernels with local memory")
