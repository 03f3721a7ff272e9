#!/usr/bin/env python
"""Static pipe mix of a kernel's SASS (cuobjdump -sass -fun NAME lib.so | python profiles/sass_pipes.py [lo hi]).
Pipes as measured by profiles/ubench/pipes.cu on B200: ALU = LOP3/SHF/PRMT/SEL/IADD3/ISETP/..., FMA-heavy = IMAD/IDP,
XU = POPC/FLO/BREV/MUFU, LSU = LDS/STS/LDG/ATOM/RED/SHFL."""
import re, sys
ALU = {"LOP3", "SHF", "PRMT", "SEL", "IADD3", "ISETP", "PLOP3", "LEA", "VIADD", "VIMNMX", "VIMNMX3", "IMNMX", "MOV", "IABS", "BMSK", "SGXT", "P2R", "R2P", "ISETP", "FSEL", "VABSDIFF4", "IADD", "LOP", "CS2R"}
FMA = {"IMAD", "IDP", "FFMA", "FMUL", "FADD", "HFMA2"}
XU = {"POPC", "FLO", "BREV", "MUFU", "I2F", "F2I"}
LSU = {"LDS", "STS", "LDG", "STG", "ATOM", "ATOMS", "ATOMG", "RED", "REDG", "SHFL", "LDC", "LDCU", "LD", "ST", "MATCH", "REDUX", "VOTE"}
lo = int(sys.argv[1], 16) if len(sys.argv) > 1 else 0
hi = int(sys.argv[2], 16) if len(sys.argv) > 2 else 1 << 30
cnt = {}; pipes = {"ALU": 0, "FMA": 0, "XU": 0, "LSU": 0, "other": 0}
for l in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if not m: continue
    a = int(m.group(1), 16)
    if a < lo or a >= hi: continue
    t = m.group(2).split()
    op = t[1] if t[0].startswith("@") else t[0]
    base = op.split(".")[0]
    key = op if base in ("IMAD", "SHF") else base
    cnt[key] = cnt.get(key, 0) + 1
    p = "ALU" if base in ALU else "FMA" if base in FMA else "XU" if base in XU else "LSU" if base in LSU else "other"
    if op.startswith("IMAD.WIDE") or op.startswith("IMAD.HI"): pipes["FMA"] += 1      # quarter rate: counts double
    pipes[p] += 1
print(pipes, "total", sum(cnt.values()))
print(sorted(cnt.items(), key=lambda x: -x[1]))
