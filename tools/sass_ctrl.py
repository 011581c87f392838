#!/usr/bin/env python
"""tools/sass_ctrl.py <binary> <kernel-substring> [first-opcode-regex] [n]: SASS with the scheduling control fields
(stall count, yield, write/read scoreboard, wait mask) decoded from the upper instruction word."""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
on = False
rows = []
cur = None
for ln in out.splitlines():
    if "Function :" in ln:
        on = sys.argv[2] in ln
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", ln)
    if m:
        cur = [m.group(1), m.group(2).strip(), int(m.group(3), 16), None]
        rows.append(cur)
        continue
    m = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", ln)
    if m and cur is not None and cur[3] is None:
        cur[3] = int(m.group(1), 16)
start = 0
if len(sys.argv) > 3:
    for i, r in enumerate(rows):
        if re.search(sys.argv[3], r[1]):
            start = i
            break
n = int(sys.argv[4]) if len(sys.argv) > 4 else 80
for r in rows[start:start + n]:
    hi = r[3] or 0
    stall, yld, wr, rd, wait = (hi >> 41) & 0xf, (hi >> 45) & 1, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3f
    print(f"{r[0]}  st={stall:2d} y={yld} wr={wr if wr != 7 else '-'} rd={rd if rd != 7 else '-'} wait={wait:06b}  {r[1][:80]}")
