#!/usr/bin/env python
"""Condense `make -C ltr-lowrank-sdp_b200/csrc ptxas-info` (ptxas -v of the whole library) into one line per kernel:
registers, stack frame, spill stores / loads, static shared memory.  CPU only (nvcc cross-compiles).

usage: make -C ltr-lowrank-sdp_b200/csrc ptxas-info 2>&1 | python scripts/ptxas_table.py [name-filter-regex] > profiles/rN_ptxas_<what>.txt
"""
import re
import subprocess
import sys


def template_head(mangled):
    """`_Z14k_mc_step_bulkILi32ELb0ELb0EEv...` -> `k_mc_step_bulk<32, false, false>` (c++filt, arguments dropped)"""
    try:
        full = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    except OSError:
        return mangled
    depth, out = 0, []
    for ch in full:
        if ch == "(" and depth == 0:
            break
        depth += ch == "<"
        depth -= ch == ">"
        out.append(ch)
    return "".join(out).replace("void ", "")


def main():
    flt = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
    name, props, rows = None, None, []
    for line in sys.stdin:
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            name, props = m.group(1), None
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            props = tuple(int(x) for x in m.groups())
            continue
        m = re.search(r"Used (\d+) registers", line)
        if m and name:
            sm = re.search(r"(\d+) bytes smem", line)
            rows.append((template_head(name), int(m.group(1)), props or (0, 0, 0), int(sm.group(1)) if sm else 0))
            name = None
    print("# ptxas -v (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo); `make -C ltr-lowrank-sdp_b200/csrc ptxas-info | scripts/ptxas_table.py`")
    print("# kernel | registers | stack / spill stores / spill loads (bytes) | static shared memory (bytes)")
    for nm, regs, (st, ss, sl), smem in rows:
        if flt and not flt.search(nm):
            continue
        print(f"{nm:72s} | {regs:3d} regs | stack {st} spill {ss}/{sl} | smem {smem}")


if __name__ == "__main__":
    main()
