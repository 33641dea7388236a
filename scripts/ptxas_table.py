"""Register / spill / shared-memory table of every kernel in libpamg.so from the ptxas -v log the build keeps
(parallel_amg_b200/build/ptxas.log):  python scripts/ptxas_table.py [substring]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def table(path=os.path.join(ROOT, "parallel_amg_b200", "build", "ptxas.log")):
    log = open(path).read().splitlines()
    rows = []
    for i, line in enumerate(log):
        m = re.search(r"Compiling entry function '([^']*)'", line)
        if not m:
            continue
        blk = " ".join(log[i:i + 5])
        regs = re.search(r"Used (\d+) registers", blk)
        spill = re.search(r"(\d+) bytes spill stores", blk)
        smem = re.search(r"(\d+) bytes smem", blk)
        stack = re.search(r"(\d+) bytes stack frame", blk)
        rows.append([m.group(1), int(regs.group(1)) if regs else -1, int(spill.group(1)) if spill else 0,
                     int(stack.group(1)) if stack else 0, int(smem.group(1)) if smem else 0])
    dem = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
    for r, d in zip(rows, dem):
        r[0] = re.sub(r"\(.*", "", d).replace("void ", "").replace("pamg::", "")
    return rows


if __name__ == "__main__":
    want = sys.argv[1] if len(sys.argv) > 1 else ""
    print(f"{'kernel':58s} {'regs':>4s} {'CTAs/SM':>7s} {'spill B':>7s} {'stack B':>7s} {'smem B':>7s}")
    for name, regs, spill, stack, smem in sorted(table()):
        if want in name:
            occ = min(2048 // 256, 65536 // (max(regs, 1) * 256), (227 * 1024) // max(smem, 1) if smem else 99)
            print(f"{name:58s} {regs:4d} {occ:7d} {spill:7d} {stack:7d} {smem:7d}")
