"""Extract one kernel's SASS from libpamg.so (cuobjdump) by a substring of its demangled name and print the
memory / control skeleton (LDG/STG/BAR/BRA/DADD/DMUL/DFMA lines).  Used to check the load batching of the SELL loop
 without a GPU:  python scripts/sass_extract.py 'k_spmv_sell<2, 0, false, 4, 3>' [--full]"""
import re
import subprocess
import sys

LIB = "parallel_amg_b200/libpamg.so"


def functions(lib=LIB):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, name, buf = {}, None, []
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                out[name] = buf
            name, buf = m.group(1), []
        elif name:
            buf.append(line)
    if name:
        out[name] = buf
    names = list(out)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return {re.sub(r"\(.*", "", d): out[n] for n, d in zip(names, dem)}


if __name__ == "__main__":
    want = sys.argv[1]
    full = "--full" in sys.argv
    for name, body in functions().items():
        if want in name:
            print("==", name, f"({len(body)} lines)")
            for line in body:
                m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
                if not m:
                    continue
                if full or re.search(r"LDG|STG|LDS|STS|BAR|BRA|DADD|DMUL|DFMA|MEMBAR|ATOM|RED|CCTL|EXIT", m.group(2)):
                    print(m.group(1), m.group(2))
