"""Derive ctypes argument codes from include/mggan_b200.h (used by tests and to refresh cuda_ext.SIGNATURES)."""
import os
import re

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "mggan_b200.h")


def parse(path=HEADER):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|const char\*)\s+(mggan_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        codes = ""
        for a in [x.strip() for x in args.split(",") if x.strip() and x.strip() != "void"]:
            if "MgganTensorTable" in a:
                codes += "t"
            elif "MgganPeerTable" in a:
                codes += "T"
            elif a.startswith("long long") and "*" not in a:
                codes += "q"
            elif "*" in a:
                codes += "p"
            elif a.startswith("cudaStream_t"):
                codes += "s"
            elif a.startswith("unsigned long long"):
                codes += "Q"
            elif a.startswith("double"):
                codes += "d"
            elif a.startswith("float"):
                codes += "f"
            elif a.startswith("int"):
                codes += "i"
            else:
                raise ValueError(f"{name}: cannot classify '{a}'")
        out[name] = (ret, codes)
    return out


if __name__ == "__main__":
    for n, (r, c) in parse().items():
        print(f'    "{n}": "{c}",')
