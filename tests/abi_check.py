"""Derives ctypes argument kinds from include/mirres_b200.h and compares them with _lib.SIGNATURES."""
import re


def header_signatures(header_path):
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|size_t)\s+(mirres_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        kinds = ""
        if args not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    kinds += "p"
                elif a.startswith("unsigned int"):
                    kinds += "u"
                elif a.startswith("size_t"):
                    kinds += "z"
                elif a.startswith("float"):
                    kinds += "f"
                elif a.startswith("int"):
                    kinds += "i"
                else:
                    raise ValueError("unknown arg %r in %s" % (a, name))
        out[name] = (ret, kinds)
    return out
