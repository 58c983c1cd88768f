"""Prints the SASS of one kernel of a shared library: tools/sass_fn.py lib.so substring [--hist]"""
import subprocess, sys, re, collections
lib, key = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
parts = re.split(r"(?m)^\s*Function : ", out)
for p in parts[1:]:
    name = p.split("\n", 1)[0]
    if key in name:
        lines = [l for l in p.split("\n") if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l) and not re.match(r"\s*/\* 0x", l)]
        if "--hist" in sys.argv:
            c = collections.Counter(re.sub(r"@!?U?P\d\s+", "", l.split("*/", 1)[1].strip()).split()[0].split(".")[0] for l in lines)
            print(name, len(lines), c.most_common(12))
        else:
            print(name); print("\n".join(lines))
