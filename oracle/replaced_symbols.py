"""TEST INFRASTRUCTURE ONLY.  Prints `name ref_name` lines for objcopy --redefine-syms: the reference entry points that
include/hsrle_b200.h (Part 1) re-declares, i.e. the symbols the product library provides in the reference's place."""
import re
import sys

txt = open(sys.argv[1]).read()
part1 = txt.split("Part 1: reference entry points")[1].split("Part 2: GPU-resident interface")[0]
part1 = re.sub(r"/\*.*?\*/", "", part1, flags=re.S)
names = sorted(set(re.findall(r"\b(rle\w+)\s*\(", part1)))
for n in names:
    print(n, "ref_" + n)
