"""Code-size report of one kernel: SASS instructions per source region (the fused window kernels are bound by
the 32 KB instruction cache as much as by anything else).
    cuobjdump -xelf all libsnn_b200.so; nvdisasm --print-line-info -c X.cubin > x.sass
    python scripts/sass_regions.py x.sass <kernel-name-substring> <source.cu> 'label=pattern' ...
"""
import re
import sys

path, kern, srcpath = sys.argv[1:4]
marks = []
src = open(srcpath).read().split("\n")
for spec in sys.argv[4:]:
    label, pat = spec.split("=", 1)
    for i, l in enumerate(src):
        if pat in l:
            marks.append((i + 1, label))
            break
marks.sort()
cnt, cur_line, infn = {}, None, False
for l in open(path):
    if l.startswith("\t.section") and ".text." in l:
        infn = kern in l
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if infn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        cnt[cur_line] = cnt.get(cur_line, 0) + 1
agg = {}
base = srcpath.split("/")[-1]
for (f, ln), c in cnt.items():
    if f != base:
        name = "other: " + f
    else:
        name = "(top)"
        for m_ln, label in marks:
            if m_ln <= ln:
                name = label
    agg[name] = agg.get(name, 0) + c
tot = sum(agg.values())
print(f"{kern}: {tot} instructions, {tot * 16 / 1024:.1f} KB")
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    print(f"  {k:34s} {v:5d}  {v * 16 / 1024:5.1f} KB")
