"""Which device functions differ between two builds of libsnn_b200.so?  (whitespace, address comments and the path hash
of the anonymous namespaces normalised).  Used to show what changed in the device code since a GPU-tested build:
    python scripts/sass_diff.py old.so bindsnet_b200/csrc/libsnn_b200.so"""
import re
import subprocess
import sys

SKIP = ("identifier =", "Fatbin", "=====", "arch =", "code version", "host =", "compile_size", "producer")


def funcs(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    out = re.sub(r"_GLOBAL__N__[0-9a-f]+_", "_GLOBAL__N__X_", out)
    d, cur, buf = {}, None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                d[cur] = buf
            cur, buf = m.group(1), []
        elif cur and not line.strip().startswith("//") and not any(s in line for s in SKIP):
            buf.append(" ".join(re.sub(r"/\*[0-9a-f]{4}\*/", "", line).split()))
    if cur:
        d[cur] = buf
    return d


if __name__ == "__main__":
    a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
    print(f"{len(a)} / {len(b)} functions")
    for k in sorted(set(a) | set(b)):
        if a.get(k) != b.get(k):
            print("differs:" if k in a and k in b else ("only in new:" if k in b else "only in old:"), k)
