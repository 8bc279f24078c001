#!/usr/bin/env python
"""Compare the device code of two liblsk builds kernel by kernel (cuobjdump -sass, instruction
text only).  Used to prove that a host-side or opt-in change left the kernels that were measured
and parity-tested on the GPU bit-for-bit unchanged:  python tools/sass_diff.py old.so new.so"""
import hashlib
import re
import subprocess
import sys


def kernels(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        mm = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if cur and mm:
            out[cur].append(mm.group(1))
    return {k: hashlib.md5("\n".join(v).encode()).hexdigest() for k, v in out.items()}


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    gone = sorted(set(a) - set(b))
    new = sorted(set(b) - set(a))
    diff = sorted(k for k in a if k in b and a[k] != b[k])
    print(f"{len(a)} kernels before, {len(b)} after; removed {len(gone)}, added {len(new)}, changed {len(diff)}")
    for tag, names in (("removed", gone), ("added", new), ("CHANGED", diff)):
        for n in names:
            print(f"  {tag}: {n[:110]}")
    sys.exit(1 if diff else 0)


if __name__ == "__main__":
    main()
