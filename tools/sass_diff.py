#!/usr/bin/env python
"""Kernel-by-kernel SASS comparison of two builds of libwowsr.so (no GPU needed: cuobjdump -sass).

Used when the library is changed without access to hardware: every kernel that already passed the GPU parity tests must come
out byte-identical (instruction encodings included); only new kernels may appear.

    python tools/sass_diff.py <reference commit> [lib.so]     # builds the commit in a temporary worktree and compares
"""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "sentinel2-super-resolution-poc_b200"


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:  # anonymous-namespace kernels carry a hash of their translation unit in the mangled name
            cur = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_(\w+?)_cu_[0-9a-f]+\d*", r"ANON_\1_", m.group(1))
            res[cur] = []
        elif cur:
            res[cur].append(line.strip())
    return {k: hashlib.sha1("\n".join(v).encode()).hexdigest() for k, v in res.items()}


def main():
    commit = sys.argv[1]
    new = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, PKG, "libwowsr.so")
    with tempfile.TemporaryDirectory() as tmp:
        wt = os.path.join(tmp, "wt")
        subprocess.run(["git", "-C", ROOT, "worktree", "add", "-f", "-q", wt, commit], check=True)
        try:
            subprocess.run(["make", "-C", os.path.join(wt, PKG, "csrc"), "-j4"], check=True, stdout=subprocess.DEVNULL)
            a = kernels(os.path.join(wt, PKG, "libwowsr.so"))
        finally:
            subprocess.run(["git", "-C", ROOT, "worktree", "remove", "--force", wt], check=False)
    b = kernels(new)
    print(f"# SASS of {new} against a fresh build of commit {commit}")
    bad = 0
    for k in sorted(set(a) | set(b)):
        st = "SAME" if a.get(k) == b.get(k) else ("NEW" if k not in a else ("GONE" if k not in b else "DIFF"))
        bad += st in ("DIFF", "GONE")
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        print(f"{st:5s} {re.sub(r'\(.*', '', name) if not name.startswith('_ZN') else k}")
    print(f"# {bad} kernel(s) changed or removed")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
