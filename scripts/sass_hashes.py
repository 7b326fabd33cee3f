"""Per-kernel hash of the device code of a build: md5 of every function's SASS instruction stream (`cuobjdump -sass`, addresses and
encodings stripped).  Two builds whose kernels hash alike run the same device code whatever the host code or the comments look like -
this is how a source change made without a GPU at hand was tied to a build that HAD run the GPU suite (round 2: the final
defaults against the A/B build `lib_MP.so` of visit X).
usage: python scripts/sass_hashes.py [lib.so] > profiles/rNN_sass_hashes.txt ; python scripts/sass_hashes.py a.so b.so  (compare)"""
import hashlib
import re
import subprocess
import sys


def hashes(lib):
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = res.setdefault(m.group(1), [])
            continue
        if cur is not None:
            m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(.*?);', line)
            if m:
                cur.append(m.group(1).strip())
    names = subprocess.run(['c++filt'] + list(res), capture_output=True, text=True).stdout.splitlines()
    return {n.split('(')[0]: (len(v), hashlib.md5('\n'.join(v).encode()).hexdigest()) for n, v in zip(names, res.values())}


if __name__ == '__main__':
    libs = sys.argv[1:] or ['openglue_b200/libopenglue_b200.so']
    if len(libs) == 1:
        for k, (n, h) in sorted(hashes(libs[0]).items()):
            print(f'{h}  {n:6d}  {k}')
    else:
        a, b = hashes(libs[0]), hashes(libs[1])
        same = [k for k in a if k in b and a[k] == b[k]]
        print(f'{len(same)} kernels identical; different: {sorted(k for k in a if k in b and a[k] != b[k])}; '
              f'only in {libs[0]}: {sorted(k for k in a if k not in b)}; only in {libs[1]}: {sorted(k for k in b if k not in a)}')
