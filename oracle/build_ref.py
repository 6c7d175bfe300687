#!/usr/bin/env python
"""Build oracle/_ref/libslvref.so: the reference's own Fortran kernels, transpiled to C by
oracle/f77c.py from the sources WHERE THEY LIE under /root/reference, plus oracle/f77_support.c.
Outputs go only into oracle/_ref/ (git-ignored, but shipped to the GPU box).  TEST INFRASTRUCTURE."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/solvshift/src'
FILES = ['efprot.f', 'solpol2.f', 'shftex.f', 'exrep.f']
EXTRA = [('/root/reference/util/lib/genres.f', {'GTCF'})]     # (path, units to translate)


def build(force=False):
    out = os.path.join(HERE, '_ref')
    lib = os.path.join(out, 'libslvref.so')
    if not os.path.isdir(REF_SRC):
        return lib if os.path.isfile(lib) else None      # GPU box: use the prebuilt library
    os.makedirs(out, exist_ok=True)
    sys.path.insert(0, HERE)
    import f77c
    cs = []
    for f in FILES:
        with open(os.path.join(REF_SRC, f), encoding='utf-8', errors='replace') as fh:
            text = fh.read()
        f77c.CALLED.clear()
        c = os.path.join(out, f.replace('.f', '_f.c'))
        with open(c, 'w') as fh:
            fh.write(f77c.transpile(text))
        cs.append(c)
    for path, only in EXTRA:
        with open(path, encoding='utf-8', errors='replace') as fh:
            text = fh.read()
        f77c.CALLED.clear()
        c = os.path.join(out, os.path.basename(path).replace('.f', '_f.c'))
        with open(c, 'w') as fh:
            fh.write(f77c.transpile(text, only=only))
        cs.append(c)
    cmd = ['gcc', '-O2', '-fPIC', '-shared', '-w', '-o', lib] + cs + [os.path.join(HERE, 'f77_support.c'), '-lm']
    subprocess.check_call(cmd)
    return lib


if __name__ == '__main__':
    print(build(force=True))
