"""Build libsurmise_b200.so in-tree with nvcc for sm_100a (no GPU needed: cross-compiles)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, 'surmise_b200', 'csrc')
OUT = os.path.join(ROOT, 'surmise_b200', 'lib', 'libsurmise_b200.so')
SOURCES = ['capi.cu', 'dgemm.cu', 'chol.cu', 'covmat.cu', 'gp.cu', 'woodbury.cu', 'nll_small.cu', 'sampler.cu']


STAMP = OUT + '.srchash'


def source_hash():
    import hashlib
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(ROOT, 'include', 'surmise_b200.h')]
    for p in deps:
        h.update(os.path.basename(p).encode())
        with open(p, 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()


def needs_build():
    """Content hash, not mtime: the snapshot shipped to the GPU box does not keep timestamps."""
    if not (os.path.exists(OUT) and os.path.exists(STAMP)):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != source_hash()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
           '-Xcompiler', '-fPIC', '-shared', '-o', OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    subprocess.check_call(cmd)
    with open(STAMP, 'w') as fh:
        fh.write(source_hash())
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
