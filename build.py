"""Build libsurmise_b200.so in-tree with nvcc for sm_100a (no GPU needed: cross-compiles)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, 'surmise_b200', 'csrc')
OUT = os.path.join(ROOT, 'surmise_b200', 'lib', 'libsurmise_b200.so')
SOURCES = ['capi.cu', 'dgemm.cu', 'chol.cu', 'covmat.cu', 'gp.cu', 'woodbury.cu', 'nll_small.cu', 'sampler.cu', 'pca.cu', 'chol_dag.cu', 'dgemm_tma.cu', 'peer.cu']


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


def _object_key(src, headers):
    import hashlib
    h = hashlib.sha256()
    for p in [src] + headers:
        h.update(os.path.basename(p).encode())
        with open(p, 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def build(force=False, verbose=False):
    """One object per source (compiled in parallel, cached by the content hash of the source and every header),
    then one link.  Only what changed is recompiled."""
    if not force and not needs_build():
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objdir = os.path.join(os.path.dirname(OUT), 'obj')
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h')))
    headers.append(os.path.join(ROOT, 'include', 'surmise_b200.h'))
    flags = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC']
    if verbose:
        flags += ['-Xptxas', '-v']
    jobs, objs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, '%s.%s.o' % (s[:-3], _object_key(src, headers)))
        objs.append(obj)
        if force or verbose or not os.path.exists(obj):
            for old in os.listdir(objdir):
                if old.startswith(s[:-3] + '.') and os.path.join(objdir, old) != obj:
                    os.remove(os.path.join(objdir, old))
            jobs.append([nvcc] + flags + ['-c', src, '-o', obj])
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        list(pool.map(subprocess.check_call, jobs))
    subprocess.check_call([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', OUT] + objs)
    with open(STAMP, 'w') as fh:
        fh.write(source_hash())
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
