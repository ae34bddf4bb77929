"""Build the C-ABI CUDA library IN-TREE: ``morgana_b200/lib/libmorgana_b200.so`` (sm_100a only).

    python -m morgana_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the source snapshot.
"""
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, 'csrc')
LIB_DIR = os.path.join(PKG_DIR, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libmorgana_b200.so')
INCLUDE = os.path.join(REPO_ROOT, 'include')

SOURCES = ['mg_core.cu', 'mg_scan.cu', 'mg_upsample.cu', 'mg_collate.cu', 'mg_normalise.cu', 'mg_reduce.cu', 'mg_objective.cu', 'mg_objective_stream.cu', 'mg_ema.cu',
           'mg_linear.cu', 'mg_linear_bwd.cu', 'mg_mlpg.cu', 'mg_segments.cu', 'mg_kld.cu']

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-std=c++17', '-lineinfo',
    '--fmad=false',          # parity: ATen's elementwise kernels round every op; never contract a*b+c behind our back
    '-Xcompiler', '-fPIC,-O2,-Wall',
    '-Xptxas', '-v',
    '-I', INCLUDE, '-I', CSRC,
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def _inputs():
    files = [os.path.join(CSRC, s) for s in SOURCES]
    files += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    files.append(os.path.join(INCLUDE, 'morgana_b200.h'))
    files.append(os.path.abspath(__file__))
    return files


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > built for f in _inputs())


def build(force=False, verbose=False):
    """Compile every kernel for sm_100a into one shared library; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, 'obj')
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, log = [], []
    failed = False
    for src, obj, proc in procs:
        out, _ = proc.communicate()
        log.append('==== %s\n%s' % (src, out))
        if proc.returncode != 0:
            failed = True
            sys.stderr.write('nvcc failed on %s:\n%s\n' % (src, out))
        objs.append(obj)
    with open(os.path.join(LIB_DIR, 'build.log'), 'w') as f:
        f.write('\n'.join(log))
    if failed:
        raise RuntimeError('morgana_b200: CUDA build failed (see above / %s)' % os.path.join(LIB_DIR, 'build.log'))
    if verbose:
        sys.stdout.write('\n'.join(log) + '\n')
    link = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB_PATH + '.tmp'] + objs
    subprocess.run(link, check=True)
    os.replace(LIB_PATH + '.tmp', LIB_PATH)
    return LIB_PATH


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(path)
