"""Build libmdevaluate_b200.so (sm_100a) in-tree with nvcc.

    python -m mdevaluate_b200.csrc.build [--force] [--verbose]

The shared library lands in mdevaluate_b200/_cuda/ (git-ignored, shipped to the GPU box with the
repository snapshot).  nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT_DIR = os.path.join(PKG, '_cuda')
LIB = os.path.join(OUT_DIR, 'libmdevaluate_b200.so')
# (source, object name, extra defines): corr_kernels.cu is compiled once per part so that its instantiations build
# in parallel (0 direct-load kernels, 1-3 bulk kernels without / with byte / with packed image counts)
UNITS = [('abi.cu', 'abi.o', []), ('corr_host.cu', 'corr_host.o', []), ('comm.cu', 'comm.o', []),
         ('nojump_kernels.cu', 'nojump_kernels.o', []), ('dense_kernels.cu', 'dense_kernels.o', []),
         ('vector_kernels.cu', 'vector_kernels.o', []), ('overlap_kernels.cu', 'overlap_kernels.o', [])] + [
    ('corr_kernels.cu', 'corr_kernels_%d.o' % part, ['-DMDB_CORR_PART=%d' % part]) for part in range(4)]
SOURCES = sorted({u[0] for u in UNITS})
HEADERS = ['common.cuh', 'corr_core.cuh', os.path.join('..', '..', 'include', 'mdevaluate_b200.h')]
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']


def nvcc_path():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=/path/to/nvcc)')


def is_stale():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, 'build')
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = nvcc_path()
    common = [nvcc, '-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--ftz=false', '--prec-div=true',
              '--prec-sqrt=true', '--fmad=true'] + ARCH
    if os.environ.get('MDB_NVCC_DEFINES'):
        common += ['-D' + d for d in os.environ['MDB_NVCC_DEFINES'].split(',')]
    if verbose:
        common += ['-Xptxas', '-v']
    procs = []
    objs = []
    for src, obj_name, defines in UNITS:
        obj = os.path.join(obj_dir, obj_name)
        objs.append(obj)
        cmd = common + defines + ['-c', os.path.join(HERE, src), '-o', obj]
        procs.append((obj_name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            failed = True
            sys.stderr.write('nvcc failed for %s:\n%s\n' % (src, out))
        elif verbose:
            sys.stdout.write(out)
    if failed:
        raise RuntimeError('nvcc compilation failed')
    link = [nvcc, '-shared', '-o', LIB] + objs + ARCH + ['-lcudart_static', '-ldl', '-lrt', '-lpthread']
    subprocess.run(link, check=True)
    return LIB


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(path)
