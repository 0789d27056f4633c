"""Build libfar3d_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m far3d_b200.build [--force]

Objects go to far3d_b200/lib/obj/, the shared library to far3d_b200/lib/libfar3d_sm100.so (git-ignored;
it travels to the GPU box with the gpurun snapshot).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
OBJDIR = os.path.join(LIBDIR, 'obj')
SO = os.path.join(LIBDIR, 'libfar3d_sm100.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
         '--expt-relaxed-constexpr', '-Xptxas', '-v'] + os.environ.get('FAR3D_NVCC_EXTRA', '').split()    # e.g. -DFAR3D_CONV_WAITSTATS


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(HERE, '..', 'include', 'far3d_b200.h'))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force):
    obj = os.path.join(OBJDIR, src[:-3] + '.o')
    sp = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(sp), _deps_mtime()):
        return obj, None
    cmd = [NVCC] + FLAGS + ['-c', sp, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJDIR, src[:-3] + '.ptxas.log')
    with open(log, 'w') as f:
        f.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{r.stderr[-6000:]}')
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in res]
    rebuilt = any(log is not None for _, log in res)
    if rebuilt or force or not os.path.exists(SO):
        cmd = [NVCC, '-shared', '-o', SO] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart_static',
                                                    '-ldl', '-lrt', '-lpthread']
        subprocess.check_call(cmd)
    if verbose:
        for _, log in res:
            if log:
                print(log)
    return SO


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
