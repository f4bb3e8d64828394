"""Builds libcmu_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m contrastive_masked_unet_b200.build          # incremental
    python -m contrastive_masked_unet_b200.build --force

(The implementation lives in `_buildlib` so that importing it never rebinds the package attribute `build`, which is the
reference-style model factory `build(cfg)` of modules.py.)

nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the repo snapshot."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libcmu_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
         '--expt-relaxed-constexpr', '-Xptxas', '-v']


# per-file extra flags: the data pipeline must round like Pillow / numpy (no contracted multiply-adds)
EXTRA_FLAGS = {'data_aug.cu': ['-fmad=false']}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    headers.append(os.path.join(os.path.dirname(HERE), 'include', 'cmu_b200.h'))
    jobs = []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + '.o')
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run([NVCC] + FLAGS + EXTRA_FLAGS.get(os.path.basename(s), []) + ['-c', s, '-o', o],
                           capture_output=True, text=True)
        return s, r

    failed = False
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                failed = True
                sys.stderr.write(f'nvcc failed on {s}\n')
            else:
                with open(os.path.join(OBJ, os.path.basename(s)[:-3] + '.ptxas.log'), 'w') as f:
                    f.write(r.stderr)
    if failed:
        raise RuntimeError('nvcc compilation failed')
    objs = [os.path.join(OBJ, src[:-3] + '.o') for src in sources()]
    if force or jobs or _stale(LIB, objs):
        r = subprocess.run([NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a',
                                                                 '-lcudart_static', '-ldl', '-lrt', '-lpthread'],
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
