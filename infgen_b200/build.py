"""Build libinfgen_b200.so for sm_100a (in-tree, so the .so travels with the repository snapshot).

    python -m infgen_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'lib', 'libinfgen_b200.so')
SOURCES = ['engine.cu']
DEPS = ['engine.cu', 'common.cuh', 'stream.cuh', 'ops.cuh', 'layer.cuh', 'decode.cuh', 'insert.cuh', 'fourier_tc.cuh', 'node.cuh', 'node_tc.cuh', 'map.cuh', 'prep.cuh', os.path.join('..', '..', 'include', 'infgen_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-shared',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(SRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', OUT] + [os.path.join(SRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(os.path.dirname(OUT), 'build.log'), 'w') as f:
        f.write(' '.join(cmd) + '\n' + log)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + log[-4000:])
    if verbose:
        print(log)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
