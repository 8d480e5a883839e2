"""Build libgcrnn_b200.so in-tree with nvcc for sm_100a:  python -m gated_gcrnns_b200.build"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = ['api.cu', 'graph.cu', 'gcrnn_f32.cu', 'gcrnn_tc.cu', 'builders.cu']
OUT = os.path.join(HERE, 'libgcrnn_b200.so')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '--use_fast_math=false',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '-shared']


def needs_build():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, 'csrc', f) for f in os.listdir(os.path.join(HERE, 'csrc'))]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'gcrnn_b200.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    flags = [f for f in FLAGS if not f.startswith('--use_fast_math')]
    cmd = [nvcc] + flags + ['-o', OUT] + [os.path.join(HERE, 'csrc', s) for s in SRC] + ['-lcudart', '-lcuda', '-ldl']
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv)
