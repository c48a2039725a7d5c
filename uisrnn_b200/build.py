"""Builds libuisrnn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libuisrnn_b200.so')
SOURCES = ['uis_api.cu', 'uis_train.cu']
DEPS = ['uis_api.cu', 'uis_train.cu', 'uis_beam.cuh', 'uis_beam_tree.cuh', 'uis_prepass.cuh', 'uis_common.cuh',
        os.path.join('..', '..', 'include', 'uisrnn_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC', '--threads', '0']


def find_nvcc():
  for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
    if cand and os.path.exists(cand):
      return cand
  raise RuntimeError('nvcc not found')


def is_stale():
  if not os.path.exists(LIB):
    return True
  t = os.path.getmtime(LIB)
  return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
  if not force and not is_stale():
    return LIB
  cmd = [find_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
      ['-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES]
  res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
  if res.returncode != 0:
    sys.stderr.write(res.stdout + res.stderr)
    raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
  if verbose:
    sys.stderr.write(res.stderr)
  return LIB


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose=True))
