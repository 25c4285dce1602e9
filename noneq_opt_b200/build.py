"""In-tree build of libnoneq_b200.so (nvcc, sm_100a only).

`python -m noneq_opt_b200.build` or `__graft_entry__.build()`.  The shared object is written
next to the sources (git-ignored, travels to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libnoneq_b200.so')
STAMP = os.path.join(HERE, 'csrc', '.build_stamp')
SOURCES = ['core.cu', 'langevin.cu', 'langevin_pot_spring.cu', 'langevin_pot_biomolecule.cu', 'langevin_pot_shifted.cu',
           'ising.cu', 'train.cu']
DEV_SOURCES = ['devtools.cu']   # -> libnoneq_b200_dev.so (microbenchmark, math checker; not the product ABI)
DEV_LIB = os.path.join(HERE, 'libnoneq_b200_dev.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr'] + os.environ.get('NQ_EXTRA_NVCC_FLAGS', '').split()


def _nvcc():
  nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
  if not os.path.exists(nvcc):
    raise RuntimeError('nvcc not found; libnoneq_b200.so cannot be built')
  return nvcc


def _digest():
  h = hashlib.sha256()
  for root in (CSRC, os.path.join(HERE, '..', 'include')):
    for name in sorted(os.listdir(root)):
      if name.endswith(('.cu', '.cuh', '.h')):
        with open(os.path.join(root, name), 'rb') as f:
          h.update(name.encode())
          h.update(f.read())
  h.update(' '.join(NVCC_FLAGS).encode())
  return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
  digest = _digest()
  if not force and os.path.exists(LIB) and os.path.exists(DEV_LIB) and os.path.exists(STAMP):
    with open(STAMP) as f:
      if f.read().strip() == digest:
        return LIB
  nvcc = _nvcc()
  objs = []
  procs = []
  for src in SOURCES + DEV_SOURCES:
    path = os.path.join(CSRC, src)
    if not os.path.exists(path):
      continue
    obj = os.path.join(CSRC, src.replace('.cu', '.o'))
    cmd = [nvcc, *NVCC_FLAGS, '-c', path, '-o', obj]
    if verbose:
      cmd.insert(1, '-Xptxas')
      cmd.insert(2, '-v')
    procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs.append(obj)
  for src, p in procs:
    out, _ = p.communicate()
    if verbose and out:
      print(out)
    if p.returncode:
      raise RuntimeError(f'nvcc failed on {src}:\n{out}')
  dev_objs = [o for o in objs if os.path.basename(o) in [d.replace('.cu', '.o') for d in DEV_SOURCES]]
  objs = [o for o in objs if o not in dev_objs]
  for lib, members in ((LIB, objs), (DEV_LIB, dev_objs)):
    cmd = [nvcc, '-shared', '-o', lib, *members, '-gencode', 'arch=compute_100a,code=sm_100a', '-ldl']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
      raise RuntimeError(f'link failed:\n{r.stdout}')
  with open(STAMP, 'w') as f:
    f.write(digest)
  return LIB


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
