"""Build libsdb200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m slotdiffusion_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the
repo snapshot (it is git-ignored, not gpurun-ignored).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SOURCES = ['api.cu', 'gemm.cu', 'elementwise.cu', 'attention.cu', 'attention_tc.cu', 'attention_fewkeys.cu', 'slot_attention.cu', 'slot_attention_fused.cu', 'slot_attention_resident.cu', 'slot_update.cu', 'backward.cu', 'boundary.cu', 'predictor.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--use_fast_math=false']
# Build variants live next to the product library under their own names (they travel to the GPU box with it and are
# selected at run time with SDB_LIB=<path>, see _lib.py); the product build is always libsdb200.so.
VARIANT = ''
# SDB_SF_EXPERIMENTAL=0: the earlier form of the fused attend kernel's wait loops / converter (csrc/slot_attention_fused.cu;
# the macro's default is 1 since both changes were verified and timed on the B200)
if os.environ.get('SDB_SF_EXPERIMENTAL', '1') == '0':
    NVCC_FLAGS.append('-DSDB_SF_EXPERIMENTAL=0')
    VARIANT += '_sfold'
# diagnostic build: the GEMM's issuer accounts its mbarrier wait cycles (csrc/gemm.cu, sdb_gemm_timing)
if os.environ.get('SDB_GEMM_TIMING', '0') == '1':
    NVCC_FLAGS.append('-DSDB_GEMM_TIMING=1')
    VARIANT += '_gtiming'
# ad-hoc experiment builds: SDB_DEFINES="-DFOO=1 -DBAR" SDB_VARIANT=name -> libsdb200_name.so
if os.environ.get('SDB_DEFINES') and os.environ.get('SDB_VARIANT'):
    NVCC_FLAGS += os.environ['SDB_DEFINES'].split()
    VARIANT += '_' + os.environ['SDB_VARIANT']
LIB = os.path.join(HERE, f'libsdb200{VARIANT}.so')
STAMP = os.path.join(HERE, f'.libsdb200{VARIANT}.stamp')
BUILD_DIR = os.path.join(HERE, 'build' + VARIANT)


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def _digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join('..', '..', 'include', 'sdb200.h')]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, 'rb').read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=True):
    """Compile every CUDA source for sm_100a into libsdb200.so (no-op if up to date)."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    objs = []
    procs = []
    os.makedirs(BUILD_DIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if not f.startswith('--use_fast_math')]
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(BUILD_DIR, src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [_nvcc()] + flags + ['-c', sp, '-o', obj]
        if verbose:
            print(' '.join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
        if verbose and out.strip():
            print(out)
    cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lpthread', '-ldl']
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    with open(STAMP, 'w') as f:
        f.write(dig)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)
    print('built', LIB)
