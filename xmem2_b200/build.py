"""Build libxmem2_b200.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libxmem2_b200.so')
SOURCES = ['common.cu', 'k1_affinity.cu', 'conv_igemm.cu', 'eltwise.cu', 'postproc.cu']
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '--use_fast_math',
         '-Xcompiler', '-fPIC', '-cudart', 'static'] + os.environ.get('XMEM_EXTRA_NVCC_FLAGS', '').split()


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    headers.append(os.path.join(os.path.dirname(HERE), 'include', 'xmem2_b200.h'))
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(HERE, 'build', s.replace('.cu', '.o'))
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if out.strip():
            print(f'--- {s}\n{out}')
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    if force or procs or _stale(OUT, objs):
        cmd = [nvcc, '-shared', '-cudart', 'static', '-o', OUT] + objs
        subprocess.check_call(cmd)
    return OUT


EXP_OUT = os.path.join(HERE, 'libxmem2_b200_exp.so')
EXP_SOURCES = ['conv_igemm_csk.cu', 'conv_igemm_2cta.cu', 'conv_igemm_mc.cu', 'conv_igemm_halo.cu', 'pair_dissim.cu']


def build_experimental(force=False):
    """libxmem2_b200_exp.so: the round-2 head-start kernels of csrc/experimental/ (four `xm_conv2d_nhwc_<variant>` entry points and
    `xm_pair_dissimilarity`) plus their own copy of common.cu.  NOT part of the product library and not
    built by `build()`; `XMEM_CONV_IMPL=<variant>` makes `lib.conv2d_nhwc` call into it (see lib.py)."""
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    jobs = [(os.path.join(CSRC, 'common.cu'), os.path.join(HERE, 'build', 'exp_common.o'))]
    jobs += [(os.path.join(CSRC, 'experimental', s), os.path.join(HERE, 'build', 'exp_' + s.replace('.cu', '.o'))) for s in EXP_SOURCES]
    procs = []
    for src, obj in jobs:
        if force or _stale(obj, [src] + headers):
            procs.append((src, subprocess.Popen([nvcc] + FLAGS + ['-c', src, '-o', obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if out.strip():
            print(f'--- {src}\n{out}')
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed (experimental)')
    objs = [o for _, o in jobs]
    if force or procs or _stale(EXP_OUT, objs):
        subprocess.check_call([nvcc, '-shared', '-cudart', 'static', '-o', EXP_OUT] + objs)
    return EXP_OUT


if __name__ == '__main__':
    if '--experimental' in sys.argv:
        print(build_experimental(force='--force' in sys.argv))
    else:
        print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
