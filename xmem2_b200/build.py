"""Build libxmem2_b200.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libxmem2_b200.so')
SOURCES = ['common.cu', 'k1_affinity.cu', 'conv_igemm.cu', 'conv_igemm_pair.cu', 'pair_dissim.cu', 'consolidate.cu', 'eltwise.cu', 'stem7x7.cu', 'postproc.cu']
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


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
