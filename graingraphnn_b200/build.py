"""Build libgraingnn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIBNAME = 'libgraingnn_b200.so'
SOURCES = ['api.cu', 'csr.cu', 'edge.cu', 'gather.cu', 'gather_tiled.cu', 'geometry.cu', 'raster.cu', 'topology.cu', 'events.cu', 'gemm_simt.cu', 'gemm_tc.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-Wall']


def lib_path():
    """GG_LIB=<path> selects another build of the same library (A/B measurements of kernel variants, scripts/ab_gather.sh)."""
    return os.environ.get('GG_LIB') or os.path.join(LIBDIR, LIBNAME)


def build_variant(name, defines=(), replace=None, verbose=False):
    """A second library lib/variants/<name>.so that differs from the main one by -D defines (applied to every source) or
    by replaced source files ({'gather_tiled.cu': '/path/to/other.cu'}); objects go to lib/variants/<name>/."""
    import tempfile
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    os.makedirs(os.path.join(LIBDIR, 'variants'), exist_ok=True)
    vdir = tempfile.mkdtemp(prefix=f'gg_{name}_')        # objects stay out of the tree (the tree travels to the GPU box)
    objs = []
    for s in SOURCES:
        src = (replace or {}).get(s, os.path.join(CSRC, s))
        o = os.path.join(vdir, s[:-3] + '.o')
        cmd = [nvcc] + NVCC_FLAGS + ['-I', CSRC] + [f'-D{d}' for d in defines] + ['-c', '-o', o, src]
        if verbose:
            print(' '.join(cmd))
        subprocess.run(cmd, check=True, stderr=None if verbose else subprocess.DEVNULL)
        objs.append(o)
    out = os.path.join(LIBDIR, 'variants', name + '.so')
    subprocess.run([nvcc, '-Wno-deprecated-gpu-targets', '-shared', '-o', out] + objs + ['-lcuda'], check=True)
    shutil.rmtree(vdir, ignore_errors=True)
    return out


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library. Returns its path."""
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(HERE, '..', 'include', 'graingnn_b200.h'))
    objs = []
    for s in srcs:
        o = os.path.join(LIBDIR, os.path.basename(s)[:-3] + '.o')
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + (['-DGG_TILED_PROFILE'] if os.environ.get('GG_TILED_PROFILE') else []) + [f'-D{d}' for d in os.environ.get('GG_DEFINES', '').split()] + ['-c', '-o', o, s]
            if verbose:
                print(' '.join(cmd))
            subprocess.run(cmd, check=True)
        objs.append(o)
    out = lib_path()
    if force or _stale(out, objs):
        cmd = [nvcc, '-Wno-deprecated-gpu-targets', '-shared', '-o', out] + objs + ['-lcuda']
        if verbose:
            print(' '.join(cmd))
        subprocess.run(cmd, check=True)
    return out


if __name__ == '__main__':
    print(build(force=True, verbose=True))
