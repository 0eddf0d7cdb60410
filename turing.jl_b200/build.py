"""Builds libtnuts.so (sm_100a) in-tree with nvcc.  No torch involved: the product is a C-ABI
shared library; PyTorch is only used by the host side for process-group plumbing."""
import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libtnuts.so")

# per-file extra flags (none at present)
EXTRA_FLAGS = {}

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _nccl_paths():
    """header + library of NCCL (torch-bundled wheel first, system second)."""
    cands = []
    for base in sys.path:
        d = os.path.join(base, "nvidia", "nccl")
        if os.path.isdir(d):
            cands.append((os.path.join(d, "include"), os.path.join(d, "lib")))
    cands.append(("/usr/include", "/usr/lib/x86_64-linux-gnu"))
    for inc, lib in cands:
        libs = glob.glob(os.path.join(lib, "libnccl.so*"))
        if os.path.exists(os.path.join(inc, "nccl.h")) and libs:
            return inc, sorted(libs)[0]
    return None, None


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = glob.glob(os.path.join(CSRC, "*")) + [os.path.join(HERE, "..", "include", "tnuts.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    inc, ncclib = _nccl_paths()
    flags = list(NVCC_FLAGS)
    if inc:
        flags += ["-I", inc]
    if verbose:
        flags += ["-Xptxas", "-v"]
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "tnuts.h")]
    hdr_t = max(os.path.getmtime(h) for h in hdrs)

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t)):
            return obj, ""
        extra = EXTRA_FLAGS.get(os.path.basename(src), [])
        r = subprocess.run([nvcc, *flags, *extra, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
        return obj, r.stderr

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(compile_one, sources()))
    objs = [o for o, _ in res]
    if verbose:
        for _, log in res:
            sys.stderr.write(log)
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")
    link = ["g++", "-shared", "-o", LIB, *objs, "-L", cuda_lib, "-lcudart_static", "-ldl", "-lrt", "-lpthread",
            ]
    if ncclib:
        # link the exact file (torch's wheel only ships libnccl.so.2) and leave an rpath
        link += ["-L", os.path.dirname(ncclib), "-l:" + os.path.basename(ncclib),
                 "-Wl,-rpath," + os.path.dirname(ncclib)]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
