"""Builds libbhr.so (the sm_100a CUDA kernels + C-ABI) in-tree with nvcc.

Each translation unit is compiled to an object file under csrc/_obj/ (in parallel, only when it
or a header is newer), then linked; nvcc cross-compiles without a GPU."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
SOURCES = ["api.cu", "raymarch.cu", "post.cu", "bloom.cu", "texture.cu", "background.cu", "stats.cu", "peer.cu", "png.cu"]
EXTRA_FLAGS = {}          # per-file nvcc flags
LIB = os.path.join(HERE, "libbhr.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [
        os.path.join(HERE, "..", "include", "bhr.h")]


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in _sources()] + _headers()
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into black_hole_renderer_b200/libbhr.so."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("BHR_NVCC_EXTRA", "").split()
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    tag = os.path.join(OBJ, ".flags")
    flags_now = " ".join(NVCC_FLAGS + extra)
    if not os.path.exists(tag) or open(tag).read() != flags_now:
        force = True

    def compile_one(src):
        s, o = os.path.join(CSRC, src), os.path.join(OBJ, src[:-3] + ".o")
        if not force and os.path.exists(o) and os.path.getmtime(o) > max(os.path.getmtime(s), hdr_t):
            return ""
        cmd = [nvcc] + NVCC_FLAGS + extra + EXTRA_FLAGS.get(src, []) + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", "-o", o, s, "-ccbin", "/usr/bin/g++"]
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        return res.stderr

    with ThreadPoolExecutor(max_workers=8) as pool:
        logs = list(pool.map(compile_one, _sources()))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in _sources()]
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", LIB] + objs +
                         ["-ccbin", "/usr/bin/g++"], capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    open(tag, "w").write(flags_now)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    import sys
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
