"""Build the C-ABI CUDA library in-tree with nvcc for sm_100a.

    python -m movenet_b200.build

produces movenet_b200/lib/libmovenet_b200.so (git-ignored, shipped to the GPU
box with the snapshot).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libmovenet_b200.so")
SOURCES = ["api.cu", "pack.cu", "wavenet.cu", "mulaw.cu", "decode.cu", "layer_tc.cu", "layer_tc_bwd.cu", "layer_tc_bwd_db.cu", "head_tc.cu", "input_tc.cu", "ce.cu", "decode_tc.cu", "upsample_tc.cu", "optim.cu", "wide.cu", "peer.cu", "video_tc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "movenet_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def source_hash():
    """sha256 over every source the library is built from (csrc/*, the public header, the compile flags)"""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(HERE, "..", "include", "movenet_b200.h")]
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS + SOURCES + [os.environ.get("MOVENET_B200_NVCC_EXTRA", "")]).encode())
    return h.hexdigest()


HASH_FILE = os.path.join(LIB_DIR, "SOURCE_HASH")


def built_hash():
    try:
        with open(HASH_FILE) as fh:
            return fh.read().strip()
    except OSError:
        return None


def build(force=False, verbose=False):
    """(Re)build when the library is missing, older than a source, or was built from different sources: the hash of the
    sources a prebuilt .so came from is stored next to it and logged, so a stale shipped binary is detectable."""
    want = source_hash()
    if not force and not _stale() and built_hash() == want:
        print(f"movenet_b200.build: up to date (sources sha256 {want[:16]})", file=sys.stderr)
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        # MOVENET_B200_NVCC_EXTRA: extra compile flags for instrumented builds (e.g. -DMVN_PHASE_CLOCKS=1)
        extra = os.environ.get("MOVENET_B200_NVCC_EXTRA", "").split()
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart", "-lcuda", "-ldl"]
    subprocess.run(cmd, check=True)
    with open(HASH_FILE, "w") as fh:
        fh.write(want + "\n")
    print(f"movenet_b200.build: compiled {len(SOURCES)} sources for sm_100a (sources sha256 {want[:16]})", file=sys.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
