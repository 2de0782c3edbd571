"""In-tree build of libnsff_b200.so (hand-written sm_100a CUDA + the C ABI of include/nsff_b200.h).

    python -m nsff_pl_b200.build            # or: __graft_entry__.build()

nvcc cross-compiles without a GPU.  The shared object lands next to this file so that it travels
with the repo snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("NSFF_LIB_OUT") or os.path.join(HERE, "libnsff_b200.so")      # NSFF_LIB_OUT: side build (tools/timeline.py)
STAMP = LIB + ".stamp"
SOURCES = ["nsff_points.cu", "nsff_composite.cu", "nsff_composite_bwd.cu", "nsff_resample.cu", "nsff_interpolate.cu", "nsff_loss.cu", "nsff_field_fp32.cu", "nsff_field_tc.cu", "nsff_field_bwd_tc.cu"]
EXTRA = os.environ.get("NSFF_NVCC_EXTRA", "").split()      # e.g. -DNSFF_TIMELINE for tools/timeline.py
NVCC_FLAGS = EXTRA + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
              "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "nsff_b200.h")]
    for f in files:
        with open(f, "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link libnsff_b200.so.  Returns the library path."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    nvcc = nvcc_path()
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", ("x_" if os.environ.get("NSFF_LIB_OUT") else "") + src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed (see output above)")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
