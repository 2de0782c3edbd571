"""Stage and import the UNMODIFIED reference (kwea123/nsff_pl) for CPU checks and the CPU bench arm.  TEST INFRASTRUCTURE ONLY.

The reference is Python, so "building" it means copying the handful of files of the hot path (and its direct
consumer, losses.py) from /root/reference into the git-ignored, gpurun-shipped ``baseline/_ref/``:

    models/__init__.py  models/nerf.py  models/rendering.py  models/softsplat.py
    datasets/ray_utils.py  losses.py            (+ an EMPTY datasets/__init__.py: the original imports the
                                                  dataset classes, which need cv2 / COLMAP readers)

``stage()`` runs in the build container (``__graft_entry__.build()``); nothing here is committed and the product
package never imports this module.  ``load()`` imports the staged files with in-memory shims for the two absent
third-party imports that the hot path never executes:

    kornia.create_meshgrid            (used by interpolate() / ray generation only)
    kornia.filters.filter2d           (losses.py:3; cross-correlation with zero padding, as kornia does it)
    cupy                              (models/softsplat.py:3; only the CUDA splat op uses it)

The staged modules are imported under their own names (``models``, ``datasets``, ``losses``) in an isolated
sys.modules / sys.path window and handed back in a namespace, so they never clash with
``nsff_pl_b200.install_as_reference_modules()``.
"""
import importlib
import os
import shutil
import sys
import types
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.environ.get("NSFF_REFERENCE", "/root/reference")
REF_DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["models/__init__.py", "models/nerf.py", "models/rendering.py", "models/softsplat.py",
         "datasets/ray_utils.py", "losses.py"]


def available():
    return all(os.path.exists(os.path.join(REF_DST, f)) for f in FILES)


def stage(force=False):
    """Copy the reference's hot-path files into baseline/_ref/ (build container only).  Returns True if staged."""
    if not os.path.isdir(REF_SRC):
        return available()
    for f in FILES:
        src, dst = os.path.join(REF_SRC, f), os.path.join(REF_DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or open(src, "rb").read() != open(dst, "rb").read():
            shutil.copyfile(src, dst)
    init = os.path.join(REF_DST, "datasets", "__init__.py")
    if not os.path.exists(init):
        open(init, "w").close()
    return True



def _nvrtc_compile_with_cache(source):
    """What models/softsplat.py:200-202 asks of cupy: compile a CUDA source string and hand back an object whose
    get_function(name) returns a callable launched as f(grid=..., block=..., args=[...]).  Implemented with NVRTC and
    the driver API from cuda-python, so that the reference's OWN splat kernel string (softsplat.py:6-44, specialised by
    its own cupy_kernel(), softsplat.py:148-196) can be executed unmodified on a GPU box that has no cupy
    (tests/golden/make_splat_golden.py).  Only ever called when the reference's interpolate() runs on CUDA tensors."""
    import ctypes
    import torch
    try:
        from cuda.bindings import driver as cu, nvrtc
    except ImportError:                               # older cuda-python layout
        from cuda import cuda as cu, nvrtc

    def ok(res):
        err, *rest = res
        if int(err) != 0:
            raise RuntimeError(f"cuda-python call failed: {err}")
        return rest[0] if len(rest) == 1 else rest

    major, minor = torch.cuda.get_device_capability()
    prog = ok(nvrtc.nvrtcCreateProgram(source.encode(), b"softsplat.cu", 0, [], []))
    opts = [f"--gpu-architecture=compute_{major}{minor}".encode()]
    res = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if int(res[0]) != 0:
        log = b" " * ok(nvrtc.nvrtcGetProgramLogSize(prog))
        nvrtc.nvrtcGetProgramLog(prog, log)
        raise RuntimeError("NVRTC failed:\n" + log.decode(errors="replace"))
    ptx = b" " * ok(nvrtc.nvrtcGetPTXSize(prog))
    ok(nvrtc.nvrtcGetPTX(prog, ptx))
    module = ok(cu.cuModuleLoadData(ptx))

    class _Module:
        def get_function(self, name):
            fn = ok(cu.cuModuleGetFunction(module, name.encode()))

            def launch(grid, block, args):
                vals = tuple(int(a) for a in args)
                types_ = tuple([ctypes.c_int] + [ctypes.c_void_p] * (len(vals) - 1))   # (n, pointers...)
                stream = torch.cuda.current_stream().cuda_stream
                ok(cu.cuLaunchKernel(fn, grid[0], grid[1], grid[2], block[0], block[1], block[2], 0, stream,
                                     (vals, types_), 0))
            return launch

    return _Module()


def _shims():
    import torch
    import torch.nn.functional as F

    kornia = types.ModuleType("kornia")

    def create_meshgrid(h, w, normalized_coordinates=True, device="cpu"):
        xs = torch.linspace(0, w - 1, w, device=device)
        ys = torch.linspace(0, h - 1, h, device=device)
        if normalized_coordinates:
            xs, ys = (xs / (w - 1) - 0.5) * 2, (ys / (h - 1) - 0.5) * 2
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        return torch.stack([gx, gy], -1)[None]

    def filter2d(input, kernel, border_type="reflect", normalized=False):
        # kornia.filters.filter2d: per-channel cross-correlation; padding (k-1)//2 in front, the rest behind
        b, c, h, w = input.shape
        k = kernel.to(input)
        if normalized:
            k = k / k.abs().sum(dim=(-2, -1), keepdim=True)
        kh, kw = k.shape[-2:]
        pad = [(kw - 1) // 2, kw - 1 - (kw - 1) // 2, (kh - 1) // 2, kh - 1 - (kh - 1) // 2]
        x = F.pad(input, pad, mode=border_type)
        wgt = k.reshape(-1, 1, kh, kw).expand(c, 1, kh, kw) if k.shape[0] == 1 else k.reshape(c, 1, kh, kw)
        return F.conv2d(x, wgt, groups=c)

    filters = types.ModuleType("kornia.filters")
    filters.filter2d = filter2d
    kornia.create_meshgrid = create_meshgrid
    kornia.filters = filters
    cupy = types.ModuleType("cupy")
    cupy.memoize = lambda **kw: (lambda f: f)
    cupy.ndarray = type("ndarray", (), {})          # einops probes sys.modules['cupy'].ndarray
    cupy.cuda = types.SimpleNamespace(compile_with_cache=_nvrtc_compile_with_cache)
    return {"kornia": kornia, "kornia.filters": filters, "cupy": cupy}


_CACHE = None


def load():
    """Import the staged reference.  Returns a namespace with .nerf, .rendering, .ray_utils, .losses, .softsplat."""
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    if not available():
        if not stage():
            raise RuntimeError("the reference is not staged (baseline/_ref is empty and /root/reference is absent)")
    taken = ("models", "datasets", "losses", "kornia", "cupy")
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in taken}
    for k in saved:
        del sys.modules[k]
    sys.modules.update(_shims())
    sys.path.insert(0, REF_DST)
    try:
        warnings.filterwarnings("ignore", category=SyntaxWarning)     # regex literals in models/softsplat.py:152-184
        ns = types.SimpleNamespace(
            nerf=importlib.import_module("models.nerf"),
            rendering=importlib.import_module("models.rendering"),
            softsplat=importlib.import_module("models.softsplat"),
            ray_utils=importlib.import_module("datasets.ray_utils"),
            losses=importlib.import_module("losses"))
    finally:
        sys.path.remove(REF_DST)
        for k in [k for k in sys.modules if k.split(".")[0] in taken]:
            del sys.modules[k]
        sys.modules.update(saved)
    _CACHE = ns
    return ns


def build_reference_fields(models, emb):
    """Reference nn.Modules (models/nerf.py:33-116) carrying the weights of nsff_pl_b200 modules (same state_dict
    names), plus the reference's embedding dict, all on CPU."""
    import torch
    ref = load()
    out = {}
    for key, m in models.items():
        r = ref.nerf.NeRF(m.typ, D=m.D, W=m.W, skips=list(m.skips), in_channels_xyz=m.in_channels_xyz,
                          use_viewdir=m.use_viewdir, in_channels_dir=m.in_channels_dir,
                          encode_appearance=bool(m.encode_appearance), in_channels_a=m.in_channels_a,
                          encode_transient=bool(m.encode_transient), in_channels_t=m.in_channels_t,
                          output_flow=bool(m.output_flow), flow_scale=float(getattr(m, "flow_scale", 0.2)))
        r.load_state_dict({k: v.detach().float().cpu() for k, v in m.state_dict().items()})
        out[key] = r.requires_grad_(False)
    remb = {"xyz": ref.nerf.PosEmbedding(9, 10), "dir": ref.nerf.PosEmbedding(3, 4)}
    for k in ("t", "a"):
        if k in emb:
            e = torch.nn.Embedding(*emb[k].weight.shape)
            e.weight.data.copy_(emb[k].weight.detach().float().cpu())
            remb[k] = e.requires_grad_(False)
    return out, remb


if __name__ == "__main__":
    print("staged" if stage(force="--force" in sys.argv) else "reference not available", "->", REF_DST)
