"""CPU: libnsff_b200.so builds for sm_100a, loads, and exports every symbol include/nsff_b200.h declares
(no compute calls - there is no GPU here).  Also checks the ctypes struct mirrors against the C layout."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nsff_b200.h")


@pytest.fixture(scope="module")
def libpath():
    from nsff_pl_b200 import build
    return build.build()          # no-op when the in-tree .so is current; nvcc cross-compiles otherwise


def declared_symbols():
    txt = open(HEADER).read()
    return sorted(set(re.findall(r"NSFF_API\s+[\w\s\*]+?\b(nsff_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    assert len(syms) == 40, syms
    for must in ("nsff_field_forward", "nsff_composite", "nsff_sample_points", "nsff_sample_pdf", "nsff_field_pack",
                 "nsff_field_backward_fp32", "nsff_composite_backward"):
        assert must in syms


def test_library_exports_every_declared_symbol(libpath):
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\sT\s+(nsff_[a-z0-9_]+)", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    extra = [s for s in exported if s not in declared_symbols()]
    assert not extra, f"exported but not declared: {extra}"


def test_ctypes_binding_covers_the_header_and_loads(libpath):
    from nsff_pl_b200 import _cabi
    assert sorted(_cabi.PROTOTYPES) == declared_symbols()
    lib = _cabi.lib()
    assert lib.nsff_abi_version() == 1
    assert lib.nsff_error_string(-2).decode().startswith("configuration not supported")
    assert lib.nsff_launch_count() == 0      # nothing has run: no GPU in this container


def test_library_contains_blackwell_tensor_core_code(libpath):
    """tcgen05.mma / tcgen05.ld / bulk-copy appear as UTCHMMA / LDTM / UBLKCP in the sm_100a SASS."""
    sass = subprocess.run(["cuobjdump", "-sass", libpath], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass or "SM100" in sass.upper()
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS"
    assert "HMMA.16816" not in sass, "legacy mma.sync path must not be present"


def test_struct_layouts_match_the_c_header(libpath, tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirrors."""
    from nsff_pl_b200 import _cabi
    src = tmp_path / "layout.c"
    fields = {
        "nsff_field_desc": ["W", "flow_scale"],
        "nsff_field_params": ["static_w", "static_final_w", "trans_w", "trans_bw_b"],
        "nsff_points": ["xyz", "offset_ld", "xyz_out", "t_shift", "a_table", "zs", "B", "S", "xyz_fw_out", "xyz_bw_out"],
        "nsff_composite_args": ["R", "S", "raw_ld", "noise_std", "vis", "raw_w_ld", "xyz_fw", "static_sigmas",
                                "disoccs_bw", "depth", "disocc_bw"],
        "nsff_composite_grads": ["static_sigmas", "xyzs_fw_bw", "depth", "xyz_bw_ray", "rgb_bw", "d_raw", "d_raw_bw"],
        "nsff_field_grads": ["static_w", "static_dir_b", "trans_w", "trans_bw_b"],
        "nsff_interpolate_args": ["R", "S", "h", "dt", "cy", "P", "xyzs", "static_alpha", "ld_t", "rgb_tp1", "ld_tp1"],
        "nsff_loss_args": ["R", "S", "n_reg", "max_t", "lambda_geo_d", "cross_entropy_w", "rgb_fine", "tw", "disoccs_bw",
                           "xyzs_bw", "rgbs", "cam_ids", "Ps", "medians", "terms", "g_terms", "d_rgb_fine", "d_xyzs_bw"],
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "nsff_b200.h"', "int main(void){"]
    for st, fs in fields.items():
        lines.append(f'printf("{st} %zu\\n", sizeof({st}));')
        for f in fs:
            lines.append(f'printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    lines.append("return 0;}")
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.rsplit(" ", 1) for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    mirror = {"nsff_field_desc": _cabi.FieldDesc, "nsff_field_params": _cabi.FieldParams,
              "nsff_points": _cabi.Points, "nsff_composite_args": _cabi.CompositeArgs,
              "nsff_composite_grads": _cabi.CompositeGrads, "nsff_field_grads": _cabi.FieldGrads,
              "nsff_interpolate_args": _cabi.InterpolateArgs, "nsff_loss_args": _cabi.LossArgs}
    for st, cls in mirror.items():
        assert int(got[st]) == C.sizeof(cls), st
        for f in fields[st]:
            assert int(got[f"{st}.{f}"]) == getattr(cls, f).offset, f"{st}.{f}"
