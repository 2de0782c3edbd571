"""Host-only invariants of the per-tile program ProgramBuilder emits for the backward chain (tools/dump_program.cu
compiles nsff_field_tc.cu's builder into a host tool; no GPU is touched).

The barrier protocol of the tile machinery waits by PARITY, so the MMA issuer may only ever be one phase behind any barrier
it waits on.  For the "hidden half published" barriers that means: every hidden layer's publish of either half must be
waited for by some op before the NEXT hidden layer completes - also when nothing reads that layer (the first layer of a
branch in the backward chain).  A missing wait once let the issuer run a layer ahead of the epilogue (hang on the GPU)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "bin", "dump_program")


@pytest.fixture(scope="module")
def tool():
    nvcc = "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    os.makedirs(os.path.dirname(TOOL), exist_ok=True)
    src = os.path.join(ROOT, "tools", "dump_program.cu")
    newest = max(os.path.getmtime(os.path.join(ROOT, "nsff_pl_b200", "csrc", f)) for f in os.listdir(os.path.join(ROOT, "nsff_pl_b200", "csrc")))
    if not os.path.exists(TOOL) or os.path.getmtime(TOOL) < max(newest, os.path.getmtime(src)):
        subprocess.run([nvcc, "-std=c++17", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a", "-w",
                        "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "nsff_pl_b200", "csrc"), "-o", TOOL, src],
                       check=True, cwd=ROOT)
    return TOOL


def _program(tool, *args):
    out = subprocess.run([tool, *map(str, args)], check=True, capture_output=True, text=True).stdout.splitlines()
    assert out[0].startswith("rc 0"), out[0]
    ops = []
    for line in out:
        m = re.match(r"op\s+(\d+) acc (\d+) d_col\s+(\d+) .*? (SS|TS) (.*?)wait_hid (\d) wait_empty 0x([0-9a-f]+) bytes (\d+)", line)
        if m:
            flags = m.group(5).split()
            ops.append(dict(acc=int(m.group(2)), ts=m.group(4) == "TS", first="FIRST" in flags, last="LAST" in flags,
                            adv="ADV" in flags, wait_hid=int(m.group(6)), wait_empty=int(m.group(7), 16), bytes=int(m.group(8))))
    subs = [line for line in out if line.startswith("sub ")]
    return subs, ops


@pytest.mark.parametrize("cfg", [(1, 0, 1), (0, 1, 1), (1, 1, 1), (0, 0, 1), ("fwd", 15, 1), ("fwd", 15, 0, 1), ("fwd", 10, 1),
                                 ("fwd", 1, 1), ("fwd", 19, 1)],
                         ids=["main", "warp", "main+pos", "warp-no-pos", "fwd-inference", "fwd-train", "fwd-warp", "fwd-static",
                              "fwd-sigma-only"])
def test_every_hidden_publish_is_waited_for_before_the_next_one(tool, cfg):
    subs, ops = _program(tool, *cfg)
    assert ops and subs
    two_tiles = ops + ops                                   # the table is replayed tile after tile
    adv = [i for i, o in enumerate(two_tiles) if o["adv"]]
    for a, b in zip(adv, adv[1:]):                          # ops after layer L's last op, up to layer L+1's last op
        waited = {o["wait_hid"] for o in two_tiles[a + 1:b + 1]}
        assert {1, 2} <= waited, f"publish of the layer ending at op {a % len(ops)} is not waited for (waits seen: {waited})"


@pytest.mark.parametrize("cfg", [(1, 0, 1), (0, 1, 1)], ids=["main", "warp"])
def test_program_shape(tool, cfg):
    subs, ops = _program(tool, *cfg)
    # hidden accumulators are never waited for explicitly (their drain is implied by the publish waits); the narrow
    # accumulators of the fp32 row blocks (ids 4, 5) are
    for o in ops:
        assert o["wait_empty"] & 0x0F == 0, o
    assert any(o["wait_empty"] & 0x30 for o in ops)
    # merged N = 256 ops stage one 16 KB piece per CTA; half ops two 8 KB pieces; row blocks 4 KB pieces
    assert {o["bytes"] for o in ops} <= {16384, 8192, 4096}
    # every sub-step is completed by exactly one LAST op
    assert sum(o["last"] for o in ops) == len(subs)
