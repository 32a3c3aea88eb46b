"""CPU-only: the C-ABI library builds/loads and exports exactly the symbols include/moegan_sm100.h declares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "moegan_sm100.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(moegan_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    import __graft_entry__ as ge
    ge.build()
    lib = pkg._cabi.lib()
    decl = declared_symbols()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(pkg._cabi.SIGNATURES) == decl, "ctypes signature table and header disagree"
    assert lib.moegan_version() == pkg._cabi.ABI_VERSION == 6


def test_sizes_and_error_reporting(pkg):
    C = pkg._cabi
    lib = C.lib()
    assert C.m_cap(4096, 8, 2, False) == 4096 * 2 + 256 * 8
    assert C.m_cap(75, 4, 4, True) == 4 * 256
    assert lib.moegan_plan_ints(1024) == 88 + 8 + 8
    # bad shapes are rejected on the host with a message, nothing is launched (no GPU needed)
    bad = C.Shape(100, 100, 64, 4, 4, 4, 25, 0, 1, 0, 1024)
    rc = lib.moegan_convert_tokens(bad, None, None, None)
    assert rc == -1 and "multiple of 128" in C.last_error()
    s = C.make_shape(128, 128, 64, 4, 4, 2, 64, 0, True, False)
    rc = lib.moegan_convert_tokens(s, None, None, None)
    assert rc == -2 and "NULL" in C.last_error()
    assert lib.moegan_router_bwd_workspace_bytes(s) > 0


def test_ctypes_struct_layouts_match_the_header(pkg):
    """sizeof / field order of the argument structs as a C compiler lays them out (values from g++ on the header)."""
    import ctypes
    C = pkg._cabi
    assert ctypes.sizeof(C.Shape) == 44
    assert ctypes.sizeof(C.BlockLayout) == 216
    assert ctypes.sizeof(C.BlockFwdArgs) == 248 and C.BlockFwdArgs.x_op.offset == 240 and C.BlockFwdArgs.out.offset == 176 and C.BlockFwdArgs.residual.offset == 224
    assert C.BlockFwdArgs.out_is_bf16.offset == 236
    assert ctypes.sizeof(C.BlockBwdArgs) == 296 and C.BlockBwdArgs.x_op.offset == 288 and C.BlockBwdArgs.dx.offset == 160 and C.BlockBwdArgs.workspace_bytes.offset == 272
    assert C.BlockBwdArgs.dout_is_bf16.offset == 280
    s = C.make_shape(4096, 512, 512, 8, 2, 8, 512, 0, False, False)
    lay = C.BlockLayout()
    n = C.lib().moegan_block_fwd_workspace(s, 1, ctypes.byref(lay))
    assert n == lay.total > 0 and lay.plan % 256 == 0 and lay.pre > 0
    n0 = C.lib().moegan_block_fwd_workspace(s, 0, ctypes.byref(lay))
    assert 0 < n0 < n and lay.pre == -1 and lay.fp == -1
    assert C.lib().moegan_block_bwd_workspace(s) > 0
