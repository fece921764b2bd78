"""The C-ABI library loads without a GPU and exports every symbol the header
declares (no compute calls here)."""

import os
import re

from conftest import ROOT
from cup1d_b200 import _abi


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "cup1d_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(c1d_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_header_symbols():
    if not os.path.exists(_abi.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    lib = _abi.load_library()
    syms = _header_symbols()
    assert set(syms) == set(_abi.SYMBOLS), (syms, _abi.SYMBOLS)
    for s in syms:
        assert getattr(lib, s) is not None
    assert lib.c1d_abi_version() == _abi.ABI_VERSION


def test_config_struct_matches_header():
    """field order of the ctypes mirror == order in the C struct"""
    txt = open(os.path.join(ROOT, "include", "cup1d_b200.h")).read()
    start = txt.index("typedef struct c1d_config {") + len("typedef struct c1d_config {")
    body = txt[start : txt.index("} c1d_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        m = re.match(r"(int32_t|double)\s+(.*)", decl, flags=re.S)
        if not m:
            continue
        for part in m.group(2).split(","):
            names.append(re.sub(r"\[.*", "", part.strip()))
    assert names == [f[0] for f in _abi.Config._fields_]


def test_flags_and_version_match_header():
    """evaluation flags and the ABI version of the ctypes mirror == the header's #defines"""
    txt = open(os.path.join(ROOT, "include", "cup1d_b200.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(C1D_[A-Z0-9_]+)\s+(\d+)u?\b", txt)}
    assert defs["C1D_ABI_VERSION"] == _abi.ABI_VERSION
    for name in ("ZMASK", "NO_HULL", "NO_CUBE", "EMU_TC", "EMU_I8", "CHI2_I8"):
        assert defs["C1D_FLAG_" + name] == getattr(_abi, "FLAG_" + name), name
    flags = [v for k, v in defs.items() if k.startswith("C1D_FLAG_")]
    assert len(set(flags)) == len(flags) and all(v & (v - 1) == 0 for v in flags)  # distinct single bits
