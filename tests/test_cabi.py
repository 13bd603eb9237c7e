"""CPU: the C-ABI library loads and exports exactly the entry points include/mdvit_b200.h declares (no compute)."""
import ctypes
import subprocess

from mdvit_b200 import _lib


def test_header_parses():
    sigs = _lib.header_signatures()
    assert len(sigs) >= 30
    assert "mdv_gemm_nt" in sigs and "mdv_attn_fwd" in sigs and "mdv_adamw" in sigs


def test_library_exports_every_declared_symbol(built_lib):
    sigs = _lib.header_signatures()
    for name in sigs:
        assert hasattr(built_lib, name), f"{name} declared in include/mdvit_b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln and ln.split()[-1].startswith("mdv_")}
    assert exported == set(sigs), f"header/library mismatch: {exported ^ set(sigs)}"


def test_version_and_arg_checks(built_lib):
    assert built_lib.mdv_version() == 100
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert built_lib.mdv_gemm_nt(None, 0, None, 0, 0, 0, 0, None, None) == -1
    assert built_lib.mdv_layernorm_fwd(None, None, None, 1e-6, None, None, None, 0, 63, None) == -1
    assert built_lib.mdv_adamw(None, None, None, None, None, 0, None) == -1


def test_epilogue_struct_layout_matches_header():
    # 9 pointers + 12 ints/floats/uint32; and the field ORDER of the ctypes mirror is the header's
    assert ctypes.sizeof(_lib.GemmEpi) == 9 * 8 + 12 * 4
    import re
    text = open(_lib.HEADER_PATH).read()
    body = text[text.index("typedef struct MdvGemmEpi {"):text.index("} MdvGemmEpi;")]
    body = re.sub(r"/\*.*?\*/", " ", body, flags=re.S)
    names = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if decl:
            names += [n.strip().lstrip("*") for n in decl.rsplit(" ", 1)[0:0]] or [d.strip().split()[-1].lstrip("*") for d in decl.split(",")]
    assert names == [f[0] for f in _lib.GemmEpi._fields_], (names, [f[0] for f in _lib.GemmEpi._fields_])
