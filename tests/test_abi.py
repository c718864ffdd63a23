"""C-ABI library: loads without a GPU, exports every symbol include/b200lp.h declares, and the ctypes binding
declares exactly those symbols (no compute calls here)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    text = (ROOT / "include" / "b200lp.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200lp_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_header_symbols():
    from b200lp import lib
    handle = lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(handle, s), f"libb200lp.so does not export {s}"
    assert handle.b200lp_abi_version() == lib.ABI_VERSION
    m = re.search(r"#define\s+B200LP_ABI_VERSION\s+(\d+)", (ROOT / "include" / "b200lp.h").read_text())
    assert int(m.group(1)) == lib.ABI_VERSION


def test_binding_covers_header_exactly():
    from b200lp import lib
    assert sorted(lib.SIGNATURES.keys()) == header_symbols()


def test_struct_layouts_match_header():
    from b200lp import lib
    # b200lp_conv_args: 7 pointers + 15 int32 (+pad) + workspace pointer + int64 ; b200lp_wgrad_args: 4 pointers + int64 + 6 int32 + float
    assert ctypes.sizeof(lib.ConvArgs) == 7 * 8 + 18 * 4 + 2 * 8
    assert ctypes.sizeof(lib.WgradArgs) == 4 * 8 + 8 + 6 * 4 + 4 + 4 * 4 + 4   # + tail padding to 8
    assert lib.ConvArgs.block_n.offset == 7 * 8 + 9 * 4 and lib.ConvArgs.precision.offset == 7 * 8 + 10 * 4
    assert lib.WgradArgs.scale.offset == 4 * 8 + 8 + 6 * 4


def test_sass_contains_blackwell_tensor_core_and_tma_instructions():
    """The shipped binary really is tcgen05 + TMA code (UTC*MMA / UTMALDG / LDTM), not a legacy mma.sync path."""
    import shutil
    import subprocess
    from b200lp import lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        import pytest
        pytest.skip("cuobjdump not available")
    lib.load()
    sass = subprocess.run([cuobjdump, "-sass", str(lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass or re.search(r"UTC\w*MMA", sass)
    assert "UTMALDG" in sass
    assert "LDTM" in sass
    assert "HMMA.16816" not in sass and "HGMMA" not in sass


def test_integration_md_stub_matches_the_binding():
    """The ctypes stub INTEGRATION.md shows a reference maintainer must describe the same struct as the binding this
    repo uses (field names, order, types) and build its ConvArgs with one positional value per field."""
    import ctypes
    import re
    from conftest import ROOT
    from b200lp import lib as L
    text = (ROOT / "INTEGRATION.md").read_text()
    block = text.split("class ConvArgs(Structure):", 1)[1].split("_lib.b200lp_conv_fwd.argtypes", 1)[0]
    fields = [(n, getattr(ctypes, t)) for n, t in re.findall(r'\("(\w+)",\s*(c_\w+)\)', block)]   # c_int32 is c_int
    want = list(L.ConvArgs._fields_)
    assert fields == want, (fields, want)
    call = text.split("a = ConvArgs(", 1)[1].split(")  #", 1)[0]
    depth, n_args = 0, 1
    for ch in call:
        depth += ch == "("
        depth -= ch == ")"
        n_args += ch == "," and depth == 0
    assert n_args == len(want), (n_args, len(want))
    assert ctypes.sizeof(L.ConvArgs) % 8 == 0


def test_struct_field_offsets_match_the_c_compiler():
    """include/b200lp.h compiled by gcc as plain C: sizeof and every field's offsetof must equal the ctypes structures
    of the binding (the header is the contract a C / cgo / JNI caller would compile against)."""
    import shutil
    import subprocess
    import tempfile
    import pytest
    from b200lp import lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    pairs = {"b200lp_conv_args": lib.ConvArgs, "b200lp_wgrad_args": lib.WgradArgs, "b200lp_sn_item": lib.SnItem}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200lp.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        src.append(f'  printf("{cname} sizeof %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            src.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    src += ['  return 0;', '}']
    with tempfile.TemporaryDirectory() as d:
        c = Path(d) / "abi.c"
        c.write_text("\n".join(src))
        exe = Path(d) / "abi"
        r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(c), "-o", str(exe)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr          # the header is valid C99 and names every field the binding names
        out = subprocess.run([str(exe)], capture_output=True, text=True).stdout
    got = {}
    for line in out.strip().splitlines():
        cname, fname, val = line.split()
        got[(cname, fname)] = int(val)
    for cname, cls in pairs.items():
        assert got[(cname, "sizeof")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)


def test_every_prototype_matches_the_ctypes_signature():
    """Argument by argument: each prototype of include/b200lp.h against the binding's (restype, argtypes) — pointer,
    int32, int64, float or double in the same position, same count (a silent mismatch would shift every later argument)."""
    from b200lp import lib
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "b200lp.h").read_text(), flags=re.S)
    protos = re.findall(r"([\w][\w\s\*]*?)\s*\b(b200lp_\w+)\s*\(([^)]*)\)\s*;", text)
    assert len(protos) == len(lib.SIGNATURES) >= 100

    def c_kind(t):
        t = t.strip()
        if t == "void":
            return None
        for key, kind in (("*", "P"), ("int32_t", "I"), ("int64_t", "L"), ("float", "F"), ("double", "D")):
            if key in t:
                return kind
        raise AssertionError(f"unclassified C type {t!r}")

    def ct_kind(t):
        if t is None:
            return None
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents"):
            return "P"
        return {ctypes.c_int32: "I", ctypes.c_int64: "L", ctypes.c_float: "F", ctypes.c_double: "D"}[t]

    for ret, name, params in protos:
        want = [k for k in (c_kind(p) for p in params.split(",")) if k is not None]
        restype, argtypes = lib.SIGNATURES[name]
        assert [ct_kind(a) for a in argtypes] == want, name
        assert ct_kind(restype) == c_kind(ret), name
