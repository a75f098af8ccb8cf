"""CPU: the C-ABI shared library loads here (no GPU, no libcuda) and exports every symbol include/teco.h declares;
the ctypes table in tecogan_b200/_ffi.py covers exactly the same set with matching arity."""
import ctypes
import os
import re

from tests.conftest import ROOT


def _header_decls():
    src = open(os.path.join(ROOT, "include", "teco.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char\*)\s+(teco_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(1)] = n
    return decls


def test_library_loads_without_a_gpu_and_exports_every_header_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "tecogan_b200", "libteco.so"))
    decls = _header_decls()
    assert len(decls) >= 35
    for name in decls:
        assert hasattr(lib, name), "libteco.so does not export " + name
    lib.teco_version.restype = ctypes.c_int
    assert lib.teco_version() >= 100
    lib.teco_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.teco_last_error(), bytes)


def test_ctypes_table_matches_header():
    from tecogan_b200 import _ffi
    decls = _header_decls()
    table = dict(_ffi.SIGNATURES)
    table["teco_last_error"] = []
    assert set(table) == set(decls), (set(table) ^ set(decls))
    for name, args in table.items():
        assert len(args) == decls[name], (name, len(args), decls[name])


def test_struct_layouts_match_header_field_order():
    from tecogan_b200 import _ffi
    src = open(os.path.join(ROOT, "include", "teco.h")).read()
    for struct, cls in (("teco_conv_desc", _ffi.ConvDesc), ("teco_tc_desc", _ffi.TcDesc)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for line in body.split(";"):
            line = line.strip()
            if not line:
                continue
            typ, rest = line.split(None, 1)
            names += [n.strip() for n in rest.split(",")]
        assert names == [f[0] for f in cls._fields_], struct


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under tecogan_b200/, main.py or runGan.py may import it."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "tecogan_b200")):
        for f in files:
            if f.endswith(".py") and re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(base, f)).read(), flags=re.M):
                bad.append(f)
    for f in ("main.py", "runGan.py"):
        if re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(ROOT, f)).read(), flags=re.M):
            bad.append(f)
    assert not bad, bad


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import pytest
    from tecogan_b200 import _ffi
    monkeypatch.setattr(_ffi, "_lib", None)
    monkeypatch.setattr(_ffi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _ffi.lib()
