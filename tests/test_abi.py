"""CPU-side checks of the C-ABI boundary: the library builds/loads and exports every symbol include/zero_b200.h
declares; the host-side config surface and plugin registry behave like the reference's (run.py, models/model.py)."""
import ctypes
import os
import re

import pytest

from zero_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "zero_b200.h")).read()
    return sorted(set(re.findall(r"\b(zb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(L.LIB_PATH):
        from zero_b200 import build
        build.build()
    lib = ctypes.CDLL(L.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 20
    for sym in declared:
        assert hasattr(lib, sym), "libzero_b200.so does not export %s" % sym
    assert sorted(L.EXPORTS) == declared
    lib.zb_abi_version.restype = ctypes.c_int
    assert lib.zb_abi_version() == 4


def test_ctypes_struct_sizes_match_header_layout():
    # spot-check: natural C alignment of the mirrored structs (pointers 8, int64 8, int32/float 4)
    assert ctypes.sizeof(L.GemmArgs) == 3 * 8 + 6 * 8 + 4 * 4 + 8 + 8 + 8 + 4 + 4
    assert ctypes.sizeof(L.AdamArgs) == 5 * 8 + 8 + 5 * 4 + 4 + 8 + 8


def test_mirrored_struct_layouts_match_the_library():
    lib = ctypes.CDLL(L.LIB_PATH)
    lib.zb_abi_struct_size.argtypes = [ctypes.c_int32]
    lib.zb_abi_struct_size.restype = ctypes.c_int64
    for which, cls in enumerate(L.STRUCTS):
        assert lib.zb_abi_struct_size(which) == ctypes.sizeof(cls), cls.__name__
    assert lib.zb_abi_struct_size(99) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(L.ZeroB200Error):
        L.load()


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from zero_b200.engine import Engine
    from zero_b200.params import transformer_base
    with pytest.raises(L.ZeroB200Error):
        Engine(transformer_base(), 1000, 1000)


def test_every_mirrored_field_offset_matches_the_header(tmp_path):
    """The binding's ctypes structs against include/zero_b200.h compiled by the C compiler: every field name exists in
    the header's struct and sits at the same offset (the size check above would miss two swapped same-size fields)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    pairs = [("zb_gemm_args", L.GemmArgs), ("zb_attention_args", L.AttentionArgs), ("zb_add_ln_args", L.AddLnArgs),
             ("zb_embed_args", L.EmbedArgs), ("zb_ce_args", L.CeArgs), ("zb_adam_args", L.AdamArgs),
             ("zb_beam_args", L.BeamArgs), ("zb_colsum_args", L.ColsumArgs), ("zb_shard_adam_args", L.ShardAdamArgs),
             ("zb_vocab_ce_args", L.VocabCeArgs), ("zb_vocab_topk_args", L.VocabTopkArgs)]
    assert [c for _, c in pairs] == L.STRUCTS
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "%s"' % os.path.join(ROOT, "include", "zero_b200.h"),
             'int main(void) {']
    for cname, cls in pairs:
        for field, _ in cls._fields_:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, field, cname, field))
        lines.append('  printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
    lines.append('  return 0; }')
    src = tmp_path / "offsets.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "offsets"
    subprocess.run([cc, str(src), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True,
                                                       text=True).stdout.splitlines())
    for cname, cls in pairs:
        for field, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, field)]) == getattr(cls, field).offset, (cname, field)
        assert int(got[cname]) == ctypes.sizeof(cls), cname


def test_bound_argument_lists_match_the_prototypes():
    """Number and kind (pointer / integer / float) of the arguments ctypes passes against the header's prototypes."""
    lib = L.load()
    src = open(os.path.join(ROOT, "include", "zero_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = dict(re.findall(r"\bint\s+(zb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src))
    checked = 0
    for name, args in protos.items():
        fn = getattr(lib, name)
        if fn.argtypes is None:
            continue
        want = [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"]
        assert len(fn.argtypes) == len(want), (name, want)
        for ct, decl in zip(fn.argtypes, want):
            is_ptr = "*" in decl or "zb_stream_t" in decl
            if is_ptr:
                assert ct is ctypes.c_void_p or hasattr(ct, "contents"), (name, decl, ct)
            elif "float" in decl:
                assert ct is ctypes.c_float, (name, decl, ct)
            elif "int64_t" in decl:
                assert ct is ctypes.c_int64, (name, decl, ct)
            else:
                assert ct in (ctypes.c_int32, ctypes.c_uint32, ctypes.c_int), (name, decl, ct)
        checked += 1
    assert checked >= 30
