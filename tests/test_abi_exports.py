"""The C-ABI shared library loads on a CPU-only box and exports every entry point include/minorg_b200.h declares
(no compute calls here).  Also: the Python product path refuses to work without the library (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "minorg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from minorg_b200 import _lib, build
    lib_path = build.build()
    h = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 24
    missing = [n for n in names if not hasattr(h, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == names            # the ctypes binding covers the whole header
    assert h.mg_version() >= 100


def test_error_reporting_without_gpu():
    from minorg_b200 import _lib
    L = _lib.lib()
    assert L.mg_device_count() >= 0
    rc = L.mg_queries_create(None, 3, 0, None, ctypes.byref(ctypes.c_void_p()))      # invalid length -> MG_ERR_ARG
    assert rc == -2 and b"invalid argument" in L.mg_last_error()
    with pytest.raises(_lib.MinorgB200Error):
        _lib.check(rc, "mg_queries_create")


def test_no_cpu_fallback_when_library_is_missing(monkeypatch, tmp_path):
    from minorg_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MinorgB200Error, match="no CPU fallback"):
        _lib.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "minorg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "screen_oracle" not in src, f


def test_shipped_library_matches_the_sources():
    """build() rebuilds by CONTENT (a SHA-256 of source + headers + flags next to every object): after it, the library's recorded
    digest is the digest of the tree, and a second call compiles nothing."""
    import time
    from minorg_b200 import build
    lib_path = build.build()
    d = build.source_digest()
    assert d is not None and len(d) == 64
    before = os.path.getmtime(lib_path)
    t0 = time.time()
    build.build()
    assert os.path.getmtime(lib_path) == before and time.time() - t0 < 5.0
    assert build.source_digest() == d
