"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol that
include/cake_b200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest
import torch

from cake_b200 import capi
from cake_b200.build import build
from cake_b200.config import CConfig, llama3_8b

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_loads():
    build()
    L = capi.lib()
    assert L.cake_b200_version().startswith(b"cake_b200")


def test_every_declared_symbol_is_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "cake_b200.h")).read()
    declared = set(re.findall(r"\b(cake_b200_[a-z0-9_]+)\s*\(", hdr))
    bound = {n for n, _, _ in capi.SYMBOLS}
    assert declared == bound, f"header/binding mismatch: {declared ^ bound}"
    L = capi.lib()
    for name in declared:
        assert hasattr(L, name)


def _prototypes(src: str, pattern: str) -> dict:
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(pattern, src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_rust_binding_declares_every_entry_point_with_the_same_arity():
    """rust/ffi.rs cannot be compiled here (no rustc): at least its extern block must list exactly the header's
    functions with the header's argument counts, and its config struct the header's 24 fields in order."""
    hdr = open(os.path.join(ROOT, "include", "cake_b200.h")).read()
    rs = open(os.path.join(ROOT, "rust", "ffi.rs")).read()
    c = _prototypes(hdr, r"\b(cake_b200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;")
    r = _prototypes(rs, r"pub fn (cake_b200_[a-z0-9_]+)\s*\(([^;]*?)\)\s*(?:->\s*[^;]+)?;")
    assert set(c) == set(r), set(c) ^ set(r)
    assert {k: (c[k], r[k]) for k in c if c[k] != r[k]} == {}
    body = re.search(r"pub struct cake_b200_config \{(.*?)\n\}", rs, flags=re.S).group(1)
    rust_fields = re.findall(r"pub (\w+):", body)
    assert rust_fields == [n for n, _ in CConfig._fields_]
    rust_types = dict(re.findall(r"pub (\w+): (\w+)", body))
    for name, ct in CConfig._fields_:
        assert rust_types[name] == ("f32" if ct is ctypes.c_float else "c_int"), name


def test_cconfig_matches_header_layout():
    # 24 4-byte fields, no padding
    assert ctypes.sizeof(CConfig) == 24 * 4
    c = CConfig.from_config(llama3_8b(), "bf16")
    assert (c.hidden, c.inter, c.n_heads, c.n_kv_heads, c.head_dim, c.n_layers, c.vocab) == \
           (4096, 14336, 32, 8, 128, 32, 128256)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    c = CConfig.from_config(llama3_8b(), "bf16")
    h = ctypes.c_void_p()
    rc = capi.lib().cake_b200_ctx_create(0, ctypes.byref(c), ctypes.byref(h))
    assert rc != 0 and capi.lib().cake_b200_last_error()
    from cake_b200.model import Context
    with pytest.raises(RuntimeError):
        Context(llama3_8b(), {}, "bf16")


def test_product_never_imports_the_oracle():
    # the oracle is the checker, never the product path: nothing under cake_b200/ may import, link or call it
    pkg = os.path.join(ROOT, "cake_b200")
    bad = re.compile(r"^\s*(import|from)\s+oracle\b|cake_oracle|libcake_oracle|\bora_[a-z_]+\s*\(", re.M)
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                assert not bad.search(src), f"{f} reaches into the oracle"


def test_null_handles_are_errors_not_crashes():
    """`errors never abort`: every entry point called with null handles / zero sizes returns a
    status (or a null/negative value) and leaves a message — checked in a child process so that a regression shows up as
    a failed assertion, not a dead test runner."""
    import subprocess
    import sys
    code = r'''
import ctypes, sys
sys.path.insert(0, %r)
from cake_b200 import capi
L = capi.lib()
bad = []
for name, restype, argtypes in capi.SYMBOLS:
    args = [0.0 if t is ctypes.c_float else 0 if t in (ctypes.c_int, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_uint64) else None
            for t in argtypes]
    rc = getattr(L, name)(*args)
    if name in ("cake_b200_version", "cake_b200_last_error", "cake_b200_ctx_destroy", "cake_b200_block_free", "cake_b200_cache_free"):
        continue
    if name == "cake_b200_stream":
        ok = rc in (None, 0)
    elif name in ("cake_b200_block_layer", "cake_b200_cache_len"):
        ok = rc < 0
    else:
        ok = rc != 0 and bool(L.cake_b200_last_error())
    if not ok:
        bad.append((name, rc))
print("BAD", bad)
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, f"crashed with {r.returncode}: {r.stderr[-500:]}"
    assert "BAD []" in r.stdout, r.stdout
