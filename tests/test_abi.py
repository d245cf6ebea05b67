"""The C-ABI shared library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import re

from conftest import ROOT


def _declared():
    text = (ROOT / "include" / "kdiffusion_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kdb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from k_diffusion import _native
    names = _declared()
    assert len(names) >= 25
    handle = ctypes.CDLL(str(_native.LIB_PATH))
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_native.SIGNATURES), set(names) ^ set(_native.SIGNATURES)
    assert _native.lib().kdb_abi_version() == _native.ABI_VERSION


def test_error_reporting_without_gpu():
    from k_diffusion import _native
    L = _native.lib()
    rc = L.kdb_solver_lincomb(None, None, 0, None, 0, None)
    assert rc == -1 and b"n_in" in L.kdb_last_error()
    cfg = _native.KdbModelConfig()
    h = ctypes.c_void_p()
    assert L.kdb_model_create(ctypes.byref(cfg), ctypes.byref(h)) == -1      # n_levels = 0
    assert _native.launch_count() == 0 or _native.launch_count() > 0        # callable


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from k_diffusion import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", tmp_path / "nope.so")
    import pytest
    with pytest.raises(_native.NativeLibraryError):
        _native.lib()


def test_every_entry_point_rejects_null_arguments_without_touching_the_gpu():
    """Each compute entry validates its arguments before its first CUDA call: all-NULL / all-zero arguments come back as a
    negative kdb error code with a message (never a crash, never a cudaError from a launch attempt)."""
    from k_diffusion import _native
    L = _native.lib()
    no_args_or_void = {"kdb_abi_version", "kdb_last_error", "kdb_launch_count", "kdb_launch_breakdown", "kdb_model_destroy",
                       "kdb_profile_end", "kdb_profile_gate"}          # gate(0 ns) is a legal request: it launches
    size_queries = {"kdb_model_workspace_bytes": 0, "kdb_model_tap_count": 0}   # return a size, 0 for a NULL model
    checked = 0
    for name, (res, args) in _native.SIGNATURES.items():
        if name in no_args_or_void:
            continue
        zeros = [0.0 if a in (ctypes.c_float, ctypes.c_double) else
                 0 if a in (ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_size_t) else None for a in args]
        rc = getattr(L, name)(*zeros)
        if name in size_queries:
            assert rc == size_queries[name], (name, rc)
        else:
            assert rc < 0, f"{name}(NULL...) returned {rc}"
            assert L.kdb_last_error(), name
        checked += 1
    assert checked >= 24


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """The boundary is a C ABI, not a C++ one: the header compiles as pedantic C99 and a C program that calls into the
    library links against libkdb200.so and runs without a GPU (kdb_abi_version, kdb_last_error, an argument-validation failure)."""
    import shutil
    import subprocess
    import pytest
    from k_diffusion import _native
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "kdiffusion_b200.h"\n'
                   'int main(void) {\n'
                   '  if (kdb_abi_version() != KDB_ABI_VERSION) return 1;\n'
                   '  if (kdb_solver_lincomb(NULL, NULL, 0, NULL, 0, NULL) >= 0) return 2;\n'
                   '  if (strlen(kdb_last_error()) == 0) return 3;\n'
                   '  printf("abi %d\\n", kdb_abi_version());\n'
                   '  return 0;\n}\n')
    inc, libdir = str(ROOT / "include"), str(_native.LIB_PATH.parent)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)], check=True)
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-l:" + _native.LIB_PATH.name,
                    "-Wl,-rpath," + libdir, "-Wl,--unresolved-symbols=ignore-in-shared-libs"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out.strip() == f"abi {_native.ABI_VERSION}"


def test_integration_guide_names_every_entry_point():
    """INTEGRATION.md is the reference-side binding guide: it must mention every function the header declares."""
    doc = (ROOT / "INTEGRATION.md").read_text()
    missing = [n for n in _declared() if n not in doc]
    assert not missing, missing
