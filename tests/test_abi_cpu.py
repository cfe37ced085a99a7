"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/mip360_b200.h declares; the ctypes table matches the header; the product path refuses CPU tensors."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mip360_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mip360_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from mipnerf360_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = header_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_ctypes_table_matches_header(lib_path):
    from mipnerf360_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_functions()
    # argument counts: every comma-separated parameter in the header has a ctypes entry
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, args in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), (name, n, len(args))
    lib = _lib.load()
    assert lib.mip360_version() == 100
    assert lib.mip360_partials_len(16) >= 1024


def test_header_is_plain_c_and_a_c_program_links(lib_path, tmp_path):
    """The boundary is a C ABI: the header compiles as C99 (no C++-isms, no torch types) and a C translation unit
    that includes it links against the library and calls it (no device needed for these calls)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Wextra", "-Werror", HEADER], check=True)
    src = tmp_path / "abi_demo.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "mip360_b200.h"\n'
        "int main(void) {\n"
        "  struct mip360_layer l; memset(&l, 0, sizeof l);\n"
        "  if (mip360_version() != 100) return 1;\n"
        "  /* argument errors come back as codes + message, never as exceptions or aborts */\n"
        "  if (mip360_adamw(0, 0, 0, 0, 0, 1e-3f, 0.9f, 0.999f, 1e-8f, 0.f, 1, 0) != MIP360_ERR_ARG) return 2;\n"
        "  if (strstr(mip360_last_error(), \"adamw\") == 0) return 3;\n"
        "  if (mip360_set_option(99, 1) != MIP360_ERR_ARG) return 4;\n"
        '  printf("ok %d %d\\n", mip360_partials_len(1), (int)sizeof l);\n'
        "  return 0;\n}\n")
    exe = tmp_path / "abi_demo"
    libdir = os.path.dirname(lib_path)
    subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir,
                    "-lmip360_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok "), (out.returncode, out.stdout, out.stderr)


def test_no_cpu_fallback(lib_path):
    from mipnerf360_b200 import _lib, ops
    with pytest.raises(_lib.Mip360Error):
        ops.viewdir_enc(torch.randn(4, 3))
    with pytest.raises(_lib.Mip360Error):
        ops.composite(torch.rand(2, 4, 3), torch.rand(2, 4, 1), torch.rand(2, 5), torch.rand(2, 3), False)


def test_argument_errors_are_reported_not_thrown(lib_path):
    from mipnerf360_b200 import _lib
    lib = _lib.load()
    # N out of range is rejected before any launch (no device needed)
    rc = lib.mip360_resample_cdf(ctypes.c_void_p(8), 1, 4096, ctypes.c_void_p(8), None)
    assert rc == -1 and b"N=4096" in lib.mip360_last_error()
    rc = lib.mip360_linear_fwd(None, None, None, 1, 1, 1, 0, None, None, 0, None)
    assert rc == -1


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "mipnerf360_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# oracle", ""), os.path.join(dirpath, f)
