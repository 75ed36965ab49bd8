"""The C-ABI library loads and exports exactly what include/pixo_b200.h declares.  CPU only:
no compute entry point is called with a device."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "pixo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pixo_b200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in pixo_b200.h but not exported"


def test_binding_table_matches_header(lib):
    from pixo_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _header_functions()


def test_version_and_host_only_entry_points(lib):
    assert lib.pixo_b200_version() == 0x000100
    import numpy as np
    from pixo_b200 import jpeg
    lz, cz, ln, cn = jpeg.quant_tables(50)
    assert ln[0] == 16 and lz[0] == 16
    assert jpeg.block_counts(3840, 2160) == (129600, 32400)
    assert jpeg.block_counts(1920, 1080) == (4 * 120 * 68, 120 * 68)
    assert jpeg.block_counts(9, 9, 0, 0) == (4, 0)


def test_no_cpu_fallback_without_device(lib):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import pixo_b200
    if lib.pixo_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pixo_b200.PixoError) as e:
        pixo_b200.Context(0)
    assert e.value.code == pixo_b200._lib.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_link_or_import_the_oracle():
    """oracle/ is test infrastructure: the package sources and the built library never
    reference it."""
    import subprocess
    pkg = os.path.join(ROOT, "pixo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pixo_oracle" not in txt and "pyoracle" not in txt, f
    ldd = subprocess.run(["ldd", os.path.join(pkg, "libpixo_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_plain_c_client_compiles_links_and_runs(lib, tmp_path):
    """include/pixo_b200.h is a C header (strict C11, -Wall -Wextra -Werror -pedantic) and the
    library links from C: tests/c/abi_client.c exercises the host-only entry points and the loud
    no-device failure exactly as a cgo / Rust FFI / JNI binding would see them."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no C compiler")
    exe = str(tmp_path / "abi_client")
    pkg = os.path.join(ROOT, "pixo_b200")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_client.c"), "-o", exe, "-L", pkg, "-lpixo_b200",
                    f"-Wl,-rpath,{pkg}"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "abi_client ok" in out.stdout, (out.returncode, out.stdout, out.stderr)
