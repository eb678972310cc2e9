"""CPU-side checks of the drop-in boundary: the library builds, loads, and exports every
symbol include/oss_b200.h declares; without a GPU it fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

import openskystacker_b200 as ob
from openskystacker_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so():
    return ob.build()


def test_header_symbols_are_exported(so):
    header = open(os.path.join(ROOT, "include", "oss_b200.h")).read()
    declared = set(re.findall(r"OSS_API\s+[\w\s\*]+?\b(oss_\w+)\s*\(", header))
    assert len(declared) >= 35
    L = ctypes.CDLL(so)
    missing = [name for name in sorted(declared) if not hasattr(L, name)]
    assert not missing, missing
    assert declared == set(binding.EXPORTS), declared ^ set(binding.EXPORTS)


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/oss_b200.h must compile as C99 (what cgo / a C caller / a bindgen run would see), with
    plain pointers and sizes only — no C++ or torch types in any signature."""
    import shutil
    import subprocess
    cc = shutil.which("gcc")
    if not cc:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include "oss_b200.h"\nint main(void) { oss_ctx *c = 0; oss_star s; oss_frame_report r; (void)c; (void)s; (void)r; return oss_device_count() < 0; }\n')
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    code = re.sub(r"/\*.*?\*/|//[^\n]*", "", open(os.path.join(ROOT, "include", "oss_b200.h")).read(), flags=re.S)   # comments aside
    assert "std::" not in code and "torch" not in code.lower() and "at::" not in code and "class " not in code


def test_error_strings_match_the_reference():
    L = ob.lib()
    assert L.oss_status_string(-3) == b"Images must be the same type."                      # imagestacker.cpp:243
    assert L.oss_status_string(-4) == b"Images must all be the same size."                  # imagestacker.cpp:256
    assert L.oss_status_string(-5) == b"Number of threads must be at least 1."              # imagestacker.cpp:290
    assert L.oss_status_string(-6).startswith(b"No images could be aligned to the reference image.")  # :357


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ob.OssError) as e:
        ob.Engine()
    assert e.value.status == -1


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "openskystacker_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "oss_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f
