"""The drop-in boundary: libapj_b200.so must load and export every symbol include/apj_b200.h
declares, the ctypes stub (the binding INTEGRATION.md shows) must cover all of them, and the library
must FAIL LOUDLY without a CUDA device -- there is no CPU fallback (no compute calls here)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "apj_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(apj_[a-z0-9_]+)\s*\(", src)))


def test_header_compiles_as_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "apj_b200.h"\nint main(void){ apj_config c; apj_state s; (void)c; (void)s; return APJ_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c), "-o",
                    str(tmp_path / "t.o")], check=True)


def test_library_exports_every_declared_symbol():
    from active_particle_jamming_b200 import load_library
    from active_particle_jamming_b200.device import EXPORTS
    lib = load_library()
    decl = declared_functions()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(lib, name), "libapj_b200.so does not export " + name
    assert sorted(EXPORTS) == decl, "ctypes stub and header disagree"
    assert b"sm_100a" in lib.apj_version()


def test_config_struct_layout_matches_header():
    """ctypes mirror of apj_config must have the C layout (checked against gcc's sizeof/offsetof)."""
    from active_particle_jamming_b200.device import _Config, _State
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "apj_b200.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(apj_config), offsetof(apj_config, seed), offsetof(apj_config, flags),
  offsetof(apj_config, lanes_per_particle), sizeof(apj_state), offsetof(apj_state, box)); return 0; }'''
    import tempfile
    d = tempfile.mkdtemp()
    open(os.path.join(d, "l.c"), "w").write(prog)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "l.c"), "-o", os.path.join(d, "l")], check=True)
    got = list(map(int, subprocess.run([os.path.join(d, "l")], capture_output=True, text=True).stdout.split()))
    assert got == [C.sizeof(_Config), _Config.seed.offset, _Config.flags.offset, _Config.lanes_per_particle.offset,
                   C.sizeof(_State), _State.box.offset]


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_cuda_available(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback_create_fails_loudly():
    from active_particle_jamming_b200 import ApjError, DeviceEngine
    with pytest.raises(ApjError) as ei:
        DeviceEngine(1024, 60.0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_invalid_arguments_are_rejected_without_a_device():
    from active_particle_jamming_b200 import load_library
    lib = load_library()
    h = C.c_void_p()
    assert lib.apj_create(None, None, C.byref(h)) == -1                    # APJ_E_INVALID
    assert lib.apj_step(None, 1) == -1
    assert lib.apj_destroy(None) == 0
    assert b"null" in lib.apj_last_error(None)


def test_product_package_never_touches_the_oracle():
    """A product path that routes through oracle/ (or any CPU stand-in) voids parity claims."""
    pkg = os.path.join(ROOT, "active_particle_jamming_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "pyoracle" not in src and "apj_oracle" not in src and "libapj_ref" not in src, os.path.join(dp, f)
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)
