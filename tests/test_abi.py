"""The C-ABI library loads and exports every symbol include/spada_b200.h declares (no GPU needed),
and without a device the product path fails loudly instead of falling back to CPU code."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "spada_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spada_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(spada):
    declared = header_symbols()
    assert len(declared) >= 25
    assert sorted(spada._abi.SYMBOLS) == declared


def test_library_exports_every_declared_symbol(spada):
    lib = C.CDLL(spada._abi.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} declared in include/spada_b200.h but not exported"
    assert spada._abi.lib().spada_b200_abi_version() == spada._abi.ABI_VERSION == 2


def test_struct_layout_matches_header(spada):
    # sizes the C compiler gives the ABI structs (checked against ctypes mirrors)
    import subprocess, tempfile, textwrap
    src = textwrap.dedent("""
        #include <stdio.h>
        #include "spada_b200.h"
        int main(void) { printf("%zu %zu %zu %zu %zu\\n", sizeof(spada_csr_view), sizeof(spada_csr_view32),
                                sizeof(spada_b200_opts), sizeof(spada_b200_launch), sizeof(spada_b200_stats)); return 0; }
    """)
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o",
                               os.path.join(d, "t")])
        sizes = list(map(int, subprocess.check_output([os.path.join(d, "t")]).split()))
    a = spada._abi
    assert sizes == [C.sizeof(a.CsrView), C.sizeof(a.CsrView32), C.sizeof(a.Opts), C.sizeof(a.Launch), C.sizeof(a.Stats)]


def test_no_device_fails_loudly(spada):
    if spada.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(spada.SpadaB200Error) as e:
        spada.Engine()
    assert e.value.status == "NO_DEVICE"
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    # the oracle is test infrastructure: nothing under the product package may reference it
    pkg = os.path.join(ROOT, "spada-sim_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ".rs")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "liboracle" not in text and "import oracle" not in text and "oracle_spgemm" not in text, f


def test_sass_is_sm100a_without_fma(spada):
    """The library is built for sm_100a only, and no kernel of the product path contracts a multiply and an add into
    an FMA (simulator.rs:101 then :217: one rounded multiply, separate rounded adds)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([cuobjdump, "-lelf", spada._abi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", elf))
    assert archs == {"100a"}, archs
    sass = subprocess.run([cuobjdump, "-sass", spada._abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "DMUL" in sass and "DADD" in sass
    assert "DFMA" not in sass
