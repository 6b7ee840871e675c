"""CPU-side checks of the product boundary: the C-ABI library builds, loads and exports every symbol
include/pk_collide.h declares, and it refuses to run without a GPU (no CPU fallback)."""
import os
import re

import pytest

import physkit_b200 as pk
from physkit_b200 import build as pk_build
from physkit_b200.api import EXPORTS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    pk_build.build()
    L = pk.load_library()
    hdr = open(os.path.join(ROOT, "include", "pk_collide.h")).read()
    declared = set(re.findall(r"^(?:int|const char \*)\s*\*?(pk_[a-z0-9_]+)\(", hdr, flags=re.M))
    assert len(declared) >= 30
    assert declared == set(EXPORTS), declared ^ set(EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.pk_abi_version() == 1


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pk.PkError) as e:
        pk.Context(16, 16)
    assert e.value.status == -2  # PK_E_NO_DEVICE


def test_product_does_not_reference_oracle():
    """Nothing under physkit_b200/ or include/ may import, link or mention the oracle."""
    for base in ("physkit_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in txt and "from oracle" not in txt and "libpk_oracle" not in txt, f
                    assert "pk_oracle.hpp\"" not in txt or "#include" not in txt.split("pk_oracle.hpp\"")[0][-12:], f


def test_sass_has_no_tensor_or_foreign_paths():
    """The path is FP64 scalar work: no tensor-core mnemonics are expected in the SASS."""
    import shutil
    import subprocess

    if shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump missing")
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([exe, "-sass", pk.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "DADD" in sass and "DMUL" in sass
    assert "HMMA" not in sass and "UTCHMMA" not in sass
