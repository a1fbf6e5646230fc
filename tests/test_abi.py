"""CPU: the C-ABI library loads and exports every symbol include/dfmir_b200.h declares; the host
wrappers refuse to run without CUDA tensors (no silent fallback)."""
import pytest
import torch


def test_library_exports_every_declared_symbol():
    from dfmir_b200 import _lib
    lib = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), n
    import re
    want = int(re.search(r"#define DFMIR_ABI_VERSION (\d+)", open(_lib.HEADER_PATH).read()).group(1))
    assert lib.dfmir_abi_version() == want


def test_no_cpu_fallback():
    from dfmir_b200 import _lib, layers, losses
    x = torch.zeros(1, 1, 8, 8)
    f = torch.zeros(1, 2, 8, 8)
    with pytest.raises(_lib.DfmirError):
        layers.SpatialTransformer((8, 8))(x, f)
    with pytest.raises(_lib.DfmirError):
        losses.NCC_Loss('cpu', kernel_var=[9, 9])(x, x)
    with pytest.raises(_lib.DfmirError):
        losses.smooothing_loss(f)


def test_product_does_not_import_oracle():
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dfmir_b200")
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("(the oracle)", ""), f
