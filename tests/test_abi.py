"""The C-ABI library loads on a CPU-only box and exports every symbol include/gpplus_b200.h declares;
compute entry points fail loudly without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gpplus_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpp_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    from gpplus_b200 import _engine as E
    assert sorted(E.EXPORTS) == _declared_symbols()


def test_library_exports_every_declared_symbol():
    from gpplus_b200 import _engine as E
    lib = E.load_library()
    for sym in _declared_symbols():
        assert hasattr(lib, sym), sym
    assert lib.gpp_version() >= 100


def test_header_constants_match_binding():
    from gpplus_b200 import _engine as E
    text = open(os.path.join(ROOT, "include", "gpplus_b200.h")).read()
    consts = dict(re.findall(r"#define\s+(GPP_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", text))
    assert int(consts["GPP_OK"]) == E.GPP_OK and int(consts["GPP_ERR_NOT_PD"]) == E.GPP_ERR_NOT_PD
    assert int(consts["GPP_ERR_NAN"]) == E.GPP_ERR_NAN and int(consts["GPP_ERR_CUDA"]) == E.GPP_ERR_CUDA
    assert int(consts["GPP_KERNEL_MATERN52"]) == E.KERNEL_MATERN52 and int(consts["GPP_ACQ_EI"]) == E.ACQ_EI
    assert int(consts["GPP_MAX_DQ"]) == E.MAX_DQ and int(consts["GPP_MAX_DZ"]) == E.MAX_DZ


def test_no_cpu_fallback_without_gpu():
    import torch
    from gpplus_b200 import _engine as E
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert E.device_count() == 0
    with pytest.raises(E.EngineError):
        E.Engine(xq=np.zeros((4, 2)), y=np.arange(4.0), kernel=0)


def test_argument_validation_is_reported_not_crashed():
    from gpplus_b200 import _engine as E
    with pytest.raises((ValueError, E.EngineError)):
        E.Engine(xq=np.zeros((4, 40)), y=np.arange(4.0), kernel=0)  # dq > GPP_MAX_DQ


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gp-plus_b200", "gpplus_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(base, f)
