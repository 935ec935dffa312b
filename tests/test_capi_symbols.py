"""The C-ABI library loads on a CPU-only box and exports every symbol declared in include/*.h; compute entry
points fail loudly (no CPU fallback). No GPU compute is attempted here."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        names += re.findall(r"(?:GPB200_EXPORT|GPBDEV_EXPORT)\s+[\w\s\*]+?\b(\w+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(product_lib):
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(product_lib, n), "missing export: " + n


def test_reference_api_names_present(product_lib):
    for n in ("GPB_CreateREModel", "GPB_REModelFree", "GPB_SetOptimConfig", "GPB_OptimCovPar", "GPB_EvalNegLogLikelihood",
              "GPB_GetCovPar", "GPB_GetInitCovPar", "GPB_GetNumIt", "GPB_GetCurrentNegLogLikelihood", "LGBM_GetLastError"):
        assert hasattr(product_lib, n)


def test_no_cpu_fallback_when_no_device(product_lib):
    if product_lib.gpbdev_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    from gpboost_b200 import GPModel, GPBoostError
    with pytest.raises(GPBoostError, match="no CPU fallback"):
        GPModel(gp_coords=np.random.rand(20, 2), gp_approx="vecchia", num_neighbors=5)


def test_unsupported_configurations_error_out(product_lib):
    from gpboost_b200 import GPModel, GPBoostError
    c = np.random.rand(20, 2)
    with pytest.raises(GPBoostError):
        GPModel(gp_coords=c, gp_approx="fitc")
    with pytest.raises(GPBoostError):
        GPModel(gp_coords=c, gp_approx="vecchia", cov_function="matern", cov_fct_shape=0.8)
    with pytest.raises(GPBoostError):
        GPModel(gp_coords=c, gp_approx="vecchia", vecchia_ordering="time")
    assert product_lib.LGBM_GetLastError() is not None


def test_product_does_not_link_or_import_oracle():
    """The product path must not route through the oracle: no file under gpboost_b200/ mentions it."""
    for p in glob.glob(os.path.join(ROOT, "gpboost_b200", "**", "*"), recursive=True):
        if os.path.isfile(p) and p.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
            src = open(p, errors="ignore").read()
            assert "import oracle" not in src and "from oracle" not in src and "liborc" not in src and "oracle/" not in src.replace("oracle/_ref/lib_gpboost.so", ""), p
