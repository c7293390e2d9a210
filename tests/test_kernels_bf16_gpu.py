"""The per-kernel numerics tests of test_kernels_gpu.py once more on the bf16 library build
(libinstantir_b200.so): the same test functions, executed with H16 = torch.bfloat16."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu

_HERE = os.path.dirname(os.path.abspath(__file__))
_prev = os.environ.get("IIR_TEST_H16")
os.environ["IIR_TEST_H16"] = "bf16"
try:
    _spec = importlib.util.spec_from_file_location("_kernels_bf16", os.path.join(_HERE, "test_kernels_gpu.py"))
    _mod = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_mod)
finally:
    if _prev is None:
        os.environ.pop("IIR_TEST_H16", None)
    else:
        os.environ["IIR_TEST_H16"] = _prev
# dtype-independent tests (fp32 SIMT variants are parametrised inside the shared functions and simply run twice)
globals().update({k: v for k, v in vars(_mod).items() if k.startswith("test_")})
