"""No-GPU checks of the drop-in boundary: liboctozk.so exports every symbol include/octozk.h declares, the three JNI
shim libraries export the reference's six Java_* symbols (algebra_msm_VariableBaseMSM.h:10-24,
algebra_msm_FixedBaseMSM.h:10-32, algebra_fft_FFTAuxiliary.h:10-16), and compute entry points fail loudly without a
device instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from octopuszk_b200 import build
    return build.build()


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "octozk.h")).read()
    declared = set(re.findall(r"OZK_API\s+[\w\s\*]+?\b(ozk_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = ctypes.CDLL(built)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    # the Python binding covers exactly the same set
    from octopuszk_b200 import lib as binding
    assert set(binding.SIGNATURES) == declared
    binding.load_library()


def test_jni_shims_export_reference_symbols(built):
    from tests.jni_harness import SHIMS, FakeJvm
    jvm = FakeJvm()
    for lib, names in SHIMS.items():
        for n in names:
            assert hasattr(jvm.libs[lib], n), (lib, n)
    assert jvm.env


def test_no_cpu_fallback_without_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from octopuszk_b200 import Context, OzkError
    with pytest.raises(OzkError):
        Context(0)
    # and through the JNI boundary: a RuntimeException, not a result
    from oracle import dizk_oracle as O
    from tests.jni_harness import FakeJvm
    jvm = FakeJvm()
    f = jvm.fn("libAlgebraMSMFixedBaseMSM.so", "Java_algebra_msm_FixedBaseMSM_fieldBatchMSMNativeHelper", [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32])
    r = f(jvm.env, None, jvm.bytes_(O.le32(3) + O.le32(5)), 1, 0)
    assert not r and "RuntimeException" in jvm.exception()
    jvm.clear()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under octopuszk_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "octopuszk_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "dizk_oracle" not in src and "c_oracle" not in src and "from oracle" not in src, os.path.join(dirpath, f)
