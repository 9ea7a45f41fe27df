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


def test_shims_cover_the_reference_javah_headers(built):
    """Every native the reference's javah headers declare (algebra_msm_VariableBaseMSM.h, algebra_msm_FixedBaseMSM.h,
    algebra_fft_FFTAuxiliary.h: name, return type and JNI signature) is exported by the shim library the Java loads for that
    class, with the same parameter list in jni_shim.cc.  Needs the reference tree (the build container has it; skipped on
    the GPU box)."""
    ref = "/root/reference"
    headers = {"algebra_msm_VariableBaseMSM.h": "libAlgebraMSMVariableBaseMSM.so", "algebra_msm_FixedBaseMSM.h": "libAlgebraMSMFixedBaseMSM.so",
               "algebra_fft_FFTAuxiliary.h": "libAlgebraFFTAuxiliary.so"}
    if not all(os.path.exists(os.path.join(ref, h)) for h in headers):
        pytest.skip("reference tree not mounted")
    shim_src = open(os.path.join(ROOT, "octopuszk_b200", "csrc", "jni", "jni_shim.cc")).read()
    libdir = os.path.join(ROOT, "octopuszk_b200", "lib")
    jni_letter = {"jbyteArray": "[B", "jint": "I", "jobject": "L", "jlong": "J"}
    total = 0
    for hdr, lib in headers.items():
        text = open(os.path.join(ref, hdr)).read()
        decls = re.findall(r"Signature:\s*(\S+)\s*\*/\s*JNIEXPORT\s+(\w+)\s+JNICALL\s+(Java_\w+)\s*\(([^)]*)\)", text)
        assert decls, hdr
        so = ctypes.CDLL(os.path.join(libdir, lib))
        for sig, ret, name, params in decls:
            total += 1
            assert hasattr(so, name), (lib, name)
            m = re.search(r"JNIEXPORT\s+(\w+)\s+JNICALL\s+" + name + r"\s*\(([^)]*)\)", shim_src)
            assert m, name
            assert m.group(1) == ret, (name, m.group(1), ret)
            ours = [p.strip().split()[0].rstrip("*") for p in m.group(2).split(",")]
            theirs = [p.strip().split()[0].rstrip("*") for p in params.split(",")]
            assert ours == theirs, (name, ours, theirs)
            # and the declared JNI signature agrees with that parameter list (after JNIEnv*, jclass)
            letters = "".join(jni_letter[t] if t != "jobject" else "Ljava/util/List;" for t in theirs[2:])
            assert sig.startswith("(" + letters + ")"), (name, sig, letters)
    assert total == 6


def test_header_is_plain_c(tmp_path):
    """include/octozk.h is the C ABI: it must compile as C99 (no C++ types, no torch types), and a C caller can link
    against liboctozk.so (link check only; nothing is executed)."""
    import subprocess
    src = tmp_path / "caller.c"
    src.write_text('#include "octozk.h"\n'
                   "int main(void) {\n"
                   "    ozk_ctx* c = 0; uint8_t out[96]; ozk_bases* k = 0;\n"
                   "    if (ozk_device_count() <= 0) return 0;\n"
                   "    if (ozk_ctx_create(0, &c) != OZK_OK) return 1;\n"
                   "    ozk_msm_g1(c, 0, 0, 0, out);\n"
                   "    ozk_bases_free(c, k);\n"
                   "    ozk_ctx_destroy(c);\n"
                   "    return 0;\n}\n")
    inc = os.path.join(ROOT, "include")
    libdir = os.path.join(ROOT, "octopuszk_b200", "lib")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, "-c", str(src), "-o", str(tmp_path / "caller.o")])
    subprocess.check_call(["gcc", str(tmp_path / "caller.o"), "-o", str(tmp_path / "caller"), "-L", libdir, "-loctozk", "-Wl,-rpath," + libdir,
                           "-Wl,--unresolved-symbols=ignore-in-shared-libs"])
