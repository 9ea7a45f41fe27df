"""Python side of the fake-JNIEnv harness (tests/fake_jni.cc): loads the three shim libraries the reference's Java
would System.loadLibrary and calls their Java_* symbols with fake byte[] / List / ByteBuffer objects."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "octopuszk_b200", "lib")
vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int32

SHIMS = {
    "libAlgebraMSMVariableBaseMSM.so": ["Java_algebra_msm_VariableBaseMSM_variableBaseSerialMSMNativeHelper",
                                        "Java_algebra_msm_VariableBaseMSM_variableBaseDoubleMSMNativeHelper",
                                        "Java_algebra_msm_VariableBaseMSM_variableBaseMSMDirect",
                                        "Java_algebra_msm_VariableBaseMSM_uploadBasesDirect",
                                        "Java_algebra_msm_VariableBaseMSM_freeBases",
                                        "Java_algebra_msm_VariableBaseMSM_variableBaseMSMKeyedDirect"],
    "libAlgebraMSMFixedBaseMSM.so": ["Java_algebra_msm_FixedBaseMSM_batchMSMNativeHelper",
                                     "Java_algebra_msm_FixedBaseMSM_doubleBatchMSMNativeHelper",
                                     "Java_algebra_msm_FixedBaseMSM_fieldBatchMSMNativeHelper",
                                     "Java_algebra_msm_FixedBaseMSM_batchMSMDirect"],
    "libAlgebraFFTAuxiliary.so": ["Java_algebra_fft_FFTAuxiliary_serialRadix2FFTNativeHelper",
                                  "Java_algebra_fft_FFTAuxiliary_serialRadix2FFTDirect"],
}


class FakeJvm:
    def __init__(self):
        os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
        so = os.path.join(ROOT, "build", "libfake_jni.so")
        src = os.path.join(ROOT, "tests", "fake_jni.cc")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
        self.fj = ctypes.CDLL(so)
        self.fj.fj_env.restype = vp
        self.fj.fj_new_bytes.restype = vp
        self.fj.fj_new_bytes.argtypes = [ctypes.c_char_p, sz]
        self.fj.fj_new_direct.restype = vp
        self.fj.fj_new_direct.argtypes = [ctypes.c_char_p, sz]
        self.fj.fj_new_list.restype = vp
        self.fj.fj_list_add.argtypes = [vp, vp]
        self.fj.fj_len.restype = sz
        self.fj.fj_len.argtypes = [vp]
        self.fj.fj_data.restype = vp
        self.fj.fj_data.argtypes = [vp]
        self.fj.fj_free.argtypes = [vp]
        self.fj.fj_exception.restype = ctypes.c_char_p
        self.env = self.fj.fj_env()
        self.libs = {name: ctypes.CDLL(os.path.join(LIBDIR, name)) for name in SHIMS}

    def fn(self, lib, name, argtypes, restype=vp):
        f = getattr(self.libs[lib], name)
        f.argtypes = [vp, vp] + argtypes
        f.restype = restype
        return f

    def bytes_(self, b):
        return self.fj.fj_new_bytes(b, len(b))

    def direct(self, b=None, size=None):
        return self.fj.fj_new_direct(b, len(b) if b is not None else size)

    def list_(self, items):
        lst = self.fj.fj_new_list()
        for it in items:
            self.fj.fj_list_add(lst, self.bytes_(it))
        return lst

    def read(self, handle):
        n = self.fj.fj_len(handle)
        return ctypes.string_at(self.fj.fj_data(handle), n)

    def exception(self):
        e = self.fj.fj_exception()
        return e.decode() if e else None

    def clear(self):
        self.fj.fj_clear_exception()

    def live_pins(self):
        return self.fj.fj_live_pins()

    def critical_violations(self):
        """JNI calls made by a thread while it held a GetPrimitiveArrayCritical pin (forbidden by the JNI specification)."""
        return self.fj.fj_critical_violations()

    def pin_calls(self):
        return self.fj.fj_pin_calls()

    def fail_next_pins(self, k):
        self.fj.fj_fail_next_pins(k)
