"""octopuszk_b200 -- B200-native (sm_100a) Groth16 arithmetic hot path (variable-base MSM, fixed-base batch MSM,
radix-2 NTT over BN254a) behind the C ABI of include/octozk.h, plus a Python mirror of the reference's Java
operator interface (algebra.msm.VariableBaseMSM / FixedBaseMSM, algebra.fft.SerialFFT)."""
from .lib import Context, OzkError, load_library, library_path  # noqa: F401

__all__ = ["Context", "OzkError", "load_library", "library_path"]
