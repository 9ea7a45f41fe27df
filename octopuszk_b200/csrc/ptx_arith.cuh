// 32-bit carry-chain primitives.
//
// Device code: one PTX instruction per wrapper (add.cc / addc / mad.lo.cc / madc.hi.cc ...).  ptxas pairs a
// mad.lo.cc + madc.hi.cc on the same operands into one IMAD.WIDE.U32(.X), so a chain of 2k wrappers costs k
// integer-pipe issues.  The wrappers are `asm volatile` so the order of the chain (and so the carry flag) is kept.
//
// Plain g++ (tests only, never nvcc): the same wrappers emulate the instructions with a thread-local carry flag.
// That lets tests/host_arith_check.cc run the *identical* limb algorithms on the CPU, where no GPU exists, and
// compare them with the Python oracle.  The emulation is compiled out of every product translation unit.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
// Product build (nvcc): the arithmetic exists as DEVICE code only -- liboctozk.so contains no host implementation of
// any field or curve operation, so nothing can silently fall back to the CPU.
#define OZK_HD __device__ __forceinline__
#define OZK_D __device__ __forceinline__
#else
#define OZK_HD inline
#define OZK_D inline
#endif

namespace ozk {
namespace ptx {

#if defined(__CUDACC__)

OZK_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
OZK_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
OZK_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
OZK_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
OZK_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
OZK_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
OZK_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
OZK_D uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
OZK_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
OZK_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
OZK_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
OZK_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
OZK_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

#else  // host emulation (tests only)

inline uint32_t& cc() { static thread_local uint32_t flag = 0; return flag; }
inline uint32_t add3(uint32_t a, uint32_t b, uint32_t cin, bool set) {
    uint64_t s = (uint64_t)a + b + cin;
    if (set) cc() = (uint32_t)(s >> 32);
    return (uint32_t)s;
}
inline uint32_t sub3(uint32_t a, uint32_t b, uint32_t bin, bool set) {
    uint64_t s = (uint64_t)a - b - bin;
    if (set) cc() = (uint32_t)((s >> 32) & 1);   // borrow
    return (uint32_t)s;
}
inline uint32_t add_cc(uint32_t a, uint32_t b) { return add3(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return add3(a, b, cc(), true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return add3(a, b, cc(), false); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return sub3(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return sub3(a, b, cc(), true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return sub3(a, b, cc(), false); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_lo(a, b), c, 0, true); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_lo(a, b), c, cc(), true); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_hi(a, b), c, 0, true); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_hi(a, b), c, cc(), true); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_hi(a, b), c, cc(), false); }

#endif

}  // namespace ptx
}  // namespace ozk
