// mtg_device.cuh -- shared device helpers and the exact constant tables.
//
// Tables: H(1;r) = A(1)^-T Q(1;r) A(1)^-1 and A(1)^-1, exact rationals rounded once
// (oracle/gen_tables.py).  The kernels never form A(T)^-1 or Q(T) numerically; they use
//     A(T)^-1 = diag(T^-j) A(1)^-1 diag(T^(s mod h))
//     H(T)    = T^(1-2r) diag(T^(s mod h)) H(1) diag(T^(s mod h))
// which replace the reference's per-segment setupMappingMatrix / invertMappingMatrix /
// computeQuadraticCostJacobian / Ai^T Q Ai (impl/polynomial_optimization_linear_impl.h
// :111-121, :142-179, :567-583, :316-318) and avoid the cancellation of forming
// A^-T Q A^-1 in floating point.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mtg_tables.h"

#define MTG_MAX_N_HALF 6  // Polynomial::kMaxN / 2 (polynomial.h:44)

namespace mtg {

// Tables in constant memory.  The specialised kernels index them with compile-time constants
// (fully unrolled loops), so every entry becomes a constant-bank operand c[3][imm] of the
// consuming DFMA/DMUL -- no load and no register; the generic kernel indexes them at run time
// with warp-uniform indices (broadcast).
#define MTG_DECL_A1INV(N_) __constant__ double c_a1inv_##N_[] = MTG_A1INV_##N_;
#define MTG_DECL_H1(N_, R_) __constant__ double c_h1_##N_##_##R_[] = MTG_H1_##N_##_##R_;
MTG_DECL_A1INV(2)
MTG_DECL_A1INV(4)
MTG_DECL_A1INV(6)
MTG_DECL_A1INV(8)
MTG_DECL_A1INV(10)
MTG_DECL_A1INV(12)
MTG_DECL_H1(2, 0)
MTG_DECL_H1(4, 0)
MTG_DECL_H1(4, 1)
MTG_DECL_H1(6, 0)
MTG_DECL_H1(6, 1)
MTG_DECL_H1(6, 2)
MTG_DECL_H1(8, 0)
MTG_DECL_H1(8, 1)
MTG_DECL_H1(8, 2)
MTG_DECL_H1(8, 3)
MTG_DECL_H1(10, 0)
MTG_DECL_H1(10, 1)
MTG_DECL_H1(10, 2)
MTG_DECL_H1(10, 3)
MTG_DECL_H1(10, 4)
MTG_DECL_H1(12, 0)
MTG_DECL_H1(12, 1)
MTG_DECL_H1(12, 2)
MTG_DECL_H1(12, 3)
MTG_DECL_H1(12, 4)
MTG_DECL_H1(12, 5)

template <int N>
struct A1Inv;
template <int N, int R>
struct H1;

// Two ways to feed a table entry with a compile-time index to an FP64 instruction (sm_100 DFMA/DMUL
// take register or uniform-register operands, not c[bank][offset]):
//   A1Inv / H1        entries are loaded from constant memory (LDC/LDCU) and the compiler keeps the hot
//                     ones in registers across loop iterations: fewest instructions, ~50 more live
//                     registers.  Used by the kernels that are not register-bound.
//   A1InvImm / H1Imm  entries are compile-time immediates (zero live registers).  H1Imm is the INTEGER-
//                     SCALED table H(1;r) * lcm(denominators) (the reduced system's solution is invariant
//                     to scaling H): its entries are exact small integers whose odd part fits 21 bits, so
//                     they are encoded in the 32-bit immediate field of DFMA/DMUL -- no UMOV, no register,
//                     no load.  A(1)^-1 entries are small dyadic rationals and are immediates as they are.
//                     Used by the TMEM kernel, which runs at the register limit (2 CTAs x 128 threads/SM).
template <int N>
struct A1InvImm;
template <int N, int R>
struct H1Imm;
#define MTG_DEF_A1INV(N_)                                                                          \
  template <>                                                                                      \
  struct A1Inv<N_> {                                                                               \
    static __device__ __forceinline__ double at(int r, int c) { return c_a1inv_##N_[r * N_ + c]; } \
  };                                                                                               \
  template <>                                                                                      \
  struct A1InvImm<N_> {                                                                            \
    static __device__ __forceinline__ constexpr double at(int r, int c) {                          \
      constexpr double t[] = MTG_A1INV_##N_;                                                       \
      return t[r * N_ + c];                                                                        \
    }                                                                                              \
  };
#define MTG_DEF_H1(N_, R_)                                                                             \
  template <>                                                                                          \
  struct H1<N_, R_> {                                                                                  \
    static __device__ __forceinline__ double at(int r, int c) { return c_h1_##N_##_##R_[r * N_ + c]; } \
  };                                                                                                   \
  template <>                                                                                          \
  struct H1Imm<N_, R_> { /* integer-scaled table: exact, entries fit the FP64 immediate field */       \
    static __device__ __forceinline__ constexpr double at(int r, int c) {                              \
      constexpr double t[] = MTG_H1S_##N_##_##R_;                                                      \
      return t[r * N_ + c];                                                                            \
    }                                                                                                  \
    static constexpr double scale = MTG_H1S_SCALE_##N_##_##R_; /* table = H(1;r) * scale */            \
  };

MTG_DEF_A1INV(2)
MTG_DEF_A1INV(4)
MTG_DEF_A1INV(6)
MTG_DEF_A1INV(8)
MTG_DEF_A1INV(10)
MTG_DEF_A1INV(12)
MTG_DEF_H1(2, 0)
MTG_DEF_H1(4, 0)
MTG_DEF_H1(4, 1)
MTG_DEF_H1(6, 0)
MTG_DEF_H1(6, 1)
MTG_DEF_H1(6, 2)
MTG_DEF_H1(8, 0)
MTG_DEF_H1(8, 1)
MTG_DEF_H1(8, 2)
MTG_DEF_H1(8, 3)
MTG_DEF_H1(10, 0)
MTG_DEF_H1(10, 1)
MTG_DEF_H1(10, 2)
MTG_DEF_H1(10, 3)
MTG_DEF_H1(10, 4)
MTG_DEF_H1(12, 0)
MTG_DEF_H1(12, 1)
MTG_DEF_H1(12, 2)
MTG_DEF_H1(12, 3)
MTG_DEF_H1(12, 4)
MTG_DEF_H1(12, 5)

// Status bits (mirror include/mtg_b200.h).
constexpr int kStatusBadTime = 1;
constexpr int kStatusNotSpd = 2;

// 1/sqrt(x) and 1/x: MUFU.RSQ64H / MUFU.RCP64H seed (rsqrt/rcp.approx.ftz.f64) + the same Newton steps
// the CUDA math library uses, WITHOUT its range checks (pivots and segment times are normal, positive
// doubles here; zero / negative / NaN inputs still come out as inf / NaN and are reported in status[]).
// Dropping the checks removes a BSSY/BRA/BSYNC region per call, across which ptxas cannot interleave the
// independent FMA chains of the block factorisation.  Result within ~1 ulp.
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double e = fma(x, -(y0 * y0), 1.0);           // 1 - x*y0^2
  const double p = fma(e, 0.375, 0.5);
  return fma(p, y0 * e, y0);                           // y0 * (1 + e/2 + 3e^2/8)
}
__device__ __forceinline__ double fast_rcp(double x) {
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
  double e = fma(-x, r0, 1.0);
  e = fma(e, e, e);
  const double r1 = fma(r0, e, r0);                    // r0 * (1 + e + e^2)
  return fma(r1, fma(-x, r1, 1.0), r1);
}

}  // namespace mtg
