// mtg_twisted_tmem_kernel.cuh -- K1 (v3): twisted two-lanes-per-trajectory sweep with the
// per-thread sweep state held in TENSOR MEMORY and coalesced output.
//
// Why TMEM: the sweep state (L_v, inverse pivots, y_v per eliminated vertex; 22 doubles per
// vertex for N=10, D=3) is what bounds the number of trajectories in flight per SM.  Shared
// memory alone gives 5 warps/SM at K = 16.  Blackwell's 256 KB tensor memory is otherwise idle
// on this (tensor-core-free) path, and tcgen05.st / tcgen05.ld with the 32x32b shape give every
// thread of a warp a private, dynamically indexed row of 512 32-bit columns -- exactly a
// per-thread LIFO.  A 128-thread CTA covers the 128 TMEM lanes; two CTAs per SM take 256 columns
// each (128 doubles per thread); what does not fit spills to shared memory [vertex][slot][thread].
//
// Output: each lane writes the D*N coefficients of the segment it has just solved into its row of a
// 128-byte aligned per-warp staging tile ([half][16 trajectories][D*N doubles]); one elected lane then hands
// the two 16-row boxes (forward halves: segment j, reversed halves: segment K-1-j) to the TMA with
// cp.async.bulk.tensor.2d stores against a tensor map of coeffs viewed as [B][K*D*N].  No cooperative
// read-back, no global-store LSU wavefronts, and the ragged last tile is clipped by the tensor map.
//
// Inputs: every lane prefetches the next step's segment time and waypoint with cp.async into a small
// per-thread ring (a register prefetch would share its scoreboard slot with the value being consumed); the
// outward sweep re-reads times from a per-thread shared-memory history and positions from the sweep state.
//
// Mathematics, frames and index maps: see mtg_twisted_kernel.cuh.
#pragma once

#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "mtg_twisted_kernel.cuh"

namespace mtg {
namespace tmem {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 bit, repeated along columns: thread i of the warp owns lane (quarter base + i) and
// moves NV consecutive 32-bit columns starting at taddr.
template <int NV>
__device__ __forceinline__ void st(uint32_t taddr, const uint32_t* v);
template <int NV>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* v);

template <>
__device__ __forceinline__ void st<2>(uint32_t a, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a), "r"(v[0]), "r"(v[1]) : "memory");
}
template <>
__device__ __forceinline__ void st<4>(uint32_t a, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3])
               : "memory");
}
template <>
__device__ __forceinline__ void st<8>(uint32_t a, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
template <>
__device__ __forceinline__ void st<16>(uint32_t a, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16};" ::"r"(a),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
template <>
__device__ __forceinline__ void ld<2>(uint32_t a, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(a) : "memory");
}
template <>
__device__ __forceinline__ void ld<4>(uint32_t a, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(a)
               : "memory");
}
template <>
__device__ __forceinline__ void ld<8>(uint32_t a, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(a)
               : "memory");
}
template <>
__device__ __forceinline__ void ld<16>(uint32_t a, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(a)
      : "memory");
}

// NW 32-bit words starting at column `col` of this thread's lane, greedy chunks of 16/8/4/2.
template <int NW, int OFF = 0>
__device__ __forceinline__ void st_words(uint32_t taddr, const uint32_t* w) {
  if constexpr (NW - OFF >= 16) {
    st<16>(taddr + OFF, w + OFF);
    st_words<NW, OFF + 16>(taddr, w);
  } else if constexpr (NW - OFF >= 8) {
    st<8>(taddr + OFF, w + OFF);
    st_words<NW, OFF + 8>(taddr, w);
  } else if constexpr (NW - OFF >= 4) {
    st<4>(taddr + OFF, w + OFF);
    st_words<NW, OFF + 4>(taddr, w);
  } else if constexpr (NW - OFF >= 2) {
    st<2>(taddr + OFF, w + OFF);
    st_words<NW, OFF + 2>(taddr, w);
  }
}
template <int NW, int OFF = 0>
__device__ __forceinline__ void ld_words(uint32_t taddr, uint32_t* w) {
  if constexpr (NW - OFF >= 16) {
    ld<16>(taddr + OFF, w + OFF);
    ld_words<NW, OFF + 16>(taddr, w);
  } else if constexpr (NW - OFF >= 8) {
    ld<8>(taddr + OFF, w + OFF);
    ld_words<NW, OFF + 8>(taddr, w);
  } else if constexpr (NW - OFF >= 4) {
    ld<4>(taddr + OFF, w + OFF);
    ld_words<NW, OFF + 4>(taddr, w);
  } else if constexpr (NW - OFF >= 2) {
    ld<2>(taddr + OFF, w + OFF);
    ld_words<NW, OFF + 2>(taddr, w);
  }
}

}  // namespace tmem

struct TmemLaunch {
  int n_tmem_blocks;   // eliminated vertices whose state lives in TMEM (the rest spill to shared memory)
  int tmem_cols;       // power of two >= 32, 0 = no TMEM used
};

constexpr int kTmemThreads = 128;
constexpr int kTmemHeaderBytes = 128;  // TMEM base-address holder, padded so that the staging tiles stay 128-B aligned

template <int N, int D>
__host__ __device__ constexpr int tmem_stage_bytes_per_warp() {
  return 32 * D * (N / 2) * 16;  // one whole segment (D*N contiguous output doubles) per lane: two 16-row TMA boxes
}

template <int D>
__host__ __device__ constexpr int tmem_prefetch_bytes() {
  return 2 * (1 + D) * kTmemThreads * 8;
}

__device__ __forceinline__ void cp_async8(const double* smem_dst, const double* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tmem::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// TMA tensor store of a [16 rows][D*N doubles] shared-memory box to coeffs viewed as a 2-D tensor
// [B trajectories][K*D*N doubles]: (c0 = first double inside the trajectory, c1 = first trajectory).  Rows
// beyond the tensor (ragged last tile) are clipped by the hardware.
__device__ __forceinline__ void tma_store_box(const CUtensorMap* tmap, const void* ssrc, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(tmem::smem_u32(ssrc)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int N, int R, int D, bool FUSED = false, bool COST = false>
__global__ void __launch_bounds__(kTmemThreads, (N <= 8 ? 3 : 2))
    twisted_tmem_kernel(const WaypointParams prm, const TmemLaunch tl, const __grid_constant__ CUtensorMap tmap) {
  constexpr int h = N / 2;
  constexpr int m = h - 1;
  constexpr int kL = m * (m + 1) / 2;
  // per eliminated vertex: L (strictly lower) + inverse pivots + y + the vertex position (so that the
  // outward sweep does not re-read it from global memory)
  constexpr int kSlots = kL + m * D + D;
  constexpr int kWords = 2 * kSlots;
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int kWarps = kTmemThreads / 32;
  using G = H1Imm<N, R>;     // immediates: this kernel is register-bound (see mtg_device.cuh)
  using AI = A1InvImm<N>;

  extern __shared__ __align__(128) unsigned char smem_raw[];  // TMA tensor stores need 128-byte aligned tiles
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int half = lane & 1;
  const int K = prm.K;
  const int nf = prm.n_fixed;
  const int M = (K + 1) >> 1;
  const int nh = half ? K - M - 1 : M - 1;
  const int nmax = M - 1;

  // ---- shared memory carve-up: [holder 16 B][staging: kWarps tiles][spilled state]
  uint32_t* holder = reinterpret_cast<uint32_t*>(smem_raw);
  double2* stage = reinterpret_cast<double2*>(smem_raw + kTmemHeaderBytes) + size_t(warp) * 32 * (D * h);  // [half][16][D*h]
  // per-thread prefetch ring for the next step's inputs (segment time + D positions), filled by
  // cp.async: a register prefetch would share its scoreboard slot with the load being consumed and the
  // consumer would wait for the NEW loads as well (measured: 25 % of all stall samples).
  double* pf = reinterpret_cast<double*>(smem_raw + kTmemHeaderBytes + size_t(kWarps) * tmem_stage_bytes_per_warp<N, D>()) + threadIdx.x;
  auto PF = [&](int buf, int slot) -> double* { return pf + (size_t(buf) * (1 + D) + slot) * kTmemThreads; };
  // per-thread history of the own-frame segment times seen by the inward sweep (nmax+1 doubles): the
  // outward sweep reads them back from shared memory (in FUSED mode this also saves the sqrt/exp)
  double* thist = reinterpret_cast<double*>(smem_raw + kTmemHeaderBytes + size_t(kWarps) * tmem_stage_bytes_per_warp<N, D>() +
                                            tmem_prefetch_bytes<D>()) +
                  threadIdx.x;
  auto HT = [&](int j) -> double& { return thist[size_t(j) * kTmemThreads]; };
  double* spill = thist + size_t(nmax + 1) * kTmemThreads;
  auto SP = [&](int blk, int slot) -> double& { return spill[(size_t(blk) * kSlots + slot) * kTmemThreads]; };

  // ---- tensor memory for the sweep state
  uint32_t tbase = 0;  // assigned after the first global loads have been issued (see below)
  const int ntm = tl.n_tmem_blocks;
  auto put_state = [&](int blk, const double (&sv)[kSlots]) {
    if (blk < ntm) {  // warp-uniform
      // one tcgen05.st.x2 per double: a double already is an aligned register pair, so no packing
      // moves are needed (a wide .x16 store wants 16 consecutive registers)
#pragma unroll
      for (int i = 0; i < kSlots; ++i) {
        const uint32_t w[2] = {(uint32_t)__double2loint(sv[i]), (uint32_t)__double2hiint(sv[i])};
        tmem::st<2>(tbase + uint32_t(blk * kWords + 2 * i), w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < kSlots; ++i) SP(blk - ntm, i) = sv[i];
    }
  };
  auto get_state = [&](int blk, double (&sv)[kSlots]) {
    if (blk < ntm) {
      uint32_t w[kWords];
      tmem::ld_words<kWords>(tbase + uint32_t(blk * kWords), w);
      tmem::wait_ld();
#pragma unroll
      for (int i = 0; i < kSlots; ++i) sv[i] = __hiloint2double((int)w[2 * i + 1], (int)w[2 * i]);
    } else {
#pragma unroll
      for (int i = 0; i < kSlots; ++i) sv[i] = SP(blk - ntm, i);
    }
  };

  const long long traj0 = ((long long)blockIdx.x * kWarps + warp) * 16;  // first trajectory of this warp
  long long traj = traj0 + (lane >> 1);
  const bool valid = traj < prm.B;
  if (!valid) traj = prm.B - 1;

  // cost-only mode with the Mellinger expansion generated on the fly: inputs come from the base trajectory
  long long src = traj;
  int mel_n = -1;
  double mel_corr = 0.0;
  if constexpr (COST) {
    if (prm.mel_k1 > 0) {
      src = traj / prm.mel_k1;
      mel_n = int(traj - src * prm.mel_k1) - 1;
      mel_corr = prm.mel_inc / (K - 1.0);
    }
  }
  auto time_of = [&](double raw, int seg_index) -> double {  // seg_index: ORIGINAL segment index
    if constexpr (COST) return mellinger_time(raw, seg_index, mel_n, prm.mel_inc, mel_corr, prm.mel_lower);
    return raw;
  };
  const double* __restrict__ tt = FUSED ? nullptr : prm.times + src * K;
  const double* __restrict__ fx =
      FUSED ? prm.positions + src * (long long)(K + 1) * D : prm.dfix + src * (long long)D * nf;
  auto seg = [&](int j) -> int { return half ? K - 1 - j : j; };
  auto pidx = [&](int v) -> int {
    const int o = half ? K - v : v;
    return o == 0 ? 0 : (o < K ? h + o - 1 : h + K - 1);
  };
  auto sgn = [&](int idx) -> double { return (half && !(idx & 1)) ? -1.0 : 1.0; };
  // address of coordinate d of own-frame vertex v
  auto xaddr = [&](int v, int d) -> const double* {
    if constexpr (FUSED) {
      return fx + (half ? K - v : v) * D + d;
    } else {
      return fx + d * nf + pidx(v);
    }
  };
  // prefetch (time of own segment j, position of own vertex v) into ring buffer `buf`
  auto pf_issue = [&](int buf, int j, int v) {
    if constexpr (!FUSED) cp_async8(PF(buf, 0), tt + seg(j));
#pragma unroll
    for (int d = 0; d < D; ++d) cp_async8(PF(buf, 1 + d), xaddr(v, d));
  };
  double* __restrict__ tout = (FUSED && prm.times_out != nullptr) ? prm.times_out + traj * K : nullptr;

  // ---- issue the first global loads NOW: their latency overlaps the TMEM allocation, the CTA barrier
  // and the index set-up below (measured: the prologue loads were ~9 % of all stall samples)
  double T0e = 0.0, x0e[D], x1e[D], u0e[m][D];
  {
    const int e0 = half ? h + K : 1;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      x0e[d] = __ldg(xaddr(0, d));
      x1e[d] = __ldg(xaddr(1, d));
#pragma unroll
      for (int b = 0; b < m; ++b) u0e[b][d] = FUSED ? 0.0 : __ldg(fx + d * nf + e0 + b);
    }
    if constexpr (!FUSED) T0e = __ldg(tt + seg(0));
    pf_issue(1, 1, 2);  // inputs of sweep step v = 1 -> ring buffer (v & 1)
  }

  // ---- tensor memory for the sweep state
  if (tl.tmem_cols > 0) {
    if (warp == 0) tmem::alloc(tmem::smem_u32(holder), (uint32_t)tl.tmem_cols);
    tmem::fence_before_sync();
    __syncthreads();
    tmem::fence_after_sync();
    tbase = *holder + (uint32_t(warp * 32) << 16);  // this warp's lane quarter
  }

  // ---- output: each lane writes the D*N doubles of the segment it emits into row (half*16 + trajectory) of
  // the warp's staging tile; one elected lane hands the two 16-row boxes (forward halves: segment j, reversed
  // halves: segment K-1-j) to the TMA.  No cooperative read-back and no global-store LSU wavefronts.
  double2* my_row = stage + ((lane & 1) * 16 + (lane >> 1)) * (D * h);
  const int nhF = M - 1, nhB = K - M - 1;  // a half is active in sweep step v iff v <= its nh

  // emit own-frame segment j for every lane of the warp at once (convergent).  `act`: this lane's
  // values are meaningful; rows of inactive lanes are not stored.  v_step: the sweep step (0 = final).
  double cost_acc = 0.0;
  auto emit_all = [&](int j, int v_step, double T, double iT, const double (&sd)[h][D], const double (&ed)[h][D]) {
    if constexpr (COST) {
      // 0.5 d^T H(T) d of this segment, d = [start derivatives; end derivatives] in the own frame (the cost is
      // invariant under the time reversal of the odd half): H = T^(1-2r) S G S / scale, S = diag(T^(s mod h)).
      (void)j;
      const int nh_own = half ? nhB : nhF;
      if (v_step <= nh_own) {
        double tp[h];
        tp[0] = 1.0;
#pragma unroll
        for (int k = 1; k < h; ++k) tp[k] = tp[k - 1] * T;
        double q = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          double u[N];
#pragma unroll
          for (int k = 0; k < h; ++k) {
            u[k] = tp[k] * sd[k][d];
            u[h + k] = tp[k] * ed[k][d];
          }
#pragma unroll
          for (int s2 = 0; s2 < N; ++s2) {
            double row = 0.5 * G::at(s2, s2) * u[s2];
#pragma unroll
            for (int t2 = s2 + 1; t2 < N; ++t2) row = fma(G::at(s2, t2), u[t2], row);
            q = fma(row, u[s2], q);
          }
        }
        // T^(1-2r) = T * (1/T)^(2r)
        cost_acc = fma(q, T * pow_int<2 * R>(iT), cost_acc);
      }
      return;
    }
    // original orientation: start = J*(own end) for the reversed half; J folded into the powers
    double tp[h], itp[h];
    const double Ts = half ? -T : T;
    tp[0] = 1.0;
#pragma unroll
    for (int k = 1; k < h; ++k) tp[k] = tp[k - 1] * Ts;
    itp[0] = pow_int<h>(iT);
#pragma unroll
    for (int k = 1; k < h; ++k) itp[k] = itp[k - 1] * iT;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      double c[N], ss[h], se[h];
#pragma unroll
      for (int k = 0; k < h; ++k) {
        const double s0 = half ? ed[k][d] : sd[k][d];
        const double e0 = half ? sd[k][d] : ed[k][d];
        c[k] = s0 * ((half && (k & 1)) ? -AI::at(k, k) : AI::at(k, k));
        ss[k] = tp[k] * s0;
        se[k] = tp[k] * e0;
      }
      // Upper coefficients in Hermite form: A(1)^-1 = [[L^-1, 0], [-D^-1 C L^-1, D^-1]] and (C L^-1)[k][j] =
      // 1/(j-k)! (derivative k of the Taylor part at tau = 1), so  q = D^-1 (se - C L^-1 ss):
      // h(h+1)/2 + h^2 operations instead of 2 h^2, and the 1/(j-k)! factors are mostly dyadic immediates.
      double ee[h];
#pragma unroll
      for (int k = 0; k < h; ++k) {
        double acc = se[k] - ss[k];
#pragma unroll
        for (int j = k + 1; j < h; ++j) {
          constexpr double kInvFact[6] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0};
          acc = (j - k == 1) ? acc - ss[j] : fma(-kInvFact[j - k], ss[j], acc);
        }
        ee[k] = acc;
      }
#pragma unroll
      for (int q = 0; q < h; ++q) {
        double acc = AI::at(h + q, h) * ee[0];
#pragma unroll
        for (int k = 1; k < h; ++k) acc = fma(AI::at(h + q, h + k), ee[k], acc);
        c[h + q] = acc * itp[q];
      }
      if (d == 0) {  // the TMA must have finished reading the previous segment's tile
        if (lane == 0) bulk_wait_read();
        __syncwarp();
      }
#pragma unroll
      for (int q = 0; q < h; ++q) my_row[d * h + q] = make_double2(c[2 * q], c[2 * q + 1]);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      if (v_step <= nhF) tma_store_box(&tmap, stage, j * (D * N), (int)traj0);
      if (v_step <= nhB) tma_store_box(&tmap, stage + 16 * (D * h), (K - 1 - j) * (D * N), (int)traj0);
      bulk_commit();
    }
  };

  int stat = 0;
  double Wp[m][m], yp[m][D], Cee[m][m], cps[m], cpe[m], bcar[m][D], xm[D], xc[D];
  {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      xm[d] = x0e[d];
      xc[d] = x1e[d];
    }
    double T0;
    if constexpr (FUSED) {
      T0 = nfabian_time<D>(xm, xc, prm.v_max, prm.a_max, prm.magic);
    } else {
      T0 = time_of(T0e, seg(0));
    }
    if (!(T0 > 0.0)) stat |= kStatusBadTime;
    HT(0) = T0;
    const double iT0 = fast_rcp(T0);
    double pw[N - 1];
    segment_powers<N, R>(T0, iT0, pw);
#pragma unroll
    for (int a = 0; a < m; ++a) {
#pragma unroll
      for (int b = 0; b < m; ++b) {
        Cee[a][b] = pw[a + b + 2] * G::at(h + 1 + a, h + 1 + b);
        Wp[a][b] = 0.0;
      }
      cps[a] = pw[a + 1] * G::at(h + 1 + a, 0);
      cpe[a] = pw[a + 1] * G::at(h + 1 + a, h);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
      double u0[m];
#pragma unroll
      for (int b = 0; b < m; ++b) u0[b] = sgn(b) * u0e[b][d];
#pragma unroll
      for (int a = 0; a < m; ++a) {
        double acc = 0.0;
#pragma unroll
        for (int b = 0; b < m; ++b) acc = fma(pw[a + b + 2] * G::at(h + 1 + a, 1 + b), u0[b], acc);
        bcar[a][d] = -acc;
        yp[a][d] = 0.0;
      }
    }
  }

  // ---------------------------------------------------------------- sweep towards the middle
  for (int v = 1; v <= nmax; ++v) {
    double sv[kSlots];  // lanes with v > nh store whatever is here; they never read it back
    if (v <= nh) {
      cp_async_wait_all();
      double xn[D];
#pragma unroll
      for (int d = 0; d < D; ++d) xn[d] = *PF(v & 1, 1 + d);
      double T;
      if constexpr (FUSED) {
        T = nfabian_time<D>(xc, xn, prm.v_max, prm.a_max, prm.magic);
      } else {
        T = time_of(*PF(v & 1, 0), seg(v));
      }
      HT(v) = T;
      {  // prefetch the next step's inputs (clamped indices: never out of bounds)
        const int jn = v + 1 < K ? v + 1 : K - 1;
        const int vn = v + 2 <= K ? v + 2 : K;
        pf_issue((v + 1) & 1, jn, vn);
      }
      if (!(T > 0.0)) stat |= kStatusBadTime;
      const double iT = fast_rcp(T);
      double pw[N - 1];
      segment_powers<N, R>(T, iT, pw);

      double Dp[m][m], E[m][m], bb[m][D];
#pragma unroll
      for (int a = 0; a < m; ++a) {
#pragma unroll
        for (int b = 0; b <= a; ++b) {
          double s = fma(pw[a + b + 2], G::at(1 + a, 1 + b), Cee[a][b]);
#pragma unroll
          for (int k = 0; k < m; ++k) s = fma(-Wp[k][a], Wp[k][b], s);
          Dp[a][b] = s;
        }
#pragma unroll
        for (int b = 0; b < m; ++b) E[a][b] = pw[a + b + 2] * G::at(1 + a, h + 1 + b);
        const double gmid = fma(pw[a + 1], G::at(1 + a, 0), cpe[a]);
        const double gnext = pw[a + 1] * G::at(1 + a, h);
#pragma unroll
        for (int d = 0; d < D; ++d) {
          double s = bcar[a][d];
          s = fma(-cps[a], xm[d], s);
          s = fma(-gmid, xc[d], s);
          s = fma(-gnext, xn[d], s);
#pragma unroll
          for (int k = 0; k < m; ++k) s = fma(-Wp[k][a], yp[k][d], s);
          bb[a][d] = s;
        }
      }
      double L[m][m], inv[m];
#pragma unroll
      for (int j = 0; j < m; ++j) {
        double s = Dp[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fma(-L[j][k], L[j][k], s);
        if (!(s > 0.0)) stat |= kStatusNotSpd;
        inv[j] = fast_rsqrt(s);
#pragma unroll
        for (int i = j + 1; i < m; ++i) {
          double t = Dp[i][j];
#pragma unroll
          for (int k = 0; k < j; ++k) t = fma(-L[i][k], L[j][k], t);
          L[i][j] = t * inv[j];
        }
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
#pragma unroll
        for (int j = 0; j < m; ++j) {
          double s = bb[j][d];
#pragma unroll
          for (int k = 0; k < j; ++k) s = fma(-L[j][k], yp[k][d], s);
          yp[j][d] = s * inv[j];
        }
      }
#pragma unroll
      for (int c = 0; c < m; ++c) {
#pragma unroll
        for (int j = 0; j < m; ++j) {
          double s = E[j][c];
#pragma unroll
          for (int k = 0; k < j; ++k) s = fma(-L[j][k], Wp[k][c], s);
          Wp[j][c] = s * inv[j];
        }
      }
      {
        int slot = 0;
#pragma unroll
        for (int i = 1; i < m; ++i)
#pragma unroll
          for (int j = 0; j < i; ++j) sv[slot++] = L[i][j];
#pragma unroll
        for (int j = 0; j < m; ++j) sv[slot++] = inv[j];
#pragma unroll
        for (int j = 0; j < m; ++j)
#pragma unroll
          for (int d = 0; d < D; ++d) sv[slot++] = yp[j][d];
#pragma unroll
        for (int d = 0; d < D; ++d) sv[slot++] = xc[d];  // position of the vertex just eliminated
      }
#pragma unroll
      for (int a = 0; a < m; ++a) {
#pragma unroll
        for (int b = 0; b <= a; ++b) Cee[a][b] = pw[a + b + 2] * G::at(h + 1 + a, h + 1 + b);
        cps[a] = pw[a + 1] * G::at(h + 1 + a, 0);
        cpe[a] = pw[a + 1] * G::at(h + 1 + a, h);
#pragma unroll
        for (int d = 0; d < D; ++d) bcar[a][d] = 0.0;
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
        xm[d] = xc[d];
        xc[d] = xn[d];
      }
    }
    __syncwarp();
    put_state(v - 1, sv);  // all lanes (tcgen05.st is warp-collective)
  }
  __syncwarp();
  if (ntm > 0) tmem::wait_st();

  // ---------------------------------------------------------------- middle vertex
  double um[m][D];
  {
    double Dl[m][m], bl[m][D];
#pragma unroll
    for (int a = 0; a < m; ++a) {
#pragma unroll
      for (int b = 0; b <= a; ++b) {
        double s = Cee[a][b];
#pragma unroll
        for (int k = 0; k < m; ++k) s = fma(-Wp[k][a], Wp[k][b], s);
        Dl[a][b] = s;
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
        double s = bcar[a][d];
        s = fma(-cps[a], xm[d], s);
        s = fma(-cpe[a], xc[d], s);
#pragma unroll
        for (int k = 0; k < m; ++k) s = fma(-Wp[k][a], yp[k][d], s);
        bl[a][d] = s;
      }
    }
#pragma unroll
    for (int a = 0; a < m; ++a) {
#pragma unroll
      for (int b = 0; b <= a; ++b) {
        const double o = __shfl_xor_sync(kFull, Dl[a][b], 1);
        Dl[a][b] += ((a + b) & 1) ? -o : o;
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double o = __shfl_xor_sync(kFull, bl[a][d], 1);
        bl[a][d] += (a & 1) ? o : -o;
      }
    }
    stat |= __shfl_xor_sync(kFull, stat, 1);
    double L[m][m], inv[m];
#pragma unroll
    for (int j = 0; j < m; ++j) {
      double s = Dl[j][j];
#pragma unroll
      for (int k = 0; k < j; ++k) s = fma(-L[j][k], L[j][k], s);
      if (!(s > 0.0)) stat |= kStatusNotSpd;
      inv[j] = fast_rsqrt(s);
#pragma unroll
      for (int i = j + 1; i < m; ++i) {
        double t = Dl[i][j];
#pragma unroll
        for (int k = 0; k < j; ++k) t = fma(-L[i][k], L[j][k], t);
        L[i][j] = t * inv[j];
      }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
      double y[m];
#pragma unroll
      for (int j = 0; j < m; ++j) {
        double s = bl[j][d];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fma(-L[j][k], y[k], s);
        y[j] = s * inv[j];
      }
#pragma unroll
      for (int j = m - 1; j >= 0; --j) {
        double s = y[j];
#pragma unroll
        for (int k = j + 1; k < m; ++k) s = fma(-L[k][j], um[k][d], s);
        um[j][d] = s * inv[j];
      }
    }
  }
  if (valid && half == 0 && prm.status != nullptr) prm.status[traj] = stat;

  // ---------------------------------------------------------------- outward back-substitution
  const int np = (K - 1) * m;
  double* __restrict__ df = prm.dfree != nullptr ? prm.dfree + traj * (long long)D * np : nullptr;
  auto store_free = [&](int v_own, const double (&u)[h][D]) {
    if (df != nullptr && valid) {
      const int vo = half ? K - v_own : v_own;
#pragma unroll
      for (int d = 0; d < D; ++d)
#pragma unroll
        for (int j = 0; j < m; ++j) df[d * np + (vo - 1) * m + j] = sgn(j) * u[1 + j][d];
    }
  };

  double ed[h][D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    ed[0][d] = xc[d];
#pragma unroll
    for (int j = 0; j < m; ++j) ed[1 + j][d] = um[j][d];
  }
  if (half == 0) store_free(nh + 1, ed);

  // outward steps take the vertex position from the sweep state; only the segment time is prefetched
  // (and the position of own vertex 0 for the final emission)
  auto pf_issue_out = [&](int j) {
    if (j == 0) pf_issue(0, 0, 0);  // position of own vertex 0 for the final emission (time comes from HT)
  };
  pf_issue_out(nh);

  for (int v = nmax; v >= 1; --v) {
    double sv[kSlots];
    get_state(v - 1, sv);  // all lanes
    const bool act = v <= nh;
    double T = 1.0, iT = 1.0;
    double sd[h][D];  // inactive lanes (odd K only) emit garbage rows that are never stored
    if (act) {
      cp_async_wait_all();
      double xv[D];
#pragma unroll
      for (int d = 0; d < D; ++d) xv[d] = sv[kL + m * D + d];
      T = HT(v);
      if constexpr (FUSED) {
        if (tout != nullptr && valid) tout[seg(v)] = T;
      }
      pf_issue_out(v - 1);
      if constexpr (!FUSED) {
        if (v == 1) {  // the final emission re-reads the fixed end derivatives: pull their lines into L1 now
          const int e0 = half ? h + K : 1;
#pragma unroll
          for (int d = 0; d < D; ++d) asm volatile("prefetch.global.L1 [%0];" ::"l"(fx + d * nf + e0));
        }
      }
      iT = fast_rcp(T);
      double L[m][m], inv[m], rhs[m][D];
      {
        int slot = 0;
#pragma unroll
        for (int i = 1; i < m; ++i)
#pragma unroll
          for (int j = 0; j < i; ++j) L[i][j] = sv[slot++];
#pragma unroll
        for (int j = 0; j < m; ++j) inv[j] = sv[slot++];
#pragma unroll
        for (int j = 0; j < m; ++j)
#pragma unroll
          for (int d = 0; d < D; ++d) rhs[j][d] = sv[slot++];
      }
      double pw[N - 1];
      segment_powers<N, R>(T, iT, pw);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        double t[m];
#pragma unroll
        for (int a = 0; a < m; ++a) {
          double s = 0.0;
#pragma unroll
          for (int b = 0; b < m; ++b) s = fma(pw[a + b + 2] * G::at(1 + a, h + 1 + b), ed[1 + b][d], s);
          t[a] = s;
        }
#pragma unroll
        for (int j = 0; j < m; ++j) {
          double s = t[j];
#pragma unroll
          for (int k = 0; k < j; ++k) s = fma(-L[j][k], t[k], s);
          t[j] = s * inv[j];
          rhs[j][d] -= t[j];
        }
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
#pragma unroll
        for (int j = m - 1; j >= 0; --j) {
          double s = rhs[j][d];
#pragma unroll
          for (int k = j + 1; k < m; ++k) s = fma(-L[k][j], sd[1 + k][d], s);
          sd[1 + j][d] = s * inv[j];
        }
        sd[0][d] = xv[d];
      }
      store_free(v, sd);
    }
    __syncwarp();
    emit_all(v, v, T, iT, sd, ed);
    if (act) {
#pragma unroll
      for (int d = 0; d < D; ++d)
#pragma unroll
        for (int k = 0; k < h; ++k) ed[k][d] = sd[k][d];
    }
  }
  {
    cp_async_wait_all();
    const int e0 = half ? h + K : 1;
    double sd[h][D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      sd[0][d] = *PF(0, 1 + d);
#pragma unroll
      for (int b = 0; b < m; ++b) sd[1 + b][d] = FUSED ? 0.0 : sgn(b) * __ldg(fx + d * nf + e0 + b);
    }
    const double T = HT(0);
    if constexpr (FUSED) {
      if (tout != nullptr && valid) tout[seg(0)] = T;
    }
    const double iT = fast_rcp(T);
    __syncwarp();
    emit_all(0, 0, T, iT, sd, ed);
  }

  if constexpr (COST) {
    // q accumulated sum_{s<=t} (1 or 1/2) G u u = 0.5 u^T G u per segment: cost = sum over both halves / scale
    const double other = __shfl_xor_sync(kFull, cost_acc, 1);
    if (valid && half == 0 && prm.cost != nullptr) prm.cost[traj] = (cost_acc + other) * (1.0 / G::scale);
  }
  if (lane == 0) bulk_wait_all();  // every TMA store issued by this warp has completed
  if (tl.tmem_cols > 0) {
    tmem::fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem::dealloc(*holder, (uint32_t)tl.tmem_cols);
  }
}

}  // namespace mtg
