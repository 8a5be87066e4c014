// mtg_twisted_tmem_v4_kernel.cuh -- K1 (v4): the twisted TMEM kernel made PERSISTENT, with every global
// read issued several sweep steps (or half a tile) before it is consumed.
//
// What the ncu source page of v3 showed (profiles/r01_tmem_c3.ncu-rep): 22 % of all warp-stall samples are
// long-scoreboard waits on INPUT data -- the first loads of a tile (4.4 %), the cp.async ring that ran one
// step ahead (4.3 % + 2.9 %) -- and 4 % are CTA barriers around the per-CTA TMEM allocation.  The kernel's
// HBM traffic is 83 % writes, and reads queued behind write bursts take 2-3 us, longer than one sweep step.
// v4 therefore
//   * runs one CTA per TMEM/shared-memory slot for the whole launch (2 per SM); every WARP loops over its own
//     16-trajectory tiles (tile = blockIdx.x * 4 + warp, stride gridDim.x * 4): TMEM is allocated once, no CTA
//     barrier inside the loop;
//   * prefetches the NEXT tile's prologue inputs (end vertex, its derivatives, first waypoint, first time)
//     with cp.async half a tile ahead, into the part of shared memory that holds the spilled sweep state
//     (dead by then);
//   * deepens the per-thread input ring to RD buffers (prefetch distance RD-1 sweep steps, cp.async groups);
//   * keeps the end vertex position in a 3-double per-thread stash and prefetches the end derivatives for the
//     final emission at the start of the outward sweep -- no exposed global read at the end of a tile;
//   * folds the first step's right-hand-side carry into the (W, y) carry with an exact power-of-two scaling
//     (W = 2^-600 I, y = -2^600 b): the m*D carry registers disappear from the loop.
// The arithmetic of a trajectory is otherwise the v3 sequence (mtg_twisted_tmem_kernel.cuh); results agree to
// rounding (bitwise except for the order in which the first step's carry is added).
#pragma once

#include "mtg_twisted_tmem_kernel.cuh"

namespace mtg {

struct TmemLaunchV4 {
  int n_tmem_blocks;  // eliminated vertices whose state lives in TMEM (the rest spill to shared memory)
  int tmem_cols;      // power of two >= 32, 0 = no TMEM used
  int region_slots;   // doubles per thread of the spill / next-tile-prologue region
  unsigned long long* tile_counter;  // non-null: warps draw their 16-trajectory tiles from this counter (zeroed by
                                     // the host before the launch); null: static round-robin assignment
};

template <int N, int D>
__host__ __device__ constexpr int v4_pro_slots() {
  return 2 * D + (N / 2 - 1) * D + 1;  // x0, x1, u0[m], T0
}
template <int N, int D>
__host__ __device__ constexpr int v4_state_slots() {
  constexpr int m = N / 2 - 1;
  return m * (m + 1) / 2 + m * D + D;
}
// dynamic shared memory of a CTA
template <int N, int D, int RD>
__host__ __device__ constexpr size_t v4_smem_bytes(int K, int ntm) {
  const int nmax = (K + 1) / 2 - 1;
  const int spill = (nmax - ntm) > 0 ? (nmax - ntm) * v4_state_slots<N, D>() : 0;
  const int region = spill > v4_pro_slots<N, D>() ? spill : v4_pro_slots<N, D>();
  return size_t(kTmemHeaderBytes) + size_t(kTmemThreads / 32) * tmem_stage_bytes_per_warp<N, D>() +
         size_t(RD * (1 + D) + (nmax + 1) + D + region) * kTmemThreads * 8;
}

template <int NPEND>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

template <int N, int R, int D, bool FUSED, int RD, int MINB, int HOIST = 0>
__global__ void __launch_bounds__(kTmemThreads, MINB)
    twisted_tmem_v4_kernel(const WaypointParams prm, const TmemLaunchV4 tl, const __grid_constant__ CUtensorMap tmap) {
  constexpr int h = N / 2;
  constexpr int m = h - 1;
  constexpr int kL = m * (m + 1) / 2;
  constexpr int kSlots = kL + m * D + D;
  constexpr int kWords = 2 * kSlots;
  constexpr int kPro = v4_pro_slots<N, D>();
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int kWarps = kTmemThreads / 32;
  constexpr double kTiny = 0x1p-600, kHuge = 0x1p+600;
  static_assert(RD >= 2, "ring depth");
  // work overlapped with the asynchronous tensor-memory read of the outward sweep: 0 = none, 1 = segment time and
  // its powers, 2 = also E_v u_{v+1} (the fetched words stay live meanwhile: ~50 registers).  Measured on C3
  // (profiles/r02_k1_variants.json): 0.563 / 0.511 / 0.511 of the HBM roofline for 0 / 1 / 2 -- the extra live
  // registers cost more than the exposed tcgen05.wait::ld, so 0 is the default.
  constexpr int kHoist = HOIST;
  constexpr bool kHoistE = HOIST >= 2;
  using G = H1Imm<N, R>;
  using AI = A1InvImm<N>;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int half = lane & 1;
  const int K = prm.K;
  const int nf = prm.n_fixed;
  const int M = (K + 1) >> 1;
  const int nh = half ? K - M - 1 : M - 1;
  const int nmax = M - 1;
  const int ntm = tl.n_tmem_blocks;

  // ---- shared memory: [holder][staging x kWarps][ring RD x (1+D)][time history nmax+1][x0 stash D][region]
  uint32_t* holder = reinterpret_cast<uint32_t*>(smem_raw);
  double2* stage = reinterpret_cast<double2*>(smem_raw + kTmemHeaderBytes) + size_t(warp) * 32 * (D * h);
  double* base = reinterpret_cast<double*>(smem_raw + kTmemHeaderBytes + size_t(kWarps) * tmem_stage_bytes_per_warp<N, D>()) +
                 threadIdx.x;
  auto PF = [&](int buf, int slot) -> double* { return base + (size_t(buf) * (1 + D) + slot) * kTmemThreads; };
  double* thist = base + size_t(RD) * (1 + D) * kTmemThreads;
  auto HT = [&](int j) -> double& { return thist[size_t(j) * kTmemThreads]; };
  double* x0s = thist + size_t(nmax + 1) * kTmemThreads;
  double* region = x0s + size_t(D) * kTmemThreads;  // spilled state blocks; next tile's prologue inputs
  auto SP = [&](int blk, int slot) -> double& { return region[(size_t(blk) * kSlots + slot) * kTmemThreads]; };
  auto PRO = [&](int slot) -> double* { return region + size_t(slot) * kTmemThreads; };

  // ---- tensor memory, once per CTA
  uint32_t tbase = 0;
  if (tl.tmem_cols > 0) {
    if (warp == 0) tmem::alloc(tmem::smem_u32(holder), (uint32_t)tl.tmem_cols);
    tmem::fence_before_sync();
    __syncthreads();
    tmem::fence_after_sync();
    tbase = *holder + (uint32_t(warp * 32) << 16);
  }
  auto put_state = [&](int blk, const double (&sv)[kSlots]) {
    if (blk < ntm) {
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
  // The tensor-memory read is asynchronous until tcgen05.wait::ld: state_issue() starts it, the caller does the
  // work that does not depend on the state (segment time, its powers, E_v u_{v+1}), state_finish() waits.
  auto state_issue = [&](int blk, uint32_t (&w)[kWords]) {
    if (blk < ntm) tmem::ld_words<kWords>(tbase + uint32_t(blk * kWords), w);
  };
  auto state_finish = [&](int blk, const uint32_t (&w)[kWords], double (&sv)[kSlots]) {
    if (blk < ntm) {
      tmem::wait_ld();
#pragma unroll
      for (int i = 0; i < kSlots; ++i) sv[i] = __hiloint2double((int)w[2 * i + 1], (int)w[2 * i]);
    } else {
#pragma unroll
      for (int i = 0; i < kSlots; ++i) sv[i] = SP(blk - ntm, i);
    }
  };

  auto seg = [&](int j) -> int { return half ? K - 1 - j : j; };
  auto pidx = [&](int v) -> int {
    const int o = half ? K - v : v;
    return o == 0 ? 0 : (o < K ? h + o - 1 : h + K - 1);
  };
  auto sgn = [&](int idx) -> double { return (half && !(idx & 1)) ? -1.0 : 1.0; };
  const int e0 = half ? h + K : 1;  // first fixed end-derivative slot of own-frame vertex 0

  const long long n_wtiles = (prm.B + 15) >> 4;
  const long long wt_stride = (long long)gridDim.x * kWarps;
  const bool dyn = tl.tile_counter != nullptr;
  // lane 0 draws the tile; the broadcast is deferred to the first use so that the atomic's round trip to L2
  // overlaps the sweep
  auto draw_tile = [&]() -> long long { return lane == 0 ? (long long)atomicAdd(tl.tile_counter, 1ULL) : 0; };
  long long wt = dyn ? __shfl_sync(kFull, draw_tile(), 0) : (long long)blockIdx.x * kWarps + warp;

  // pointers of a warp tile's trajectory for this lane
  struct Ptrs {
    const double* tt;
    const double* fx;
    long long traj;
    bool valid;
  };
  auto tile_ptrs = [&](long long w) -> Ptrs {
    Ptrs p;
    p.traj = w * 16 + (lane >> 1);
    p.valid = p.traj < prm.B;
    if (!p.valid) p.traj = prm.B - 1;
    p.tt = FUSED ? nullptr : prm.times + p.traj * K;
    p.fx = FUSED ? prm.positions + p.traj * (long long)(K + 1) * D : prm.dfix + p.traj * (long long)D * nf;
    return p;
  };
  auto xaddr = [&](const Ptrs& p, int v, int d) -> const double* {
    if constexpr (FUSED) {
      return p.fx + (half ? K - v : v) * D + d;
    } else {
      return p.fx + d * nf + pidx(v);
    }
  };
  // next tile's prologue inputs -> PRO region (cp.async; the caller commits the group)
  auto pro_issue = [&](const Ptrs& p) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      cp_async8(PRO(d), xaddr(p, 0, d));
      cp_async8(PRO(D + d), xaddr(p, 1, d));
      if constexpr (!FUSED) {
#pragma unroll
        for (int b = 0; b < m; ++b) cp_async8(PRO(2 * D + b * D + d), p.fx + d * nf + e0 + b);
      }
    }
    if constexpr (!FUSED) cp_async8(PRO(kPro - 1), p.tt + seg(0));
  };
  // inputs of inward step v (time of own segment v, position of own vertex v+1) -> ring buffer v % RD
  auto ring_issue = [&](const Ptrs& p, int v) {
    const int j = v < K ? v : K - 1;
    const int vn = v + 1 <= K ? v + 1 : K;
    const int buf = v % RD;
    if constexpr (!FUSED) cp_async8(PF(buf, 0), p.tt + seg(j));
#pragma unroll
    for (int d = 0; d < D; ++d) cp_async8(PF(buf, 1 + d), xaddr(p, vn, d));
  };

  if (wt < n_wtiles) {
    const Ptrs p0 = tile_ptrs(wt);
    pro_issue(p0);
    cp_async_commit();
  }

  double2* my_row = stage + ((lane & 1) * 16 + (lane >> 1)) * (D * h);
  const int nhF = M - 1, nhB = K - M - 1;
  // outward step at which the next tile's prologue prefetch is issued: the spill blocks (index >= ntm) have all
  // been read back once step v <= ntm starts
  const int v_pro = ntm < nmax ? ntm : nmax;

  while (wt < n_wtiles) {
    long long wt_next = dyn ? draw_tile() : wt + wt_stride;  // known one tile ahead: its prologue is prefetched
    bool next_known = !dyn;
    const Ptrs P = tile_ptrs(wt);
    const long long traj0 = wt * 16;
    const bool valid = P.valid;
    double* __restrict__ tout = (FUSED && prm.times_out != nullptr) ? prm.times_out + P.traj * K : nullptr;

    // emit own-frame segment j for every lane of the warp at once (convergent)
    auto emit_all = [&](int j, int v_step, double T, double iT, const double (&sd)[h][D], const double (&ed)[h][D]) {
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
          const double e0v = half ? sd[k][d] : ed[k][d];
          c[k] = s0 * ((half && (k & 1)) ? -AI::at(k, k) : AI::at(k, k));
          ss[k] = tp[k] * s0;
          se[k] = tp[k] * e0v;
        }
        double ee[h];
#pragma unroll
        for (int k = 0; k < h; ++k) {
          double acc = se[k] - ss[k];
#pragma unroll
          for (int j2 = k + 1; j2 < h; ++j2) {
            constexpr double kInvFact[6] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0};
            acc = (j2 - k == 1) ? acc - ss[j2] : fma(-kInvFact[j2 - k], ss[j2], acc);
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

    // ---- ring prefetch of the first RD-1 inward steps, then consume the prologue inputs (issued half a tile ago)
    // (a group is committed for every step even when it is empty -- steps beyond this lane's own range --
    // so that wait_group<RD-2> always means "the data of the current step has landed")
#pragma unroll
    for (int q = 1; q < RD; ++q) {
      if (q <= nh) ring_issue(P, q);
      cp_async_commit();
    }
    cp_async_wait_group<RD - 1>();

    int stat = 0;
    double Wp[m][m], yp[m][D], Cee[m][m], cps[m], cpe[m], xm[D], xc[D];
    {
      double T0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        xm[d] = *PRO(d);
        xc[d] = *PRO(D + d);
        x0s[size_t(d) * kTmemThreads] = xm[d];
      }
      if constexpr (FUSED) {
        T0 = nfabian_time<D>(xm, xc, prm.v_max, prm.a_max, prm.magic);
      } else {
        T0 = *PRO(kPro - 1);
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
          Wp[a][b] = (a == b) ? kTiny : 0.0;
        }
        cps[a] = pw[a + 1] * G::at(h + 1 + a, 0);
        cpe[a] = pw[a + 1] * G::at(h + 1 + a, h);
      }
      // carry of the fixed end derivatives: b_1 -= H_0[end,start] u_0, stored as y = 2^600 * (H u_0) against
      // W = 2^-600 I, so that  -W^T y  reproduces it exactly and  W^T W  underflows to zero
#pragma unroll
      for (int d = 0; d < D; ++d) {
        double u0[m];
#pragma unroll
        for (int b = 0; b < m; ++b) u0[b] = FUSED ? 0.0 : sgn(b) * *PRO(2 * D + b * D + d);
#pragma unroll
        for (int a = 0; a < m; ++a) {
          double acc = 0.0;
#pragma unroll
          for (int b = 0; b < m; ++b) acc = fma(pw[a + b + 2] * G::at(h + 1 + a, 1 + b), u0[b], acc);
          yp[a][d] = acc * kHuge;
        }
      }
    }

    // ---------------------------------------------------------------- sweep towards the middle
    for (int v = 1; v <= nmax; ++v) {
      double sv[kSlots];
      if (v <= nh) {
        cp_async_wait_group<RD - 2>();
        double xn[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xn[d] = *PF(v % RD, 1 + d);
        double T;
        if constexpr (FUSED) {
          T = nfabian_time<D>(xc, xn, prm.v_max, prm.a_max, prm.magic);
        } else {
          T = *PF(v % RD, 0);
        }
        HT(v) = T;
        if (v + RD - 1 <= nh) ring_issue(P, v + RD - 1);  // nothing is left in flight after the last own step
        cp_async_commit();
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
            double s = -cps[a] * xm[d];
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
          for (int d = 0; d < D; ++d) sv[slot++] = xc[d];
        }
#pragma unroll
        for (int a = 0; a < m; ++a) {
#pragma unroll
          for (int b = 0; b <= a; ++b) Cee[a][b] = pw[a + b + 2] * G::at(h + 1 + a, h + 1 + b);
          cps[a] = pw[a + 1] * G::at(h + 1 + a, 0);
          cpe[a] = pw[a + 1] * G::at(h + 1 + a, h);
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
          xm[d] = xc[d];
          xc[d] = xn[d];
        }
      }
      __syncwarp();
      put_state(v - 1, sv);
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
          double s = -cps[a] * xm[d];
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
    if (valid && half == 0 && prm.status != nullptr) prm.status[P.traj] = stat;

    // ---------------------------------------------------------------- outward back-substitution
    const int np = (K - 1) * m;
    double* __restrict__ df = prm.dfree != nullptr ? prm.dfree + P.traj * (long long)D * np : nullptr;
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

    // The ring is idle during the outward sweep: the fixed end derivatives needed by the final emission are
    // fetched into it now (slot q of the flattened ring), many steps ahead of their use.
    constexpr bool kEndInRing = !FUSED && (RD * (1 + D) >= m * D);
    if constexpr (kEndInRing) {
#pragma unroll
      for (int d = 0; d < D; ++d)
#pragma unroll
        for (int b = 0; b < m; ++b) cp_async8(base + size_t(b * D + d) * kTmemThreads, P.fx + d * nf + e0 + b);
      cp_async_commit();
    }
    auto maybe_pro = [&](int v) {
      if (v == v_pro && !next_known) {
        wt_next = __shfl_sync(kFull, wt_next, 0);
        next_known = true;
      }
      if (v == v_pro && wt_next < n_wtiles) {  // warp-uniform
        const Ptrs pn = tile_ptrs(wt_next);
        pro_issue(pn);
        cp_async_commit();
      }
    };
    if (nmax == 0) maybe_pro(0);  // v_pro == 0

    for (int v = nmax; v >= 1; --v) {
      uint32_t w[kWords];
      if constexpr (kHoist > 0) state_issue(v - 1, w);
      const bool act = v <= nh;
      // independent of the state being fetched: segment time, its powers, t = E_v u_{v+1}
      const double T = act ? HT(v) : 1.0;
      const double iT = fast_rcp(T);
      double pw[N - 1];
      segment_powers<N, R>(T, iT, pw);
      double tE[m][D];
      auto compute_tE = [&]() {
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
          for (int a = 0; a < m; ++a) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < m; ++b) s = fma(pw[a + b + 2] * G::at(1 + a, h + 1 + b), ed[1 + b][d], s);
            tE[a][d] = s;
          }
      };
      if constexpr (kHoistE) compute_tE();
      maybe_pro(v);
      double sv[kSlots];
      if constexpr (kHoist == 0) state_issue(v - 1, w);
      state_finish(v - 1, w, sv);
      if constexpr (!kHoistE) compute_tE();
      double sd[h][D];
      if (act) {
        double xv[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xv[d] = sv[kL + m * D + d];
        if constexpr (FUSED) {
          if (tout != nullptr && valid) tout[seg(v)] = T;
        }
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
#pragma unroll
        for (int d = 0; d < D; ++d) {
          double t[m];
#pragma unroll
          for (int j = 0; j < m; ++j) {
            double s = tE[j][d];
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
    if (nmax > 0 && v_pro == 0) maybe_pro(0);  // no TMEM blocks: the spill region is dead only now
    {
      double sd[h][D];
      if constexpr (kEndInRing) {
        // everything except (possibly) the next tile's prologue group has landed
        if (wt_next < n_wtiles) cp_async_wait_group<1>(); else cp_async_wait_group<0>();
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
        sd[0][d] = x0s[size_t(d) * kTmemThreads];
#pragma unroll
        for (int b = 0; b < m; ++b) {
          if constexpr (FUSED) {
            sd[1 + b][d] = 0.0;
          } else if constexpr (kEndInRing) {
            sd[1 + b][d] = sgn(b) * base[size_t(b * D + d) * kTmemThreads];
          } else {
            sd[1 + b][d] = sgn(b) * __ldg(P.fx + d * nf + e0 + b);
          }
        }
      }
      const double T = HT(0);
      if constexpr (FUSED) {
        if (tout != nullptr && valid) tout[seg(0)] = T;
      }
      const double iT = fast_rcp(T);
      __syncwarp();
      emit_all(0, 0, T, iT, sd, ed);
    }
    wt = wt_next;
  }

  if (lane == 0) bulk_wait_all();
  if (tl.tmem_cols > 0) {
    tmem::fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem::dealloc(*holder, (uint32_t)tl.tmem_cols);
  }
}

}  // namespace mtg
