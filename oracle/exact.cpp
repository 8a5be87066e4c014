// exact.cpp -- binary128 (__float128, 113-bit significand, ~34 digits) solve of the linear
// min-derivative problem: the "fast exact" checker for parity at scale.
//
// TEST INFRASTRUCTURE ONLY (tests/, tools/parity_report.py): never linked or loaded by the product.
//
// Independent of oracle.cpp (reference operation order in fp64) and of the kernels' scaled-table
// formulation: builds A(T) and Q(T) literally (reference
// impl/polynomial_optimization_linear_impl.h:111-121 setupMappingMatrix, :567-583
// computeQuadraticCostJacobian), inverts A by Gauss-Jordan with partial pivoting, forms
// H = A^-T Q A^-1 and R = C^T H C (:307-336), solves R_pp d_p = -R_pf d_f (:360-375; here a banded
// LDL^T -- R_pp is SPD) and back-substitutes p = A^-1 C d (:262-283), all in binary128, rounding to
// fp64 once at the very end.  The reference-order fp64 arithmetic loses ~5-6 digits on these systems
// (oracle.cpp is ~1e-11 from a 60-digit solve), so binary128 is exact to well below one fp64 ulp of the
// result; tests/test_oracle.py pins this file against oracle/truth.py (mpmath, 60 digits).
//
// Why it exists: truth.py takes seconds per trajectory; this takes ~2-15 ms per trajectory and thread, so the GPU parity suite can
// assert GPU-vs-exact on >= 8192 fixture trajectories per BASELINE configuration instead of 3.
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

namespace {

typedef __float128 Q;

inline Q qabs(Q x) { return x < 0 ? -x : x; }

double base_coeff(int k, int j) {  // B(k,j) = j!/(j-k)!  (polynomial.cpp:145-160)
  if (j < k) return 0.0;
  double v = 1.0;
  for (int q = 0; q < k; ++q) v *= double(j - q);
  return v;
}

// Gauss-Jordan inverse with partial pivoting, n <= 12
bool invert(int n, const Q* A, Q* Ai) {
  Q M[12][24];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      M[i][j] = A[i * n + j];
      M[i][n + j] = (i == j) ? Q(1) : Q(0);
    }
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int i = c + 1; i < n; ++i)
      if (qabs(M[i][c]) > qabs(M[piv][c])) piv = i;
    if (M[piv][c] == 0) return false;
    if (piv != c)
      for (int j = 0; j < 2 * n; ++j) std::swap(M[piv][j], M[c][j]);
    const Q ip = Q(1) / M[c][c];
    for (int j = 0; j < 2 * n; ++j) M[c][j] *= ip;
    for (int i = 0; i < n; ++i) {
      if (i == c) continue;
      const Q f = M[i][c];
      if (f == 0) continue;
      for (int j = 0; j < 2 * n; ++j)
        if (M[c][j] != 0) M[i][j] -= f * M[c][j];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Ai[i * n + j] = M[i][n + j];
  return true;
}

struct Layout {
  int nf = 0, np = 0, bw = 0;
  std::vector<int> slot;  // [K*N]
};

// constraint reordering (linear.h:287-295, linear_impl.h:181-260): rank inside the sorted fixed / free sets
Layout make_layout(int N, int K, const uint8_t* mask) {
  const int h = N / 2;
  Layout L;
  std::vector<int> col(size_t(K + 1) * h);
  for (size_t i = 0; i < col.size(); ++i) (mask[i] ? L.nf : L.np)++;
  int cf = 0, cp = 0;
  for (size_t i = 0; i < col.size(); ++i) col[i] = mask[i] ? cf++ : L.nf + cp++;
  L.slot.resize(size_t(K) * N);
  for (int i = 0; i < K; ++i) {
    int lo = 1 << 30, hi = -1;
    for (int s = 0; s < N; ++s) {
      const int v = s < h ? i : i + 1, k = s < h ? s : s - h;
      const int c = col[size_t(v) * h + k];
      L.slot[size_t(i) * N + s] = c;
      if (c >= L.nf) {
        lo = std::min(lo, c);
        hi = std::max(hi, c);
      }
    }
    if (hi >= lo) L.bw = std::max(L.bw, hi - lo);
  }
  return L;
}

// one trajectory; returns false on a singular A or a non-positive pivot
bool solve_one(int N, int r, int K, int D, const Layout& L, const double* times, const double* dfix, double* coeffs,
               double* dfree, double* cost) {
  const int h = N / 2, nf = L.nf, np = L.np, bw = L.bw, bw1 = bw + 1;
  std::vector<Q> Ainv(size_t(K) * N * N);
  std::vector<Q> Hall(cost ? size_t(K) * N * N : 0);
  std::vector<Q> band(size_t(np) * bw1, Q(0));  // band[i*bw1 + (i-j)] = R_pp[i][j], j <= i
  std::vector<Q> rhs(size_t(np) * D, Q(0));
  std::vector<Q> d_all(size_t(nf + np) * D);
  for (int d = 0; d < D; ++d)
    for (int c = 0; c < nf; ++c) d_all[size_t(c) * D + d] = Q(dfix[size_t(d) * nf + c]);
  for (int i = 0; i < K; ++i) {
    const Q t = Q(times[i]);
    Q A[144], Qm[144], tp[24];
    tp[0] = 1;
    for (int k = 1; k < 24; ++k) tp[k] = tp[k - 1] * t;
    for (int k = 0; k < N * N; ++k) A[k] = Qm[k] = 0;
    for (int k = 0; k < h; ++k) {
      A[k * N + k] = Q(base_coeff(k, k));
      for (int j = k; j < N; ++j) A[(h + k) * N + j] = Q(base_coeff(k, j)) * tp[j - k];
    }
    for (int a = r; a < N; ++a)
      for (int b = r; b < N; ++b) {
        const int e = a + b - 2 * r + 1;
        Qm[a * N + b] = Q(2.0 * base_coeff(r, a) * base_coeff(r, b)) * tp[e] / Q(e);
      }
    Q* Ai = &Ainv[size_t(i) * N * N];
    if (!invert(N, A, Ai)) return false;
    Q QA[144], H[144];
    for (int a = 0; a < N; ++a)
      for (int b = 0; b < N; ++b) {
        Q s = 0;
        if (a >= r)  // rows a < r of Q are zero
          for (int k = r; k < N; ++k)
            if (Ai[k * N + b] != 0) s += Qm[a * N + k] * Ai[k * N + b];
        QA[a * N + b] = s;
      }
    for (int a = 0; a < N; ++a)
      for (int b = 0; b < N; ++b) {
        Q s = 0;
        for (int k = r; k < N; ++k)  // rows k < r of Q A^-1 are zero
          if (Ai[k * N + a] != 0) s += Ai[k * N + a] * QA[k * N + b];
        H[a * N + b] = s;
        if (cost) Hall[(size_t(i) * N + a) * N + b] = s;
      }
    const int* sl = &L.slot[size_t(i) * N];
    for (int a = 0; a < N; ++a) {
      const int ca = sl[a] - nf;
      if (ca < 0) continue;
      for (int b = 0; b < N; ++b) {
        const int cb = sl[b];
        if (cb >= nf) {
          const int j = cb - nf;
          if (j <= ca) band[size_t(ca) * bw1 + (ca - j)] += H[a * N + b];
        } else {
          for (int d = 0; d < D; ++d) rhs[size_t(ca) * D + d] -= H[a * N + b] * d_all[size_t(cb) * D + d];
        }
      }
    }
  }
  // banded LDL^T: band(i,0) = d_i, band(i,i-j) = l_ij
  for (int i = 0; i < np; ++i) {
    const int j0 = std::max(0, i - bw);
    for (int j = j0; j <= i; ++j) {
      Q s = band[size_t(i) * bw1 + (i - j)];
      for (int k = std::max(j0, j - bw); k < j; ++k)
        s -= band[size_t(i) * bw1 + (i - k)] * band[size_t(k) * bw1] * band[size_t(j) * bw1 + (j - k)];
      if (j < i) {
        band[size_t(i) * bw1 + (i - j)] = s / band[size_t(j) * bw1];
      } else {
        if (!(s > 0)) return false;
        band[size_t(i) * bw1] = s;
      }
    }
  }
  for (int d = 0; d < D; ++d) {
    for (int i = 0; i < np; ++i) {
      Q s = rhs[size_t(i) * D + d];
      for (int k = std::max(0, i - bw); k < i; ++k) s -= band[size_t(i) * bw1 + (i - k)] * rhs[size_t(k) * D + d];
      rhs[size_t(i) * D + d] = s;
    }
    for (int i = 0; i < np; ++i) rhs[size_t(i) * D + d] /= band[size_t(i) * bw1];
    for (int i = np - 1; i >= 0; --i) {
      Q s = rhs[size_t(i) * D + d];
      for (int k = i + 1; k <= std::min(np - 1, i + bw); ++k) s -= band[size_t(k) * bw1 + (k - i)] * rhs[size_t(k) * D + d];
      rhs[size_t(i) * D + d] = s;
    }
    for (int i = 0; i < np; ++i) {
      d_all[size_t(nf + i) * D + d] = rhs[size_t(i) * D + d];
      if (dfree) dfree[size_t(d) * np + i] = double(rhs[size_t(i) * D + d]);
    }
  }
  if (cost) {  // computeCost() (linear_impl.h:123-140) = 0.5 sum_seg sum_dim d^T H d, unrounded coefficients
    Q J = 0;
    for (int i = 0; i < K; ++i) {
      const int* sl = &L.slot[size_t(i) * N];
      for (int d = 0; d < D; ++d)
        for (int a = 0; a < N; ++a) {
          Q s = 0;
          for (int b = 0; b < N; ++b) s += Hall[(size_t(i) * N + a) * N + b] * d_all[size_t(sl[b]) * D + d];
          J += s * d_all[size_t(sl[a]) * D + d];
        }
    }
    *cost = double(J / 2);
  }
  for (int i = 0; i < K; ++i) {
    const Q* Ai = &Ainv[size_t(i) * N * N];
    const int* sl = &L.slot[size_t(i) * N];
    for (int d = 0; d < D; ++d)
      for (int j = 0; j < N; ++j) {
        Q s = 0;
        for (int c = 0; c < N; ++c) s += Ai[j * N + c] * d_all[size_t(sl[c]) * D + d];
        coeffs[(size_t(i) * D + d) * N + j] = double(s);
      }
  }
  return true;
}

}  // namespace

extern "C" {

// mask [(K+1)*(N/2)] (nullptr = createRandomVertices topology), times [B][K], dfix [B][D][n_fixed] ->
// coeffs [B][K][D][N], dfree [B][D][n_free] (nullable), cost [B] (nullable; computeCost()).  Returns the number of failed trajectories, or -1
// on a bad argument.
int exact_solve_batch(int N, int r, int K, int D, const uint8_t* mask, int64_t B, const double* times,
                      const double* dfix, double* coeffs, double* dfree, double* cost, int n_threads) {
  if (N < 2 || N > 12 || (N & 1) || r < 0 || r > N / 2 - 1 || K < 1 || D < 1 || B < 0) return -1;
  const int h = N / 2;
  std::vector<uint8_t> m(size_t(K + 1) * h, 0);
  if (mask) {
    std::memcpy(m.data(), mask, m.size());
  } else {
    for (int v = 0; v <= K; ++v) {
      m[size_t(v) * h] = 1;
      if (v == 0 || v == K)
        for (int k = 1; k < h; ++k) m[size_t(v) * h + k] = 1;
    }
  }
  const Layout L = make_layout(N, K, m.data());
  std::atomic<int64_t> next(0), failed(0);
  auto worker = [&]() {
    for (;;) {
      const int64_t b = next.fetch_add(1);
      if (b >= B) break;
      const bool ok = solve_one(N, r, K, D, L, times + b * K, dfix + b * size_t(D) * L.nf,
                                coeffs + b * size_t(K) * D * N, dfree ? dfree + b * size_t(D) * L.np : nullptr,
                                cost ? cost + b : nullptr);
      if (!ok) failed.fetch_add(1);
    }
  };
  n_threads = std::max(1, n_threads);
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  return int(failed.load());
}

}  // extern "C"
