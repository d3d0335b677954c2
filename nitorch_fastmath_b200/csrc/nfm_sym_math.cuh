// nfm_sym_math.cuh -- per-matrix register math for packed symmetric matrices.
//
// Packed layout (nitorch_fastmath/sym.py:7-14): diagonal first, then the rows
// of the strict upper triangle: [a00 .. a(N-1)(N-1) | a01 a02 .. a0(N-1) | a12 ..].
//
// Orders 1..4 use the same closed forms (adjugate / determinant) as the
// reference's own implementation (_impl/sym.py:186-324) so that results track
// it term for term, including on indefinite or singular input (NaN/inf, no
// trap).  Orders 5..10 factorise: LDL^T without pivoting (default; SPD or
// strongly regular input) or LU with partial pivoting on the expanded matrix
// (the reference's semantics for N > 4, _impl/sym.py:392-396).
// All loops are fully unrolled over compile-time N so every array lives in
// registers; pivoting is done with predicated swaps, never dynamic indexing.
#pragma once

#include "nfm_common.cuh"

namespace nfm {

__host__ __device__ constexpr int packed_len(int n) { return n * (n + 1) / 2; }

// position of a_ij in the packed record
__host__ __device__ constexpr int pidx(int n, int i, int j) {
  return i == j ? i : (i < j ? n + i * n - (i * (i + 1)) / 2 + (j - i - 1) : n + j * n - (j * (j + 1)) / 2 + (i - j - 1));
}

template <typename T>
__device__ __forceinline__ T sq(T x) { return x * x; }

template <typename T>
__device__ __forceinline__ T tabs(T x) { return x < T(0) ? -x : x; }

// ---------------------------------------------------------------------------
// y = A v      (_impl/sym.py:88-131)
// ---------------------------------------------------------------------------
template <typename T, int N>
__device__ __forceinline__ void sym_matvec_reg(const T (&a)[packed_len(N)], const T (&v)[N], T (&y)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = a[i] * v[i];
  if constexpr (N <= 4) {
    // unrolled variants: per output row, columns in increasing order (:88-119)
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j)
        if (j != i) y[i] += a[pidx(N, i, j)] * v[j];
  } else {
    // generic variant walks the strict upper triangle row by row (:123-131)
    static_for<0, N>([&](auto I) {
      constexpr int i = I;
      static_for<i + 1, N>([&](auto J) {
        constexpr int j = J;
        y[i] += a[pidx(N, i, j)] * v[j];
        y[j] += a[pidx(N, i, j)] * v[i];
      });
    });
  }
}

// ---------------------------------------------------------------------------
// closed forms, N = 1..4.  d = diagonal, u = strict upper triangle (packed order)
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T sym_det2(const T* d, const T* u) { return d[0] * d[1] - sq(u[0]); }

template <typename T>
__device__ __forceinline__ T sym_det3(const T* d, const T* u) {
  // u = [a01 a02 a12]   (_impl/sym.py:203-209)
  return d[0] * d[1] * d[2] + T(2) * (u[0] * u[1] * u[2]) - (d[0] * sq(u[2]) + d[2] * sq(u[0]) + d[1] * sq(u[1]));
}

template <typename T>
__device__ __forceinline__ T sym_det4(const T* d, const T* u) {
  // u = [a01 a02 a03 a12 a13 a23]   (_impl/sym.py:229-248)
  const T a = u[0], b = u[1], c = u[2], e = u[3], f = u[4], g = u[5];
  return d[0] * d[1] * d[2] * d[3] + (sq(a * g) + sq(b * f) + sq(c * e)) -
         T(2) * (a * b * f * g + a * c * e * g + b * c * e * f) +
         T(2) * (d[0] * e * f * g + d[1] * b * c * g + d[2] * a * c * f + d[3] * a * b * e) -
         (d[0] * d[1] * sq(g) + d[0] * d[2] * sq(f) + d[0] * d[3] * sq(e) + d[1] * d[2] * sq(c) +
          d[1] * d[3] * sq(b) + d[2] * d[3] * sq(a));
}

// adjugate (cofactor matrix, symmetric) in packed order, and determinant
template <typename T, int N>
__device__ __forceinline__ T sym_adjugate(const T (&m)[packed_len(N)], T (&adj)[packed_len(N)]) {
  static_assert(N >= 1 && N <= 4, "closed forms cover N = 1..4");
  const T* d = m;
  const T* u = m + N;
  if constexpr (N == 1) {
    adj[0] = T(1);
    return d[0];
  } else if constexpr (N == 2) {
    adj[0] = d[1];
    adj[1] = d[0];
    adj[2] = -u[0];
    return sym_det2(d, u);
  } else if constexpr (N == 3) {
    // _impl/sym.py:216-224
    adj[0] = d[1] * d[2] - sq(u[2]);
    adj[1] = d[0] * d[2] - sq(u[1]);
    adj[2] = d[0] * d[1] - sq(u[0]);
    adj[3] = u[1] * u[2] - d[2] * u[0];
    adj[4] = u[0] * u[2] - d[1] * u[1];
    adj[5] = u[0] * u[1] - d[0] * u[2];
    return sym_det3(d, u);
  } else {
    // Same adjugate / determinant as _impl/sym.py:229-322, evaluated through the
    // twelve 2x2 minors of rows (0,1) and (2,3) (Laplace expansion by
    // complementary minors): ~100 flops instead of the ~250 of the expanded
    // polynomials, which would make the 4x4 case issue-bound on B200.
    const T a = u[0], b = u[1], c = u[2], e = u[3], f = u[4], g = u[5];
    const T s0 = d[0] * d[1] - a * a, s1 = d[0] * e - a * b, s2 = d[0] * f - a * c;
    const T s3 = a * e - d[1] * b, s4 = a * f - d[1] * c, s5 = b * f - e * c;
    const T c5 = d[2] * d[3] - g * g, c4 = e * d[3] - f * g, c3 = e * g - f * d[2];
    const T c2 = b * d[3] - c * g, c1 = b * g - c * d[2], c0 = s5;
    adj[0] = d[1] * c5 - e * c4 + f * c3;     // 00
    adj[1] = d[0] * c5 - b * c2 + c * c1;     // 11
    adj[2] = c * s4 - f * s2 + d[3] * s0;     // 22
    adj[3] = b * s3 - e * s1 + d[2] * s0;     // 33
    adj[4] = -a * c5 + b * c4 - c * c3;       // 01
    adj[5] = f * s5 - g * s4 + d[3] * s3;     // 02
    adj[6] = -e * s5 + d[2] * s4 - g * s3;    // 03
    adj[7] = -c * s5 + g * s2 - d[3] * s1;    // 12
    adj[8] = b * s5 - d[2] * s2 + g * s1;     // 13
    adj[9] = -b * s4 + e * s2 - g * s0;       // 23
    return s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
  }
}

template <typename T, int N>
__device__ __forceinline__ T sym_det_closed(const T (&m)[packed_len(N)]) {
  if constexpr (N == 1) return m[0];
  else if constexpr (N == 2) return sym_det2(m, m + N);
  else if constexpr (N == 3) return sym_det3(m, m + N);
  else return sym_det4(m, m + N);
}

// x = A^-1 v via adjugate / det   (_impl/sym.py:193-324, :384-391)
template <typename T, int N>
__device__ __forceinline__ void sym_solve_closed(const T (&m)[packed_len(N)], const T (&v)[N], T (&x)[N]) {
  if constexpr (N == 1) {
    x[0] = v[0] / m[0];
  } else {
    T adj[packed_len(N)];
    const T det = sym_adjugate<T, N>(m, adj);
    if constexpr (N == 4 || sizeof(T) == 8) {
      // reference accumulates diagonal term first, then the others in column
      // order (:291-322) and divides every entry; one reciprocal here (<= 1 ulp
      // apart): N = 4 and fp64 would otherwise be issue-bound on divisions
      const T rdet = T(1) / det;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        T s = adj[i] * v[i];
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (j != i) s += adj[pidx(N, i, j)] * v[j];
        x[i] = s * rdet;
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        T s = adj[pidx(N, i, 0)] * v[0];
#pragma unroll
        for (int j = 1; j < N; ++j) s += adj[pidx(N, i, j)] * v[j];
        x[i] = s / det;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// LDL^T, no pivoting, upper-triangle storage w[i][j] (j >= i).
// After factor(): w[k][k] = d_k ; w[k][j] = l_jk  (j > k).
// ---------------------------------------------------------------------------
template <typename T, int N>
struct LDL {
  T w[N][N];
  T rd[N];  // 1 / d_k

  __device__ __forceinline__ void load_packed(const T (&m)[packed_len(N)]) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) w[i][j] = m[pidx(N, i, j)];
  }

  __device__ __forceinline__ void factor() {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      rd[k] = T(1) / w[k][k];
#pragma unroll
      for (int i = k + 1; i < N; ++i) {
        const T lik = w[k][i] * rd[k];
#pragma unroll
        for (int j = i; j < N; ++j) w[i][j] -= lik * w[k][j];
        w[k][i] = lik;
      }
    }
  }

  // Same factorisation, but reports whether every 1x1 pivot was acceptable:
  // |d_k| >= kPivotTol * max_i |a_ki| of the current Schur complement (the
  // Bunch-Kaufman style test; with 0.1 a multiplier is at most 10, so growth is
  // bounded but looser than partial pivoting -- a tighter tolerance would send
  // legitimate SPD matrices with a wide diagonal range to the slow redo).
  // Non-finite pivots are NOT caught (fmin drops NaN): such input yields
  // non-finite output on either path, as in the reference.  SPD
  // matrices always pass; a symmetric indefinite matrix that fails is re-solved
  // with pivoted LU by the caller, which is what the reference does for every
  // matrix of order > 4 (_impl/sym.py:392-396).
  static constexpr float kPivotTol = 0.1f;
  __device__ __forceinline__ bool factor_checked() {
    if constexpr (sizeof(T) == 8) {
      // fp64: do the magnitude bookkeeping on the high words as integers (the
      // ordering of |x| is the ordering of its exponent/top-mantissa bits; the
      // tolerance 2^-3 is a subtraction in the exponent field): full-rate
      // integer min/max instead of ~N*N/2 half-rate FP64 compares.
      int worst = 0x7fffffff;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        int big = 0;
#pragma unroll
        for (int i = k + 1; i < N; ++i) big = max(big, __double2hiint(w[k][i]) & 0x7fffffff);
        worst = min(worst, (__double2hiint(w[k][k]) & 0x7fffffff) - big + (3 << 20));
        rd[k] = T(1) / w[k][k];
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
          const T lik = w[k][i] * rd[k];
#pragma unroll
          for (int j = i; j < N; ++j) w[i][j] -= lik * w[k][j];
          w[k][i] = lik;
        }
      }
      return worst >= 0;
    } else {
      // branch-free: one running minimum of the pivot margins, tested once at the
      // end (fmax / fabs map to one FMNMX with an |x| source modifier each)
      T worst = fabs(w[0][0]);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        T big = T(0);
#pragma unroll
        for (int i = k + 1; i < N; ++i) big = fmax(big, fabs(w[k][i]));
        worst = fmin(worst, fabs(w[k][k]) - T(kPivotTol) * big);
        rd[k] = T(1) / w[k][k];
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
          const T lik = w[k][i] * rd[k];
#pragma unroll
          for (int j = i; j < N; ++j) w[i][j] -= lik * w[k][j];
          w[k][i] = lik;
        }
      }
      return worst >= T(0);
    }
  }

  __device__ __forceinline__ void solve(const T (&v)[N], T (&x)[N]) const {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = v[i];
#pragma unroll
    for (int k = 0; k < N; ++k)  // L y = v
#pragma unroll
      for (int i = k + 1; i < N; ++i) x[i] -= w[k][i] * x[k];
#pragma unroll
    for (int k = 0; k < N; ++k) x[k] *= rd[k];
#pragma unroll
    for (int k = N - 1; k >= 0; --k)  // L^T x = z
#pragma unroll
      for (int i = k + 1; i < N; ++i) x[k] -= w[k][i] * x[i];
  }

  // A^-1 = L^-T D^-1 L^-1, packed (or only its diagonal)
  template <bool kDiagOnly>
  __device__ __forceinline__ void invert(T* __restrict__ out) {
    // g[k][i] (i < k) = (L^-1)_ki, stored in the free lower triangle of w
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int k = i + 1; k < N; ++k) {
        T s = -w[i][k];
#pragma unroll
        for (int j = i + 1; j < k; ++j) s -= w[j][k] * w[j][i];
        w[k][i] = s;
      }
    }
    // (A^-1)_ij = sum_{k >= max(i,j)} g_ki g_kj / d_k , g_kk = 1
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int j = i; j < N; ++j) {
        if (kDiagOnly && j != i) continue;
        T s = (i == j) ? rd[j] : w[j][i] * rd[j];
#pragma unroll
        for (int k = j + 1; k < N; ++k) s += (w[k][i] * rd[k]) * w[k][j];
        out[kDiagOnly ? i : pidx(N, i, j)] = s;
      }
    }
  }
};

// ---------------------------------------------------------------------------
// dense Gaussian elimination with partial pivoting on an augmented system
// [A | B], R right-hand-side columns; registers only, predicated row swaps.
// Used by: sym solve/invert with NFM_ALGO_LU, dense solve / inverse / det.
// ---------------------------------------------------------------------------
template <typename T, int N, int R>
struct GaussPP {
  T a[N][N];
  T b[N][R > 0 ? R : 1];
  T det_sign;

  // forward elimination; on return a is upper triangular (pivots on the
  // diagonal, *not* normalised) and b is transformed accordingly
  __device__ __forceinline__ void eliminate() {
    det_sign = T(1);
    static_for<0, N>([&](auto K) {
      constexpr int k = K;
      // pivot search: first row of maximal |a_ik|, i >= k  (LAPACK getrf / idamax)
      T best = tabs(a[k][k]);
      int p = k;
      static_for<k + 1, N>([&](auto I) {
        constexpr int i = I;
        const T c = tabs(a[i][k]);
        if (c > best) {
          best = c;
          p = i;
        }
      });
      if (N < kVoteFromOrder || warp_any(p != k)) {  // skipped when no matrix of the warp exchanges rows in this step
        static_for<k + 1, N>([&](auto I) {
          constexpr int i = I;
          const bool sw = (p == i);
          static_for<k, N>([&](auto J) {
            constexpr int j = J;
            const T lo = a[k][j], hi = a[i][j];
            a[k][j] = sw ? hi : lo;
            a[i][j] = sw ? lo : hi;
          });
          static_for<0, R>([&](auto C) {
            constexpr int c = C;
            const T lo = b[k][c], hi = b[i][c];
            b[k][c] = sw ? hi : lo;
            b[i][c] = sw ? lo : hi;
          });
        });
        if (p != k) det_sign = -det_sign;
      }
      const T rp = T(1) / a[k][k];
      static_for<k + 1, N>([&](auto I) {
        constexpr int i = I;
        const T f = a[i][k] * rp;
        static_for<k + 1, N>([&](auto J) {
          constexpr int j = J;
          a[i][j] -= f * a[k][j];
        });
        static_for<0, R>([&](auto C) {
          constexpr int c = C;
          b[i][c] -= f * b[k][c];
        });
      });
    });
  }

  __device__ __forceinline__ T det() const {
    T d = det_sign;
    static_for<0, N>([&](auto K) { d *= a[K][K]; });
    return d;
  }

  // back substitution: b <- U^-1 b
  __device__ __forceinline__ void back_substitute() {
    static_for_down<0, N>([&](auto K) {
      constexpr int k = K;
      const T rp = T(1) / a[k][k];
      static_for<0, R>([&](auto C) {
        constexpr int c = C;
        T s = b[k][c];
        static_for<k + 1, N>([&](auto J) {
          constexpr int j = J;
          s -= a[k][j] * b[j][c];
        });
        b[k][c] = s * rp;
      });
    });
  }
};

// x = (A)^-1 v for a packed symmetric A through pivoted LU of the expansion
template <typename T, int N>
__device__ __forceinline__ void sym_solve_lu(const T (&m)[packed_len(N)], const T (&v)[N], T (&x)[N]) {
  GaussPP<T, N, 1> g;
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = 0; j < N; ++j) g.a[i][j] = m[pidx(N, i, j)];
    g.b[i][0] = v[i];
  }
  g.eliminate();
  g.back_substitute();
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = g.b[i][0];
}

}  // namespace nfm
