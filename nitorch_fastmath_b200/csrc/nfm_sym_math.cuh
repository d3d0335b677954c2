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
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i + 1; j < N; ++j) {
        y[i] += a[pidx(N, i, j)] * v[j];
        y[j] += a[pidx(N, i, j)] * v[i];
      }
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
    // _impl/sym.py:253-322
    const T a = u[0], b = u[1], c = u[2], e = u[3], f = u[4], g = u[5];
    adj[0] = d[1] * d[2] * d[3] - d[1] * sq(g) - d[2] * sq(f) - d[3] * sq(e) + T(2) * e * f * g;
    adj[1] = d[0] * d[2] * d[3] - d[0] * sq(g) - d[2] * sq(c) - d[3] * sq(b) + T(2) * b * c * g;
    adj[2] = d[0] * d[1] * d[3] - d[0] * sq(f) - d[1] * sq(c) - d[3] * sq(a) + T(2) * a * c * f;
    adj[3] = d[0] * d[1] * d[2] - d[0] * sq(e) - d[1] * sq(b) - d[2] * sq(a) + T(2) * a * b * e;
    adj[4] = -d[2] * d[3] * a + d[2] * c * f + d[3] * b * e + a * sq(g) - b * f * g - c * e * g;  // 01
    adj[5] = -d[1] * d[3] * b + d[1] * c * g + d[3] * a * e + b * sq(f) - a * f * g - c * e * f;  // 02
    adj[6] = -d[1] * d[2] * c + d[1] * b * g + d[2] * a * f + c * sq(e) - a * e * g - b * e * f;  // 03
    adj[7] = -d[0] * d[3] * e + d[0] * f * g + d[3] * a * b + e * sq(c) - a * c * g - b * c * f;  // 12
    adj[8] = -d[0] * d[2] * f + d[0] * e * g + d[2] * a * c + f * sq(b) - a * b * g - b * c * e;  // 13
    adj[9] = -d[0] * d[1] * g + d[0] * f * e + d[1] * b * c + g * sq(a) - a * b * f - a * c * e;  // 23
    return sym_det4(d, u);
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
    if constexpr (N == 4) {
      // reference accumulates diagonal term first, then the others in column order (:291-322)
#pragma unroll
      for (int i = 0; i < N; ++i) {
        T s = adj[i] * v[i];
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (j != i) s += adj[pidx(N, i, j)] * v[j];
        x[i] = s / det;
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
template <typename T>
__device__ __forceinline__ T tabs(T x) { return x < T(0) ? -x : x; }

template <typename T, int N, int R>
struct GaussPP {
  T a[N][N];
  T b[N][R > 0 ? R : 1];
  T det_sign;

  // forward elimination; on return a is upper triangular (pivots on the
  // diagonal, *not* normalised) and b is transformed accordingly
  __device__ __forceinline__ void eliminate() {
    det_sign = T(1);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      {
        // pivot search: first row of maximal |a_ik|, i >= k  (LAPACK getrf / idamax)
        T best = tabs(a[k][k]);
        int p = k;
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
          const T c = tabs(a[i][k]);
          if (c > best) {
            best = c;
            p = i;
          }
        }
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
          const bool sw = (p == i);
#pragma unroll
          for (int j = k; j < N; ++j) {
            const T lo = a[k][j], hi = a[i][j];
            a[k][j] = sw ? hi : lo;
            a[i][j] = sw ? lo : hi;
          }
#pragma unroll
          for (int c = 0; c < R; ++c) {
            const T lo = b[k][c], hi = b[i][c];
            b[k][c] = sw ? hi : lo;
            b[i][c] = sw ? lo : hi;
          }
        }
        if (p != k) det_sign = -det_sign;
      }
      const T rp = T(1) / a[k][k];
#pragma unroll
      for (int i = k + 1; i < N; ++i) {
        const T f = a[i][k] * rp;
#pragma unroll
        for (int j = k + 1; j < N; ++j) a[i][j] -= f * a[k][j];
#pragma unroll
        for (int c = 0; c < R; ++c) b[i][c] -= f * b[k][c];
      }
    }
  }

  __device__ __forceinline__ T det() const {
    T d = det_sign;
#pragma unroll
    for (int k = 0; k < N; ++k) d *= a[k][k];
    return d;
  }

  // back substitution: b <- U^-1 b
  __device__ __forceinline__ void back_substitute() {
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
      const T rp = T(1) / a[k][k];
#pragma unroll
      for (int c = 0; c < R; ++c) {
        T s = b[k][c];
#pragma unroll
        for (int j = k + 1; j < N; ++j) s -= a[k][j] * b[j][c];
        b[k][c] = s * rp;
      }
    }
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
