// nfm_extra.cu -- the remaining public names of nitorch_fastmath/sym.py:29
// (SURVEY.md section 8f "next" rows): sym_det, sym_to_full, sym_outer.
#include "nfm_dense_ops.cuh"
#include "nfm_impl.cuh"
#include "nfm_pipeline.cuh"
#include "nfm_sym_ops.cuh"

namespace nfm {

// det of a packed symmetric matrix (_impl/sym.py:401-452): closed form for
// N <= 4, pivoted LU of the expansion above (the reference calls torch.det).
template <typename T, int N>
struct SymDetOp {
  using scalar = T;
  static constexpr int kLen0 = packed_len(N), kLen1 = 1, kLen2 = 1, kUse = 1, kOut = 1;
  static constexpr bool kHeavy = N > 4;
  __device__ static __forceinline__ void apply(const T (&m)[kLen0], const T (&)[1], const T (&)[1], int, int, T (&out)[1]) {
    if constexpr (N <= 4) {
      out[0] = sym_det_closed<T, N>(m);
    } else {
      GaussPP<T, N, 0> g;
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) g.a[i][j] = m[pidx(N, i, j)];
      g.eliminate();
      out[0] = g.det();
    }
  }
};

// packed -> dense (_impl/sym.py:16-60)
template <typename T, int N>
struct SymToFullOp {
  using scalar = T;
  static constexpr int kLen0 = packed_len(N), kLen1 = 1, kLen2 = 1, kUse = 1, kOut = N * N;
  static constexpr bool kHeavy = false;
  __device__ static __forceinline__ void apply(const T (&m)[kLen0], const T (&)[1], const T (&)[1], int, int, T (&out)[kOut]) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) out[i * N + j] = m[pidx(N, i, j)];
  }
};

// x x^T packed (_impl/sym.py:496-528)
template <typename T, int N>
struct SymOuterOp {
  using scalar = T;
  static constexpr int kLen0 = N, kLen1 = 1, kLen2 = 1, kUse = 1, kOut = packed_len(N);
  static constexpr bool kHeavy = false;
  __device__ static __forceinline__ void apply(const T (&x)[N], const T (&)[1], const T (&)[1], int, int, T (&out)[kOut]) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) out[pidx(N, i, j)] = x[i] * x[j];
  }
};

// J^T H J (mode 0) or J H J^T (mode 1, K == D), packed output (_impl/sym.py:532-670)
template <typename T, int K, int D, int MODE>
struct SymMatmulOp {
  using scalar = T;
  static constexpr int kHN = MODE == 0 ? K : D;   // order of H
  static constexpr int kON = MODE == 0 ? D : K;   // order of the result
  static constexpr int kLen0 = K * D, kLen1 = packed_len(kHN), kLen2 = 1, kUse = 3, kOut = packed_len(kON);
  static constexpr bool kHeavy = false;
  __device__ static __forceinline__ void apply(const T (&j)[kLen0], const T (&h)[kLen1], const T (&)[1], int, int, T (&out)[kOut]) {
    // G = J as (kON x kHN) "rows = output index": mode 0 uses J^T, mode 1 uses J
    T hj[kHN][kON];  // H * G^T
#pragma unroll
    for (int a = 0; a < kHN; ++a)
#pragma unroll
      for (int o = 0; o < kON; ++o) {
        T s = T(0);
#pragma unroll
        for (int b = 0; b < kHN; ++b) {
          const T g = MODE == 0 ? j[b * D + o] : j[o * D + b];
          s += h[pidx(kHN, a, b)] * g;
        }
        hj[a][o] = s;
      }
#pragma unroll
    for (int o = 0; o < kON; ++o)
#pragma unroll
      for (int q = o; q < kON; ++q) {
        T s = T(0);
#pragma unroll
        for (int a = 0; a < kHN; ++a) {
          const T g = MODE == 0 ? j[a * D + o] : j[o * D + a];
          s += g * hj[a][q];
        }
        out[pidx(kON, o, q)] = s;
      }
  }
};

template <typename T, int K, int D>
static int matmul_run(int mode, const KParams& p, cudaStream_t s) {
  if (mode == 0) return run_op<SymMatmulOp<T, K, D, 0>>(p, s);
  if constexpr (K == D) return run_op<SymMatmulOp<T, K, D, 1>>(p, s);
  else return NFM_E_UNSUPPORTED;
}

template <typename T, int K>
static int matmul_d(int d, int mode, const KParams& p, cudaStream_t s) {
  switch (d) {
    case 1: return matmul_run<T, K, 1>(mode, p, s);
    case 2: return matmul_run<T, K, 2>(mode, p, s);
    case 3: return matmul_run<T, K, 3>(mode, p, s);
    case 4: return matmul_run<T, K, 4>(mode, p, s);
  }
  return NFM_E_UNSUPPORTED;
}

template <typename T>
static int matmul_kd(int k, int d, int mode, const KParams& p, cudaStream_t s) {
  switch (k) {
    case 1: return matmul_d<T, 1>(d, mode, p, s);
    case 2: return matmul_d<T, 2>(d, mode, p, s);
    case 3: return matmul_d<T, 3>(d, mode, p, s);
    case 4: return matmul_d<T, 4>(d, mode, p, s);
  }
  return NFM_E_UNSUPPORTED;
}

template <typename T, int ALGO> struct SolveUpdBind { template <int N> using Op = SymSolveUpdateOp<T, N, ALGO>; };

template <typename T> struct SDetBind { template <int N> using Op = SymDetOp<T, N>; };
template <typename T> struct SFullBind { template <int N> using Op = SymToFullOp<T, N>; };
template <typename T> struct SOuterBind { template <int N> using Op = SymOuterOp<T, N>; };

template <template <typename> class Bind>
static int unary_entry(int dtype, int n, i64 batch, const void* in, i64 in_stride, void* out, i64 out_stride, void* stream) {
  if (dtype != NFM_F32 && dtype != NFM_F64) { set_error("dtype must be NFM_F32 or NFM_F64"); return NFM_E_UNSUPPORTED; }
  if (n < 1 || n > NFM_MAX_N) { set_error("matrix order must be in 1..10"); return NFM_E_UNSUPPORTED; }
  if (batch < 0 || in == nullptr || out == nullptr || in_stride < 0 || out_stride < 0) { set_error("bad argument"); return NFM_E_BADARG; }
  KParams p{};
  p.in[0].ptr = in;
  p.in[0].stride = in_stride;
  p.present = 1;
  p.out = out;
  p.out_stride = out_stride;
  p.batch = batch;
  auto s = static_cast<cudaStream_t>(stream);
  if (dtype == NFM_F32) return DispatchN<Bind<float>::template Op, 1, NFM_MAX_N>::run(n, p, s);
  return DispatchN<Bind<double>::template Op, 1, NFM_MAX_N>::run(n, p, s);
}

}  // namespace nfm

using namespace nfm;

extern "C" {

int nfm_sym_det(int dtype, int n, int64_t batch, const void* mat, int64_t mat_stride, void* out, int64_t out_stride, void* stream) {
  return unary_entry<SDetBind>(dtype, n, batch, mat, mat_stride, out, out_stride, stream);
}
int nfm_sym_to_full(int dtype, int n, int64_t batch, const void* mat, int64_t mat_stride, void* out, int64_t out_stride, void* stream) {
  return unary_entry<SFullBind>(dtype, n, batch, mat, mat_stride, out, out_stride, stream);
}
int nfm_sym_outer(int dtype, int n, int64_t batch, const void* vec, int64_t vec_stride, void* out, int64_t out_stride, void* stream) {
  return unary_entry<SOuterBind>(dtype, n, batch, vec, vec_stride, out, out_stride, stream);
}

int nfm_sym_matmul(int dtype, int k, int d, int mode, int64_t batch, const void* jac, int64_t jac_stride, const void* hess,
                   int64_t hess_stride, void* out, int64_t out_stride, void* stream) {
  if (dtype != NFM_F32 && dtype != NFM_F64) { set_error("dtype must be NFM_F32 or NFM_F64"); return NFM_E_UNSUPPORTED; }
  if (k < 1 || k > NFM_MAX_N || d < 1 || d > NFM_MAX_N || (mode != 0 && mode != 1) || (mode == 1 && k != d)) {
    set_error("sym_matmul: 1 <= k, d <= 10; mode 1 needs k == d");
    return NFM_E_UNSUPPORTED;
  }
  if (batch < 0 || !jac || !hess || !out || jac_stride < 0 || hess_stride < 0 || out_stride < 0) { set_error("bad argument"); return NFM_E_BADARG; }
  if (k > 4 || d > 4) {  // run-time-sized kernel above the templated 4 x 4
    auto st = static_cast<cudaStream_t>(stream);
    const int rc = dtype == NFM_F32 ? sym_matmul_rt<float>(k, d, mode, batch, jac, jac_stride, hess, hess_stride, out, out_stride, st)
                                    : sym_matmul_rt<double>(k, d, mode, batch, jac, jac_stride, hess, hess_stride, out, out_stride, st);
    if (rc) set_error("sym_matmul kernel launch failed: %s", cudaGetErrorString(cudaError_t(rc)));
    return rc;
  }
  KParams p{};
  p.in[0].ptr = jac;
  p.in[0].stride = jac_stride;
  p.in[1].ptr = hess;
  p.in[1].stride = hess_stride;
  p.present = 3;
  p.out = out;
  p.out_stride = out_stride;
  p.batch = batch;
  auto s = static_cast<cudaStream_t>(stream);
  return dtype == NFM_F32 ? matmul_kd<float>(k, d, mode, p, s) : matmul_kd<double>(k, d, mode, p, s);
}

int nfm_sym_solve_update(int dtype, int n, int algo, int64_t batch, const void* mat, int64_t mat_stride, const void* vec,
                         int64_t vec_stride, const void* x, int64_t x_stride, double lam, double alpha, void* out,
                         int64_t out_stride, void* stream) {
  if (dtype != NFM_F32 && dtype != NFM_F64) { set_error("dtype must be NFM_F32 or NFM_F64"); return NFM_E_UNSUPPORTED; }
  if (n < 1 || n > NFM_MAX_N) { set_error("matrix order must be in 1..10"); return NFM_E_UNSUPPORTED; }
  if (algo != NFM_ALGO_AUTO && algo != NFM_ALGO_LDL) { set_error("sym_solve_update: algo must be AUTO or LDL"); return NFM_E_UNSUPPORTED; }
  if (batch < 0 || !mat || !vec || !x || !out || mat_stride < 0 || vec_stride < 0 || x_stride < 0 || out_stride < 0) {
    set_error("bad argument");
    return NFM_E_BADARG;
  }
  KParams p{};
  p.in[0].ptr = mat;
  p.in[0].stride = mat_stride;
  p.in[1].ptr = vec;
  p.in[1].stride = vec_stride;
  p.in[2].ptr = x;
  p.in[2].stride = x_stride;
  p.present = 7;
  p.out = out;
  p.out_stride = out_stride;
  p.batch = batch;
  p.scal0 = lam;
  p.scal1 = alpha;
  auto s = static_cast<cudaStream_t>(stream);
  if (algo == NFM_ALGO_LDL) {
    return dtype == NFM_F32 ? DispatchN<SolveUpdBind<float, NFM_ALGO_LDL>::template Op, 1, NFM_MAX_N>::run(n, p, s)
                            : DispatchN<SolveUpdBind<double, NFM_ALGO_LDL>::template Op, 1, NFM_MAX_N>::run(n, p, s);
  }
  return dtype == NFM_F32 ? DispatchN<SolveUpdBind<float, NFM_ALGO_AUTO>::template Op, 1, NFM_MAX_N>::run(n, p, s)
                          : DispatchN<SolveUpdBind<double, NFM_ALGO_AUTO>::template Op, 1, NFM_MAX_N>::run(n, p, s);
}

}  // extern "C"
