// nfm_fused.cu -- the Gauss-Newton neighbours of the solve,
// compiled once per scalar type (-DNFM_SCALAR) and part (-DNFM_PART):
//   part 0: sym_matmul          J^T H J -> packed           (register kernels for k, d <= 6 and for k <= 10, d <= 3)
//   part 1: sym_matmul_solve    (J^T H J + diag(d))^-1 g    fused: the packed Hessian stays in registers
//   part 2: sym_solve_update    x - alpha (A + lam I + diag(d))^-1 v
#include "nfm_impl.cuh"
#include "nfm_pipeline.cuh"
#include "nfm_sym_ops.cuh"

namespace nfm {

#if NFM_PART == 0
// J^T H J (mode 0) or J H J^T (mode 1, K == D), packed output (_impl/sym.py:532-670)
template <typename T, int K, int D, int MODE>
struct SymMatmulOp {
  using scalar = T;
  static constexpr int kHN = MODE == 0 ? K : D;   // order of H
  static constexpr int kON = MODE == 0 ? D : K;   // order of the result
  static constexpr int kLen0 = K * D, kLen1 = packed_len(kHN), kLen2 = 1, kUse = 3, kOut = packed_len(kON);
  static constexpr bool kHeavy = false;
  __device__ static __forceinline__ void apply(const T (&j)[kLen0], const T (&h)[kLen1], const T (&)[1], int, int, T (&out)[kOut]) {
    sym_jhj<T, K, D, MODE>(j, h, out);
  }
};


template <typename T, int K, int D>
static int matmul_run(int mode, const KParams& p, cudaStream_t s) {
  if (mode == 0) return run_op<SymMatmulOp<T, K, D, 0>>(p, s);
  if constexpr (K == D && K <= 3) return run_op<SymMatmulOp<T, K, D, 1>>(p, s);  // the reference's unrolled branches
  else return NFM_E_UNSUPPORTED;
}
#define NFM_FUSED_RUN matmul_run
#elif NFM_PART == 1
template <typename T, int K, int D>
static int matmul_solve_run(int mode, const KParams& p, cudaStream_t s) {
  if (mode == 0) return run_op<SymMatmulSolveOp<T, K, D, 0>>(p, s);
  if constexpr (K == D && K <= 3) return run_op<SymMatmulSolveOp<T, K, D, 1>>(p, s);
  else return NFM_E_UNSUPPORTED;
}
#define NFM_FUSED_RUN matmul_solve_run
#endif

#if NFM_PART <= 1
template <typename T, int K>
static int fused_d(int d, int mode, const KParams& p, cudaStream_t s) {
  switch (d) {
    case 1: return NFM_FUSED_RUN<T, K, 1>(mode, p, s);
    case 2: return NFM_FUSED_RUN<T, K, 2>(mode, p, s);
    case 3: return NFM_FUSED_RUN<T, K, 3>(mode, p, s);
    case 4: return NFM_FUSED_RUN<T, K, 4>(mode, p, s);
    case 5: return NFM_FUSED_RUN<T, K, 5>(mode, p, s);
    case 6: return NFM_FUSED_RUN<T, K, 6>(mode, p, s);
  }
  return NFM_E_UNSUPPORTED;
}

// tall Jacobians (many channels / features, a 1..3-parameter displacement): k = 7..10, d <= 3
template <typename T, int K>
static int fused_d_tall(int d, int mode, const KParams& p, cudaStream_t s) {
  switch (d) {
    case 1: return NFM_FUSED_RUN<T, K, 1>(mode, p, s);
    case 2: return NFM_FUSED_RUN<T, K, 2>(mode, p, s);
    case 3: return NFM_FUSED_RUN<T, K, 3>(mode, p, s);
  }
  return NFM_E_UNSUPPORTED;
}

template <typename T>
static int fused_kd(int k, int d, int mode, const KParams& p, cudaStream_t s) {
  switch (k) {
    case 1: return fused_d<T, 1>(d, mode, p, s);
    case 2: return fused_d<T, 2>(d, mode, p, s);
    case 3: return fused_d<T, 3>(d, mode, p, s);
    case 4: return fused_d<T, 4>(d, mode, p, s);
    case 5: return fused_d<T, 5>(d, mode, p, s);
    case 6: return fused_d<T, 6>(d, mode, p, s);
    case 7: return fused_d_tall<T, 7>(d, mode, p, s);
    case 8: return fused_d_tall<T, 8>(d, mode, p, s);
    case 9: return fused_d_tall<T, 9>(d, mode, p, s);
    case 10: return fused_d_tall<T, 10>(d, mode, p, s);
  }
  return NFM_E_UNSUPPORTED;
}
#endif

#if NFM_PART == 0
template <typename T>
int sym_matmul_impl(int k, int d, int mode, const KParams& p, cudaStream_t s) { return fused_kd<T>(k, d, mode, p, s); }
template int sym_matmul_impl<NFM_SCALAR>(int, int, int, const KParams&, cudaStream_t);
#elif NFM_PART == 1
template <typename T>
int sym_matmul_solve_impl(int k, int d, int mode, const KParams& p, cudaStream_t s) { return fused_kd<T>(k, d, mode, p, s); }
template int sym_matmul_solve_impl<NFM_SCALAR>(int, int, int, const KParams&, cudaStream_t);
#else
template <typename T, int ALGO> struct SolveUpdBind { template <int N> using Op = SymSolveUpdateOp<T, N, ALGO>; };
template <typename T>
int sym_solve_update_impl(int n, int algo, const KParams& p, cudaStream_t s) {
  if (algo == NFM_ALGO_LDL) return DispatchN<SolveUpdBind<T, NFM_ALGO_LDL>::template Op, 1, NFM_MAX_N>::run(n, p, s);
  return DispatchN<SolveUpdBind<T, NFM_ALGO_AUTO>::template Op, 1, NFM_MAX_N>::run(n, p, s);
}
template int sym_solve_update_impl<NFM_SCALAR>(int, int, const KParams&, cudaStream_t);
#endif

}  // namespace nfm
