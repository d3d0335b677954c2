// sym_invert for one scalar type (-DNFM_SCALAR) and one part (-DNFM_PART)
#include "nfm_impl.cuh"
#include "nfm_sym_ops.cuh"

namespace nfm {

template <typename T, int ALGO, bool DIAG_ONLY>
struct InvertBind {
  template <int N> using Op = SymInvertOp<T, N, ALGO, DIAG_ONLY>;
};

#if NFM_PART == 0
template <typename T>
int sym_invert_part0(int n, int diag_only, const KParams& p, cudaStream_t s) {
  if (diag_only) return DispatchN<InvertBind<T, NFM_ALGO_LDL, true>::template Op, 1, NFM_MAX_N>::run(n, p, s);
  return DispatchN<InvertBind<T, NFM_ALGO_LDL, false>::template Op, 1, NFM_MAX_N>::run(n, p, s);
}
template int sym_invert_part0<NFM_SCALAR>(int, int, const KParams&, cudaStream_t);
#elif NFM_PART == 1
template <typename T>
int sym_invert_part1(int n, int diag_only, const KParams& p, cudaStream_t s) {
  if (diag_only) return DispatchN<InvertBind<T, NFM_ALGO_LU, true>::template Op, 5, NFM_MAX_N>::run(n, p, s);
  return DispatchN<InvertBind<T, NFM_ALGO_LU, false>::template Op, 5, NFM_MAX_N>::run(n, p, s);
}
template int sym_invert_part1<NFM_SCALAR>(int, int, const KParams&, cudaStream_t);
#else
template <typename T>
int sym_invert_part2(int n, int diag_only, const KParams& p, cudaStream_t s) {
  if (diag_only) return DispatchN<InvertBind<T, NFM_ALGO_AUTO, true>::template Op, 5, NFM_MAX_N>::run(n, p, s);
  return DispatchN<InvertBind<T, NFM_ALGO_AUTO, false>::template Op, 5, NFM_MAX_N>::run(n, p, s);
}
template int sym_invert_part2<NFM_SCALAR>(int, int, const KParams&, cudaStream_t);
#endif

}  // namespace nfm
