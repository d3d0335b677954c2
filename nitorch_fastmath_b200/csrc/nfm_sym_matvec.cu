// sym_matvec / sym_addmatvec / sym_submatvec for one scalar type (-DNFM_SCALAR)
#include "nfm_impl.cuh"
#include "nfm_sym_ops.cuh"

namespace nfm {

template <typename T, int LAYOUT>
struct MatvecBind {
  template <int N> using Op = SymMatvecOp<T, N, LAYOUT>;
};

template <typename T>
int sym_matvec_impl(int n, int layout, const KParams& p, cudaStream_t s) {
  switch (layout) {
    case NFM_LAYOUT_SCALED_IDENTITY:
      return DispatchN<MatvecBind<T, NFM_LAYOUT_SCALED_IDENTITY>::template Op, 1, NFM_MAX_N>::run(n, p, s);
    case NFM_LAYOUT_DIAG:
      return DispatchN<MatvecBind<T, NFM_LAYOUT_DIAG>::template Op, 1, NFM_MAX_N>::run(n, p, s);
    case NFM_LAYOUT_SYM:
      return DispatchN<MatvecBind<T, NFM_LAYOUT_SYM>::template Op, 1, NFM_MAX_N>::run(n, p, s);
    case NFM_LAYOUT_FULL:
      return DispatchN<MatvecBind<T, NFM_LAYOUT_FULL>::template Op, 1, NFM_MAX_N>::run(n, p, s);
  }
  return NFM_E_UNSUPPORTED;
}

template int sym_matvec_impl<NFM_SCALAR>(int, int, const KParams&, cudaStream_t);

}  // namespace nfm
