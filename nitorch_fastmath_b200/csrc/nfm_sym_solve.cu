// sym_solve for one scalar type (-DNFM_SCALAR) and one part (-DNFM_PART)
#include "nfm_impl.cuh"
#include "nfm_sym_ops.cuh"

namespace nfm {

template <typename T, int LAYOUT, int ALGO>
struct SolveBind {
  template <int N> using Op = SymSolveOp<T, N, LAYOUT, ALGO>;
};

#if NFM_PART == 0
template <typename T>
int sym_solve_part0(int n, int layout, const KParams& p, cudaStream_t s) {
  switch (layout) {
    case NFM_LAYOUT_SCALED_IDENTITY:
      return DispatchN<SolveBind<T, NFM_LAYOUT_SCALED_IDENTITY, 0>::template Op, 1, NFM_MAX_N>::run(n, p, s);
    case NFM_LAYOUT_DIAG:
      return DispatchN<SolveBind<T, NFM_LAYOUT_DIAG, 0>::template Op, 1, NFM_MAX_N>::run(n, p, s);
    case NFM_LAYOUT_SYM:
      return DispatchN<SolveBind<T, NFM_LAYOUT_SYM, 0>::template Op, 1, 4>::run(n, p, s);
    case NFM_LAYOUT_FULL:
      return DispatchN<SolveBind<T, NFM_LAYOUT_FULL, 0>::template Op, 1, NFM_MAX_N>::run(n, p, s);
  }
  return NFM_E_UNSUPPORTED;
}
template int sym_solve_part0<NFM_SCALAR>(int, int, const KParams&, cudaStream_t);
#elif NFM_PART == 1
template <typename T>
int sym_solve_part1(int n, const KParams& p, cudaStream_t s) {
  return DispatchN<SolveBind<T, NFM_LAYOUT_SYM, NFM_ALGO_LDL>::template Op, 5, NFM_MAX_N>::run(n, p, s);
}
template int sym_solve_part1<NFM_SCALAR>(int, const KParams&, cudaStream_t);
#elif NFM_PART == 2
template <typename T>
int sym_solve_part2(int n, const KParams& p, cudaStream_t s) {
  return DispatchN<SolveBind<T, NFM_LAYOUT_SYM, NFM_ALGO_LU>::template Op, 5, NFM_MAX_N>::run(n, p, s);
}
template int sym_solve_part2<NFM_SCALAR>(int, const KParams&, cudaStream_t);
#else
template <typename T>
int sym_solve_part3(int n, const KParams& p, cudaStream_t s) {
  return DispatchN<SolveBind<T, NFM_LAYOUT_SYM, NFM_ALGO_AUTO>::template Op, 5, NFM_MAX_N>::run(n, p, s);
}
template int sym_solve_part3<NFM_SCALAR>(int, const KParams&, cudaStream_t);
#endif

}  // namespace nfm
