// nfm_warp.cuh -- NFM_ALGO_WARP: the sub-warp cooperative, shuffle-based solve
// for packed symmetric matrices of order 5..10 (the "one warp per matrix"
// design of the north star, with several matrices packed per warp so lanes do
// not idle: groups of 8 lanes for N <= 8, 16 lanes for N = 9, 10).
//
// Lane r of a group owns row r of the expanded matrix and entry r of the
// right-hand side.  Elimination step k broadcasts the pivot row from lane k
// with __shfl_sync (register indices stay compile-time constants), every lane
// below updates its own row; back substitution broadcasts each finished
// unknown.  No pivoting (same arithmetic class as LDL^T).  Data movement is
// the same TMA ring as tile_kernel: operand tiles arrive by bulk copy, lanes
// gather their row from shared memory, results leave by bulk store.
//
// This is the A/B variant.  It needs ~4x (N=6) to ~7x (N=10) the issue slots
// of the thread-per-matrix kernels, so it is not the
// default; `method='warp'` / NFM_ALGO_WARP selects it.
#pragma once

#include "nfm_pipeline.cuh"
#include "nfm_sym_math.cuh"

namespace nfm {

template <typename T, int N, int THREADS, int TILE, int STAGES>
__global__ void __launch_bounds__(THREADS) warp_solve_kernel(const __grid_constant__ KParams p, const i64 ntiles) {
  constexpr int G = N <= 8 ? 8 : 16;   // lanes per matrix
  constexpr int PER_WARP = 32 / G;     // matrices per warp per pass
  constexpr int NWARPS = THREADS / 32;
  constexpr int NN = packed_len(N);
  constexpr int kBytesMat = TILE * NN * int(sizeof(T));
  constexpr int kBytesVec = TILE * N * int(sizeof(T));
  static_assert(TILE % (NWARPS * PER_WARP) == 0, "tile must be a whole number of warp passes");

  extern __shared__ __align__(128) unsigned char smem[];
  const bool has_diag = (p.present & 4) != 0;
  const int stage_bytes = kBytesMat + kBytesVec + (has_diag ? kBytesVec : 0);
  unsigned char* const in_base = smem;
  unsigned char* const out_base = smem + STAGES * stage_bytes;
  uint64_t* const full = reinterpret_cast<uint64_t*>(out_base + 2 * kBytesVec);

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int r = lane % G, sub = lane / G;
  const bool active = r < N;
  const T* const gmat = static_cast<const T*>(p.in[0].ptr);
  const T* const gvec = static_cast<const T*>(p.in[1].ptr);
  const T* const gdiag = static_cast<const T*>(p.in[2].ptr);
  T* const gout = static_cast<T*>(p.out);

  // packed position of a_rj for this lane's row
  int off[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const int lo = r < j ? r : j, hi = r < j ? j : r;
    off[j] = !active ? 0 : (lo == hi ? lo : N + lo * N - (lo * (lo + 1)) / 2 + (hi - lo - 1));
  }

  uint64_t policy = 0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
    policy = policy_evict_first();
  }
  __syncthreads();
  grid_dependency_wait();
  grid_launch_dependents();

  auto issue = [&](int stage, i64 tile) {
    unsigned char* dst = in_base + stage * stage_bytes;
    const i64 first = tile * TILE;
    mbar_arrive_expect_tx(&full[stage], uint32_t(stage_bytes));
    bulk_g2s<false>(dst, gmat + first * NN, kBytesMat, &full[stage], policy);
    bulk_g2s<false>(dst + kBytesMat, gvec + first * N, kBytesVec, &full[stage], policy);
    if (has_diag) bulk_g2s<false>(dst + kBytesMat + kBytesVec, gdiag + first * N, kBytesVec, &full[stage], policy);
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      const i64 t = i64(blockIdx.x) + i64(s) * gridDim.x;
      if (t < ntiles) issue(s, t);
    }
  }

  int it = 0;
  for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int stage = it % STAGES;
    const uint32_t parity = uint32_t(it / STAGES) & 1u;
    const T* smat = reinterpret_cast<const T*>(in_base + stage * stage_bytes);
    const T* svec = reinterpret_cast<const T*>(in_base + stage * stage_bytes + kBytesMat);
    const T* sdiag = reinterpret_cast<const T*>(in_base + stage * stage_bytes + kBytesMat + kBytesVec);
    T* sout = reinterpret_cast<T*>(out_base + (it & 1) * kBytesVec);

    mbar_wait(&full[stage], parity);
    if (tid == 0) bulk_wait_read<1>();  // output buffer (it & 1) free again
    __syncthreads();

    for (int m0 = warp * PER_WARP; m0 < TILE; m0 += NWARPS * PER_WARP) {
      const int m = m0 + sub;
      T a[N];
      T b = T(0);
#pragma unroll
      for (int j = 0; j < N; ++j) a[j] = active ? smat[m * NN + off[j]] : (j == r ? T(1) : T(0));
      if (active) b = svec[m * N + r];
      if (has_diag && active) {
        const T d = sdiag[m * N + r];
#pragma unroll
        for (int j = 0; j < N; ++j) a[j] += (j == r) ? d : T(0);
      }
      // forward elimination, pivot row k broadcast from lane k of the group;
      // every lane keeps the reciprocal pivots (one division per step)
      T rp[N];
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const T piv = __shfl_sync(0xffffffffu, a[k], k, G);
        rp[k] = T(1) / piv;
        if (k == N - 1) break;
        const T bk = __shfl_sync(0xffffffffu, b, k, G);
        const T f = (r > k) ? a[k] * rp[k] : T(0);
#pragma unroll
        for (int j = k + 1; j < N; ++j) {
          const T ukj = __shfl_sync(0xffffffffu, a[j], k, G);
          a[j] -= f * ukj;
        }
        b -= f * bk;
      }
      // back substitution: unknown j is finished by lane j and broadcast
#pragma unroll
      for (int j = N - 1; j >= 0; --j) {
        const T xj = __shfl_sync(0xffffffffu, b * rp[j], j, G);
        if (r == j) b = xj;
        else if (r < j) b -= a[j] * xj;
      }
      if (active) sout[m * N + r] = b;
    }

    fence_proxy_async();
    __syncthreads();  // all rows read, all results staged
    if (tid == 0) {
      const i64 nxt = tile + i64(STAGES) * gridDim.x;
      if (nxt < ntiles) issue(stage, nxt);
      bulk_s2g(gout + tile * TILE * N, sout, kBytesVec);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait<0>();
}

// host: full tiles through warp_solve_kernel; returns the number of matrices done
template <typename T, int N>
int launch_warp_solve(const KParams& p, cudaStream_t stream, i64* done) {
  constexpr int kRec = (packed_len(N) + 2 * N) * int(sizeof(T));  // mat + vec + optional regulariser
  constexpr int TILE = kRec * 256 * 2 + 2 * 256 * N * int(sizeof(T)) <= 200 * 1024 ? 256 : 128;
  constexpr int THREADS = 512;
  constexpr int STAGES = kRec * TILE * 3 <= 160 * 1024 ? 3 : 2;
  *done = 0;
  const i64 ntiles = p.batch / TILE;
  if (ntiles == 0) return 0;
  auto kern = warp_solve_kernel<T, N, THREADS, TILE, STAGES>;
  const bool has_diag = (p.present & 4) != 0;
  const int stage = TILE * (packed_len(N) + N + (has_diag ? N : 0)) * int(sizeof(T));
  const int smem = STAGES * stage + 2 * TILE * N * int(sizeof(T)) + STAGES * 8 + 16;
  const DeviceInfo& dev = device_info();
  if (smem > dev.max_smem_optin) return 0;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return int(e);
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
  if (e != cudaSuccess) return int(e);
  if (per_sm < 1) return 0;
  i64 grid = i64(dev.sm_count) * per_sm;
  if (grid > ntiles) grid = ntiles;
  e = launch_pdl(kern, unsigned(grid), THREADS, size_t(smem), stream, p, ntiles);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  if (e == cudaSuccess) *done = ntiles * TILE;
  return int(e);
}

}  // namespace nfm
