// nfm_pipeline.cuh -- the two kernels every op runs through.
//
//   tile_kernel<Op, THREADS, MPT, STAGES>  (fast path)
//     Persistent CTAs, one thread per matrix (MPT matrices per thread per
//     tile).  The AoS records (coefficient dimension last) of a tile of
//     TILE = THREADS*MPT consecutive matrices are one contiguous byte range
//     per operand, so each operand tile moves HBM -> shared memory with ONE
//     1-D TMA bulk copy (cp.async.bulk, completion on an mbarrier) into a
//     STAGES-deep ring; threads pull their own record out of shared memory
//     with the widest conflict-free access, compute in registers, stage the
//     result record in shared memory and one thread sends the whole output
//     tile back with a TMA bulk store.  Every HBM access is therefore a full,
//     aligned, contiguous burst regardless of the record length.
//
//   strided_kernel<Op>  (general path)
//     One thread per matrix straight from global memory with arbitrary batch
//     strides / alignment / broadcast.  Also runs the ragged tail
//     (batch % TILE) of the fast path.
//
// An Op is a stateless struct:
//     using scalar = float|double;
//     static constexpr int kLen0, kLen1, kLen2;   // input record lengths (1 if unused)
//     static constexpr int kUse;                   // bit mask of inputs the op can take
//     static constexpr int kOut;                   // output record length
//     __device__ static void apply(const T(&)[kLen0], const T(&)[kLen1], const T(&)[kLen2],
//                                  int present, int flags, T(&out)[kOut]);
// Absent optional inputs arrive zero-filled.
#pragma once

#include <atomic>

#include "nfm_common.cuh"

namespace nfm {

extern std::atomic<unsigned long long> g_launch_count;
extern thread_local int t_last_path_tma;
void set_error(const char* fmt, ...);

constexpr int round16(int x) { return (x + 15) & ~15; }

template <class Op, int THREADS, int MPT>
struct TileGeom {
  using T = typename Op::scalar;
  static constexpr int kTile = THREADS * MPT;
  static constexpr int kBytes0 = kTile * Op::kLen0 * int(sizeof(T));
  static constexpr int kBytes1 = kTile * Op::kLen1 * int(sizeof(T));
  static constexpr int kBytes2 = kTile * Op::kLen2 * int(sizeof(T));
  static constexpr int kBytesOut = kTile * Op::kOut * int(sizeof(T));
  static_assert(kBytes0 % 16 == 0 && kBytes1 % 16 == 0 && kBytes2 % 16 == 0 && kBytesOut % 16 == 0,
                "tile byte counts must be multiples of 16 for TMA bulk copies");
};

// which operands are staged through shared memory for this launch:
// present, not broadcast.  (Eligibility -- dense stride, 16 B alignment -- is
// checked on the host.)
__host__ __device__ inline int staged_mask(const KParams& p) {
  int m = 0;
  for (int i = 0; i < kMaxIn; ++i)
    if (((p.present >> i) & 1) && p.in[i].stride != 0) m |= 1 << i;
  return m;
}

template <class Op, int THREADS, int MPT, int STAGES>
__global__ void __launch_bounds__(THREADS) tile_kernel(const __grid_constant__ KParams p, const i64 ntiles) {
  using T = typename Op::scalar;
  using G = TileGeom<Op, THREADS, MPT>;
  constexpr int TILE = G::kTile;

  extern __shared__ __align__(128) unsigned char smem[];
  const int staged = staged_mask(p);
  const int b0 = (staged & 1) ? G::kBytes0 : 0;
  const int b1 = (staged & 2) ? G::kBytes1 : 0;
  const int b2 = (staged & 4) ? G::kBytes2 : 0;
  const int stage_bytes = b0 + b1 + b2;

  unsigned char* const in_base = smem;
  unsigned char* const out_base = smem + STAGES * stage_bytes;
  uint64_t* const full = reinterpret_cast<uint64_t*>(out_base + 2 * G::kBytesOut);

  const int tid = threadIdx.x;
  const T* const g0 = static_cast<const T*>(p.in[0].ptr);
  const T* const g1 = static_cast<const T*>(p.in[1].ptr);
  const T* const g2 = static_cast<const T*>(p.in[2].ptr);
  T* const gout = static_cast<T*>(p.out);

  uint64_t policy = 0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
    policy = policy_evict_first();
  }
  __syncthreads();

  // producer: one elected thread arms the stage barrier with the byte count
  // and issues one bulk copy per staged operand
  auto issue = [&](int stage, i64 tile) {
    unsigned char* dst = in_base + stage * stage_bytes;
    const i64 first = tile * TILE;
    mbar_arrive_expect_tx(&full[stage], uint32_t(stage_bytes));
    if (staged & 1) bulk_g2s(dst, g0 + first * Op::kLen0, G::kBytes0, &full[stage], policy);
    if (staged & 2) bulk_g2s(dst + b0, g1 + first * Op::kLen1, G::kBytes1, &full[stage], policy);
    if (staged & 4) bulk_g2s(dst + b0 + b1, g2 + first * Op::kLen2, G::kBytes2, &full[stage], policy);
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
    const unsigned char* sin = in_base + stage * stage_bytes;
    T* sout = reinterpret_cast<T*>(out_base + (it & 1) * G::kBytesOut);

    mbar_wait(&full[stage], parity);

    // staged operands come out of shared memory; broadcast (stride 0) operands
    // are one record for the whole batch, re-read through L1; absent ones are 0
    T r0[MPT][Op::kLen0], r1[MPT][Op::kLen1], r2[MPT][Op::kLen2];
#pragma unroll
    for (int j = 0; j < MPT; ++j) {
      const int m = tid + j * THREADS;
      if (staged & 1) load_record(reinterpret_cast<const T*>(sin) + m * Op::kLen0, r0[j]);
      else if (p.present & 1) load_record_scalar(g0, r0[j]);
      else zero_record(r0[j]);
      if (staged & 2) load_record(reinterpret_cast<const T*>(sin + b0) + m * Op::kLen1, r1[j]);
      else if (p.present & 2) load_record_scalar(g1, r1[j]);
      else zero_record(r1[j]);
      if (staged & 4) load_record(reinterpret_cast<const T*>(sin + b0 + b1) + m * Op::kLen2, r2[j]);
      else if (p.present & 4) load_record_scalar(g2, r2[j]);
      else zero_record(r2[j]);
    }

    // the output buffer we are about to overwrite was last read by the bulk
    // store issued two tiles ago: allow one store still pending
    if (tid == 0) bulk_wait_read<1>();
    __syncthreads();  // every thread has its inputs in registers: stage is free
    if (tid == 0) {
      const i64 nxt = tile + i64(STAGES) * gridDim.x;
      if (nxt < ntiles) issue(stage, nxt);
    }

#pragma unroll
    for (int j = 0; j < MPT; ++j) {
      T o[Op::kOut];
      Op::apply(r0[j], r1[j], r2[j], p.present, p.flags, o);
      store_record(sout + (tid + j * THREADS) * Op::kOut, o);
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(gout + tile * TILE * Op::kOut, sout, G::kBytesOut);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait<0>();
}

template <class Op>
__global__ void __launch_bounds__(128) strided_kernel(const __grid_constant__ KParams p) {
  using T = typename Op::scalar;
  const T* const g0 = static_cast<const T*>(p.in[0].ptr);
  const T* const g1 = static_cast<const T*>(p.in[1].ptr);
  const T* const g2 = static_cast<const T*>(p.in[2].ptr);
  T* const gout = static_cast<T*>(p.out);
  for (i64 b = i64(blockIdx.x) * blockDim.x + threadIdx.x; b < p.batch; b += i64(gridDim.x) * blockDim.x) {
    T r0[Op::kLen0], r1[Op::kLen1], r2[Op::kLen2], o[Op::kOut];
    zero_record(r0);
    zero_record(r1);
    zero_record(r2);
    if (p.present & 1) load_record_scalar(g0 + b * p.in[0].stride, r0);
    if (p.present & 2) load_record_scalar(g1 + b * p.in[1].stride, r1);
    if (p.present & 4) load_record_scalar(g2 + b * p.in[2].stride, r2);
    Op::apply(r0, r1, r2, p.present, p.flags, o);
    store_record_scalar(gout + b * p.out_stride, o);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct DeviceInfo {
  int sm_count;
  int max_smem_optin;
};
const DeviceInfo& device_info();  // cached per current device

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Default tile geometry.  Targets (DESIGN.md "tile geometry"): >= ~16 KB of
// input per stage so that 2 resident CTAs x (STAGES-1) stages keep > 64 KB of
// loads in flight per SM; <= ~100 KB of shared memory per CTA so two CTAs fit.
template <class Op>
struct Tune {
  using T = typename Op::scalar;
  static constexpr int kInBytes = (((Op::kUse >> 0) & 1) * Op::kLen0 + ((Op::kUse >> 1) & 1) * Op::kLen1 +
                                   ((Op::kUse >> 2) & 1) * Op::kLen2) * int(sizeof(T));
  static constexpr int kOutBytes = Op::kOut * int(sizeof(T));
  static constexpr int kRec = kInBytes + kOutBytes;
  // matrices per tile: aim at ~24 KB of input per stage, clamp to [64, 1024]
  static constexpr int kWant = 24576 / (kInBytes > 0 ? kInBytes : 1);
  static constexpr int kTile = kWant >= 1024 ? 1024 : kWant >= 512 ? 512 : kWant >= 256 ? 256 : kWant >= 128 ? 128 : 64;
  static constexpr int kThreads = kTile >= 256 ? 256 : kTile;
  static constexpr int kMpt = kTile / kThreads;
  static constexpr int kStages = 3;
};

template <class Op, int THREADS, int MPT, int STAGES>
int launch_tile(const KParams& p, i64 ntiles, cudaStream_t stream) {
  using G = TileGeom<Op, THREADS, MPT>;
  auto kern = tile_kernel<Op, THREADS, MPT, STAGES>;
  const int staged = staged_mask(p);
  const int stage_bytes = ((staged & 1) ? G::kBytes0 : 0) + ((staged & 2) ? G::kBytes1 : 0) + ((staged & 4) ? G::kBytes2 : 0);
  const int smem = STAGES * stage_bytes + 2 * G::kBytesOut + STAGES * 8 + 16;
  const DeviceInfo& dev = device_info();
  if (smem > dev.max_smem_optin) return -1;  // caller falls back to the strided kernel
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return int(e);
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
  if (e != cudaSuccess) return int(e);
  if (per_sm < 1) return -1;
  i64 grid = i64(dev.sm_count) * per_sm;
  if (grid > ntiles) grid = ntiles;
  kern<<<unsigned(grid), THREADS, smem, stream>>>(p, ntiles);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}

template <class Op>
int launch_strided(const KParams& p, cudaStream_t stream) {
  if (p.batch <= 0) return 0;
  const DeviceInfo& dev = device_info();
  i64 blocks = (p.batch + 127) / 128;
  const i64 cap = i64(dev.sm_count) * 16;
  if (blocks > cap) blocks = cap;
  strided_kernel<Op><<<unsigned(blocks), 128, 0, stream>>>(p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}

// Run an op over p.batch matrices: TMA-staged tiles for the bulk when every
// operand is dense-or-broadcast and 16 B aligned, strided kernel for the rest.
template <class Op>
int run_op(KParams p, cudaStream_t stream) {
  using T = typename Op::scalar;
  using Tn = Tune<Op>;
  constexpr int TILE = Tn::kThreads * Tn::kMpt;
  t_last_path_tma = 0;
  if (p.batch == 0) return NFM_OK;
  const int lens[kMaxIn] = {Op::kLen0, Op::kLen1, Op::kLen2};
  bool fast = p.batch >= TILE && p.out_stride == Op::kOut && aligned16(p.out);
  int nstaged = 0;
  for (int i = 0; i < kMaxIn && fast; ++i) {
    if (!((p.present >> i) & 1)) continue;
    if (p.in[i].stride == 0) continue;  // broadcast
    if (p.in[i].stride != lens[i] || !aligned16(p.in[i].ptr)) fast = false;
    ++nstaged;
  }
  if (fast && nstaged == 0) fast = false;
  if (fast) {
    const i64 ntiles = p.batch / TILE;
    int rc = launch_tile<Op, Tn::kThreads, Tn::kMpt, Tn::kStages>(p, ntiles, stream);
    if (rc == 0) {
      t_last_path_tma = 1;
      const i64 done = ntiles * TILE;
      if (done == p.batch) return NFM_OK;
      // ragged tail through the strided kernel
      for (int i = 0; i < kMaxIn; ++i)
        if ((p.present >> i) & 1) p.in[i].ptr = static_cast<const T*>(p.in[i].ptr) + done * p.in[i].stride;
      p.out = static_cast<T*>(p.out) + done * p.out_stride;
      p.batch -= done;
    } else if (rc > 0) {
      set_error("tile kernel launch failed: %s", cudaGetErrorString(cudaError_t(rc)));
      return rc;
    }
  }
  int rc = launch_strided<Op>(p, stream);
  if (rc != 0) set_error("strided kernel launch failed: %s", cudaGetErrorString(cudaError_t(rc)));
  return rc;
}

}  // namespace nfm
