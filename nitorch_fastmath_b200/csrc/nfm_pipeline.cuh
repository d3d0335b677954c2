// nfm_pipeline.cuh -- the two kernels every op runs through.
//
//   tile_kernel<Op, THREADS, MPT, STAGES, SEG>  (fast path)
//     Persistent CTAs, one thread per matrix (MPT matrices per thread per
//     tile).  The AoS records (coefficient dimension last) of a tile of
//     TILE = THREADS*MPT consecutive matrices are one contiguous byte range
//     per operand, so each operand tile moves HBM -> shared memory with 1-D
//     TMA bulk copies (cp.async.bulk, completion on an mbarrier) into a
//     STAGES-deep ring; threads pull their own record out of shared memory
//     with the widest conflict-free access, compute in registers, stage the
//     result record in shared memory and the tile goes back with TMA bulk
//     stores.  Every HBM access is therefore a full, aligned, contiguous
//     burst regardless of the record length.
//
//     SEG = false: one bulk copy per operand per tile, record r of the tile at
//       r * record_bytes.  Conflict free whenever record_bytes is an odd
//       multiple of its widest access (all packed-symmetric sizes but two).
//     SEG = true: for records that are a multiple of 32 B (dense 4x4, ...),
//       where consecutive threads would hit the same banks 2..8-way: the tile
//       is cut into 8 segments, each landing 16 B further than a dense layout
//       would put it (8 bulk copies per operand, issued by 8 lanes), and
//       thread m takes record (m / 8) of segment (m % 8).  The 8 lanes of a
//       128-bit shared-memory phase then sit in 8 different segments, i.e. 8
//       different 16 B bank groups: conflict free for every record size.
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
//     static constexpr bool kHeavy;                // pivoted / long dependent chains (tile geometry hint)
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

// Ops that take the two scalar parameters of KParams declare `kScalars`.
template <class Op, class = void>
struct has_scalars : std::false_type {};
template <class Op>
struct has_scalars<Op, std::void_t<decltype(Op::kScalars)>> : std::true_type {};

template <class Op, class R0, class R1, class R2, class O>
__device__ __forceinline__ void apply_op(const KParams& p, const R0& r0, const R1& r1, const R2& r2, O& o) {
  using T = typename Op::scalar;
  if constexpr (has_scalars<Op>::value) Op::apply(r0, r1, r2, p.present, p.flags, T(p.scal0), T(p.scal1), o);
  else Op::apply(r0, r1, r2, p.present, p.flags, o);
}

constexpr int kSegs = 8;     // segments per tile in the SEG layout
constexpr int kSegPad = 16;  // bytes of skew per segment

template <class Op, int THREADS, int MPT, bool SEG>
struct TileGeom {
  using T = typename Op::scalar;
  static constexpr int kTile = THREADS * MPT;
  static constexpr int kPad = SEG ? kSegs * kSegPad : 0;
  // payload bytes of one operand tile / its footprint in shared memory
  static constexpr int bytes(int len) { return kTile * len * int(sizeof(T)); }
  static constexpr int footprint(int len) { return bytes(len) + kPad; }
  static constexpr int kBytes0 = bytes(Op::kLen0), kBytes1 = bytes(Op::kLen1), kBytes2 = bytes(Op::kLen2);
  static constexpr int kBytesOut = bytes(Op::kOut);
  static constexpr int kFootOut = footprint(Op::kOut);
  static_assert(kTile % 64 == 0, "tile must be a multiple of 64 matrices");
  static_assert(!SEG || (bytes(Op::kLen0) / kSegs) % 16 == 0, "segments must keep the 16 B alignment of bulk copies");
  static_assert(kBytes0 % 16 == 0 && kBytes1 % 16 == 0 && kBytes2 % 16 == 0 && kBytesOut % 16 == 0,
                "tile byte counts must be multiples of 16 for TMA bulk copies");

  // byte offset of record m (0 <= m < kTile, in thread order) inside an operand tile
  template <int LEN>
  __device__ static __forceinline__ int rec_offset(int m) {
    if constexpr (SEG) {
      constexpr int seg_bytes = bytes(LEN) / kSegs;
      return (m & (kSegs - 1)) * (seg_bytes + kSegPad) + (m >> 3) * LEN * int(sizeof(T));
    } else {
      return m * LEN * int(sizeof(T));
    }
  }
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

// element e (0 <= e < count*LEN, tile-global order) -> byte offset in the staged tile
template <class G, int LEN>
__device__ __forceinline__ int elem_offset(int e) {
  using T = typename G::T;
  if constexpr (G::kPad != 0) {
    constexpr int per_seg = G::kTile / kSegs * LEN;  // elements per segment
    const int seg = e / per_seg;
    return seg * (per_seg * int(sizeof(T)) + kSegPad) + (e - seg * per_seg) * int(sizeof(T));
  } else {
    return e * int(sizeof(T));
  }
}

// ragged last tile: all threads copy `count` records between global memory and
// the staged layout with plain coalesced accesses
template <class G, int LEN>
__device__ __forceinline__ void coop_load(unsigned char* smem_tile, const typename G::T* src, int count) {
  using T = typename G::T;
  constexpr int U = 4;  // independent loads in flight per thread
  const int total = count * LEN, step = int(blockDim.x);
  for (int e0 = threadIdx.x; e0 < total; e0 += U * step) {
    T v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = (e0 + u * step < total) ? src[e0 + u * step] : T(0);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (e0 + u * step < total) *reinterpret_cast<T*>(smem_tile + elem_offset<G, LEN>(e0 + u * step)) = v[u];
  }
}

template <class G, int LEN>
__device__ __forceinline__ void coop_store(typename G::T* dst, const unsigned char* smem_tile, int count) {
  using T = typename G::T;
  const int total = count * LEN;
  for (int e = threadIdx.x; e < total; e += blockDim.x)
    dst[e] = *reinterpret_cast<const T*>(smem_tile + elem_offset<G, LEN>(e));
}

template <class Op, int THREADS, int MPT, int STAGES, bool SEG>
__global__ void __launch_bounds__(THREADS) tile_kernel(const __grid_constant__ KParams p, const i64 ntiles) {
  // ntiles = number of FULL tiles (moved by TMA).  A ragged last tile of
  // rem = p.batch - ntiles*TILE matrices is handled by the CTA whose turn it
  // is, with a cooperative guarded copy in place of the bulk copies.
  using T = typename Op::scalar;
  using G = TileGeom<Op, THREADS, MPT, SEG>;
  constexpr int TILE = G::kTile;
  const int rem = int(p.batch - ntiles * TILE);
  const i64 ntiles_all = ntiles + (rem > 0 ? 1 : 0);
  constexpr int kIssuers = SEG ? kSegs : 1;  // threads that issue bulk copies (one segment each)
  // L2 evict_first on the loads only for ops that write at least as much as they read
  constexpr bool kHint = Op::kOut >= ((Op::kUse & 1) ? Op::kLen0 : 0) + ((Op::kUse & 2) ? Op::kLen1 : 0);

  extern __shared__ __align__(128) unsigned char smem[];
  const int staged = staged_mask(p);
  const int f0 = (staged & 1) ? G::footprint(Op::kLen0) : 0;
  const int f1 = (staged & 2) ? G::footprint(Op::kLen1) : 0;
  const int f2 = (staged & 4) ? G::footprint(Op::kLen2) : 0;
  const int stage_bytes = f0 + f1 + f2;
  const uint32_t tx_bytes = ((staged & 1) ? G::kBytes0 : 0) + ((staged & 2) ? G::kBytes1 : 0) + ((staged & 4) ? G::kBytes2 : 0);

  unsigned char* const in_base = smem;
  unsigned char* const out_base = smem + STAGES * stage_bytes;
  uint64_t* const full = reinterpret_cast<uint64_t*>(out_base + 2 * G::kFootOut);

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
  }
  if (tid < kIssuers) policy = policy_evict_first();
  __syncthreads();
  // programmatic dependent launch: everything above overlapped the previous
  // kernel's tail; its results are visible only after this wait
  grid_dependency_wait();
  grid_launch_dependents();

  // producer (threads 0..kIssuers-1): thread 0 arms the stage barrier with the
  // tile's byte count; each issuer sends its segment of every staged operand
  auto issue = [&](int stage, i64 tile) {
    unsigned char* dst = in_base + stage * stage_bytes;
    const i64 first = tile * TILE;
    if (tid == 0) mbar_arrive_expect_tx(&full[stage], tx_bytes);
    constexpr int nseg = SEG ? kSegs : 1;
    const int seg = SEG ? tid : 0;
    if (staged & 1) {
      constexpr int sb = G::kBytes0 / nseg;
      bulk_g2s<kHint>(dst + seg * (sb + (SEG ? kSegPad : 0)), reinterpret_cast<const unsigned char*>(g0 + first * Op::kLen0) + seg * sb,
               sb, &full[stage], policy);
    }
    if (staged & 2) {
      constexpr int sb = G::kBytes1 / nseg;
      bulk_g2s<kHint>(dst + f0 + seg * (sb + (SEG ? kSegPad : 0)),
               reinterpret_cast<const unsigned char*>(g1 + first * Op::kLen1) + seg * sb, sb, &full[stage], policy);
    }
    if (staged & 4) {
      constexpr int sb = G::kBytes2 / nseg;
      bulk_g2s<kHint>(dst + f0 + f1 + seg * (sb + (SEG ? kSegPad : 0)),
               reinterpret_cast<const unsigned char*>(g2 + first * Op::kLen2) + seg * sb, sb, &full[stage], policy);
    }
  };

  if (tid < kIssuers) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      const i64 t = i64(blockIdx.x) + i64(s) * gridDim.x;
      if (t < ntiles) issue(s, t);
    }
  }

  int it = 0;
  for (i64 tile = blockIdx.x; tile < ntiles_all; tile += gridDim.x, ++it) {
    const int stage = it % STAGES;
    const uint32_t parity = uint32_t(it / STAGES) & 1u;
    unsigned char* sin = in_base + stage * stage_bytes;
    unsigned char* sout = out_base + (it & 1) * G::kFootOut;
    const bool ragged = tile >= ntiles;  // at most once, as this CTA's last tile

    if (!ragged) {
      mbar_wait(&full[stage], parity);
    } else {
      // no bulk copy was issued into this stage: fill it by hand in the same layout
      const i64 first = tile * TILE;
      if (staged & 1) coop_load<G, Op::kLen0>(sin, g0 + first * Op::kLen0, rem);
      if (staged & 2) coop_load<G, Op::kLen1>(sin + f0, g1 + first * Op::kLen1, rem);
      if (staged & 4) coop_load<G, Op::kLen2>(sin + f0 + f1, g2 + first * Op::kLen2, rem);
      __syncthreads();
    }

    // staged operands come out of shared memory; broadcast (stride 0) operands
    // are one record for the whole batch, re-read through L1; absent ones are 0
    T r0[MPT][Op::kLen0], r1[MPT][Op::kLen1], r2[MPT][Op::kLen2];
#pragma unroll
    for (int j = 0; j < MPT; ++j) {
      const int m = tid + j * THREADS;
      if (staged & 1) load_record(reinterpret_cast<const T*>(sin + G::template rec_offset<Op::kLen0>(m)), r0[j]);
      else if (p.present & 1) load_record_scalar(g0, r0[j]);
      else zero_record(r0[j]);
      if (staged & 2) load_record(reinterpret_cast<const T*>(sin + f0 + G::template rec_offset<Op::kLen1>(m)), r1[j]);
      else if (p.present & 2) load_record_scalar(g1, r1[j]);
      else zero_record(r1[j]);
      if (staged & 4) load_record(reinterpret_cast<const T*>(sin + f0 + f1 + G::template rec_offset<Op::kLen2>(m)), r2[j]);
      else if (p.present & 4) load_record_scalar(g2, r2[j]);
      else zero_record(r2[j]);
    }

    // the output buffer we are about to overwrite was last read by the bulk
    // store issued two tiles ago: allow one store (group) still pending
    if (tid < kIssuers) bulk_wait_read<1>();
    __syncthreads();  // every thread has its inputs in registers: stage is free
    if (tid < kIssuers) {
      const i64 nxt = tile + i64(STAGES) * gridDim.x;
      if (nxt < ntiles) issue(stage, nxt);
    }

#pragma unroll
    for (int j = 0; j < MPT; ++j) {
      T o[Op::kOut];
      apply_op<Op>(p, r0[j], r1[j], r2[j], o);
      store_record(reinterpret_cast<T*>(sout + G::template rec_offset<Op::kOut>(tid + j * THREADS)), o);
    }
    if (ragged) {
      __syncthreads();
      coop_store<G, Op::kOut>(gout + tile * TILE * Op::kOut, sout, rem);
      break;
    }
    fence_proxy_async();
    __syncthreads();
    if (tid < kIssuers) {
      constexpr int nseg = SEG ? kSegs : 1;
      constexpr int sb = G::kBytesOut / nseg;
      const int seg = SEG ? tid : 0;
      bulk_s2g(reinterpret_cast<unsigned char*>(gout + tile * TILE * Op::kOut) + seg * sb,
               sout + seg * (sb + (SEG ? kSegPad : 0)), sb);
      bulk_commit();
    }
  }
  if (tid < kIssuers) bulk_wait<0>();
}

template <class Op>
__global__ void __launch_bounds__(128) strided_kernel(const __grid_constant__ KParams p) {
  using T = typename Op::scalar;
  const T* const g0 = static_cast<const T*>(p.in[0].ptr);
  const T* const g1 = static_cast<const T*>(p.in[1].ptr);
  const T* const g2 = static_cast<const T*>(p.in[2].ptr);
  T* const gout = static_cast<T*>(p.out);
  grid_dependency_wait();
  grid_launch_dependents();
  for (i64 b = i64(blockIdx.x) * blockDim.x + threadIdx.x; b < p.batch; b += i64(gridDim.x) * blockDim.x) {
    T r0[Op::kLen0], r1[Op::kLen1], r2[Op::kLen2], o[Op::kOut];
    zero_record(r0);
    zero_record(r1);
    zero_record(r2);
    if (p.present & 1) load_record_scalar(g0 + b * p.in[0].stride, r0, elem_stride(p.in[0].estride));
    if (p.present & 2) load_record_scalar(g1 + b * p.in[1].stride, r1, elem_stride(p.in[1].estride));
    if (p.present & 4) load_record_scalar(g2 + b * p.in[2].stride, r2, elem_stride(p.in[2].estride));
    apply_op<Op>(p, r0, r1, r2, o);
    store_record_scalar(gout + b * p.out_stride, o, elem_stride(p.out_estride));
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
int current_device();

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Tile geometry.  Measured on B200 (profiles/r1_tile_geometry_sweep.txt): the
// best configurations keep ONE large CTA per SM with 2 (big records) or 3
// stages and ~130-150 KB of shared memory -- large bulk copies beat many small
// ones -- so the rule is: stages = 2 when a matrix moves >= 100 B, else 3;
// TILE = the largest of {64,...,1024} (2048 / 4096 for records of a few bytes) whose ring (mandatory
// operands) + double-buffered output fits 152 KB; THREADS = 512 / 384 / 256 /
// TILE.  Optional operands (regulariser, addend) grow the ring up to the
// 227 KB limit.  TuneFixed pins ops whose measured optimum differs.
// -DNFM_TUNE_TILE / NFM_TUNE_THREADS / NFM_TUNE_STAGES / NFM_TUNE_SEG override
// the rule for tuning builds.
template <class Op>
struct TuneBase {
  using T = typename Op::scalar;
  // operand 2 is the optional one in every op that has it
  static constexpr int kInBytes = (((Op::kUse >> 0) & 1) * Op::kLen0 + ((Op::kUse >> 1) & 1) * Op::kLen1) * int(sizeof(T));
  static constexpr int kOutBytes = Op::kOut * int(sizeof(T));
  // Records that are a multiple of 32 B bank-conflict in the dense layout:
  // 2-way at 32 B, 4-way at 64 B, 8-way at 128 B.  The segmented layout removes
  // that but its 8x smaller bulk copies cost ~7-15 %, so it is used when an
  // operand conflicts >= 4-way, or 2-way on an operand that carries at least
  // half of the traffic (a 2-way conflict on a small vector is cheaper).
  static constexpr int rec_bytes(int len) { return len * int(sizeof(T)); }
  static constexpr bool heavy(int len) { return rec_bytes(len) % 64 == 0; }
  static constexpr bool light(int len) { return rec_bytes(len) % 32 == 0 && 2 * rec_bytes(len) >= kInBytes + kOutBytes; }
  static constexpr bool conflicts(int len) { return heavy(len) || light(len); }
#ifdef NFM_TUNE_SEG
  static constexpr bool kSeg = NFM_TUNE_SEG;
#else
  static constexpr bool kSeg = (((Op::kUse >> 0) & 1) && conflicts(Op::kLen0)) || (((Op::kUse >> 1) & 1) && conflicts(Op::kLen1)) ||
                               conflicts(Op::kOut);
#endif
};

constexpr int kSmemTarget = 152 * 1024;

constexpr int pick_tile(int stages, int in_bytes, int out_bytes) {
  const int per_matrix = stages * in_bytes + 2 * out_bytes;
  const int cands[9] = {4096, 2048, 1024, 768, 512, 384, 256, 128, 64};
  for (int i = (per_matrix <= 48 ? 0 : per_matrix <= 96 ? 1 : 2); i < 9; ++i)
    if (cands[i] * per_matrix <= kSmemTarget) return cands[i];
  return 64;
}

constexpr int pick_threads(int tile) { return tile >= 1024 ? 512 : tile == 768 ? 384 : tile; }  // one matrix per thread up to 512; MPT <= 8

// Compute-heavy ops (pivoted elimination, Gauss-Jordan: Op::kHeavy) are bound by
// the latency of their dependent chains rather than by HBM alone; they do best
// with SEVERAL small CTAs per SM whose barrier phases interleave: 128 threads,
// one matrix each, 3 stages (2 when the ring would pass 72 KB), tile 64 when
// even that passes 100 KB  (profiles/r1_tile_geometry_sweep.txt, sweep 3).
constexpr int heavy_stages(int in_bytes, int out_bytes) { return 128 * (3 * in_bytes + 2 * out_bytes) <= 72 * 1024 ? 3 : 2; }
constexpr int heavy_tile(int in_bytes, int out_bytes) {
  return 128 * (heavy_stages(in_bytes, out_bytes) * in_bytes + 2 * out_bytes) <= 100 * 1024 ? 128 : 64;
}

template <class Op>
struct TuneRule : TuneBase<Op> {
  using B = TuneBase<Op>;
#ifdef NFM_TUNE_STAGES
  static constexpr int kStages = NFM_TUNE_STAGES;
#else
  static constexpr int kStages = Op::kHeavy ? heavy_stages(B::kInBytes, B::kOutBytes) : (B::kInBytes + B::kOutBytes >= 100) ? 2 : 3;
#endif
#ifdef NFM_TUNE_TILE
  static constexpr int kTile = NFM_TUNE_TILE;
#else
  static constexpr int kTile = Op::kHeavy ? heavy_tile(B::kInBytes, B::kOutBytes) : pick_tile(kStages, B::kInBytes, B::kOutBytes);
#endif
#ifdef NFM_TUNE_THREADS
  static constexpr int kThreads = NFM_TUNE_THREADS < kTile ? NFM_TUNE_THREADS : kTile;
#else
  static constexpr int kThreads = Op::kHeavy ? kTile : pick_threads(kTile);
#endif
  static constexpr int kMpt = kTile / kThreads;
};

template <class Op, int TILE, int THREADS, int STAGES>
struct TuneFixed : TuneBase<Op> {
#if defined(NFM_TUNE_TILE) || defined(NFM_TUNE_THREADS) || defined(NFM_TUNE_STAGES)
  static constexpr int kTile = TuneRule<Op>::kTile, kThreads = TuneRule<Op>::kThreads, kStages = TuneRule<Op>::kStages;
#else
  static constexpr int kTile = TILE, kThreads = THREADS, kStages = STAGES;
#endif
  static constexpr int kMpt = kTile / kThreads;
};

template <class Op>
struct Tune : TuneRule<Op> {};

// Second, smaller geometry for light ops with small records when a launch has
// only a few tiles per CTA (mid-size batches: a multi-GPU slab, config 1).
// Measured for the 3x3 fp32 solve (tune_main "small"): 512-matrix tiles beat
// 1024 by 10 / 6 / 3.5 % at 0.5M / 1M / 2M matrices and lose by 3 % at 4M.
template <class Op>
struct TuneSmall {
  using B = TuneBase<Op>;
  static constexpr bool kEnabled = !Op::kHeavy && (B::kInBytes + B::kOutBytes <= 64) && Tune<Op>::kTile >= 1024;
  static constexpr int kThreads = 256, kMpt = 2, kStages = 3;
  static constexpr int kTilesPerCtaBelow = 20;  // use it when the big geometry would give fewer tiles per SM than this
};

bool pdl_enabled();  // nfm_entry.cu: false when the environment has NFM_DISABLE_PDL=1

// Launch with programmatic stream serialization: the kernel may become resident
// while the previous kernel in the stream drains; it touches global memory only
// after griddepcontrol.wait, so ordering is unchanged.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;  // NFM_DISABLE_PDL=1 turns it off
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// per (kernel instantiation, device, staged mask): resident CTAs per SM, 0 = not yet known
struct LaunchCache {
  std::atomic<int> per_sm[16][8];
  std::atomic<int> attr_set[16];
};

template <class Op, int THREADS, int MPT, int STAGES, bool SEG>
int launch_tile(const KParams& p, i64 ntiles, cudaStream_t stream) {
  using G = TileGeom<Op, THREADS, MPT, SEG>;
  static LaunchCache cache;  // zero-initialised
  auto kern = tile_kernel<Op, THREADS, MPT, STAGES, SEG>;
  const int staged = staged_mask(p);
  auto smem_for = [](int mask) {
    const int stage = ((mask & 1) ? G::footprint(Op::kLen0) : 0) + ((mask & 2) ? G::footprint(Op::kLen1) : 0) +
                      ((mask & 4) ? G::footprint(Op::kLen2) : 0);
    return STAGES * stage + 2 * G::kFootOut + STAGES * 8 + 16;
  };
  const int smem = smem_for(staged);
  const DeviceInfo& dev = device_info();
  if (smem > dev.max_smem_optin) return -1;  // caller falls back to the strided kernel
  const int d = current_device() & 15;
  if (!cache.attr_set[d].load(std::memory_order_acquire)) {
    int most = smem_for(Op::kUse & 7);
    if (most > dev.max_smem_optin) most = dev.max_smem_optin;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
    if (e != cudaSuccess) return int(e);
    cache.attr_set[d].store(1, std::memory_order_release);
  }
  int per_sm = cache.per_sm[d][staged].load(std::memory_order_acquire);
  if (per_sm == 0) {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
    if (e != cudaSuccess) return int(e);
    if (per_sm < 1) return -1;
    cache.per_sm[d][staged].store(per_sm, std::memory_order_release);
  }
  const i64 ntiles_all = ntiles + (p.batch > ntiles * G::kTile ? 1 : 0);
  i64 grid = i64(dev.sm_count) * per_sm;
  if (grid > ntiles_all) grid = ntiles_all;
  cudaError_t e = launch_pdl(kern, unsigned(grid), THREADS, size_t(smem), stream, p, ntiles);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(e);
}

template <class Op>
int launch_strided(const KParams& p, cudaStream_t stream) {
  if (p.batch <= 0) return 0;
  const DeviceInfo& dev = device_info();
  i64 blocks = (p.batch + 127) / 128;
  const i64 cap = i64(dev.sm_count) * 16;
  if (blocks > cap) blocks = cap;
  cudaError_t e = launch_pdl(strided_kernel<Op>, unsigned(blocks), 128, size_t(0), stream, p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(e);
}

// Run an op over p.batch matrices: TMA-staged tiles for the bulk when every
// operand is dense-or-broadcast and 16 B aligned, strided kernel for the rest.
template <class Op>
int run_op(KParams p, cudaStream_t stream) {
  using Tn = Tune<Op>;
  constexpr int TILE = Tn::kThreads * Tn::kMpt;
  t_last_path_tma = 0;
  if (p.batch == 0) return NFM_OK;
  const int lens[kMaxIn] = {Op::kLen0, Op::kLen1, Op::kLen2};
  bool fast = p.out_stride == Op::kOut && elem_stride(p.out_estride) == 1 && aligned16(p.out);
  int nstaged = 0;
  for (int i = 0; i < kMaxIn && fast; ++i) {
    if (!((p.present >> i) & 1)) continue;
    if (elem_stride(p.in[i].estride) != 1) fast = false;
    if (p.in[i].stride == 0) continue;  // broadcast
    if (p.in[i].stride != lens[i] || !aligned16(p.in[i].ptr)) fast = false;
    ++nstaged;
  }
  if (fast && nstaged == 0) fast = false;
  if (fast) {
    // full tiles by TMA, the ragged remainder inside the same launch
    int rc;
    bool small = false;
    if constexpr (TuneSmall<Op>::kEnabled) small = p.batch / TILE < i64(TuneSmall<Op>::kTilesPerCtaBelow) * device_info().sm_count;
    if constexpr (TuneSmall<Op>::kEnabled) {
      using Ts = TuneSmall<Op>;
      rc = small ? launch_tile<Op, Ts::kThreads, Ts::kMpt, Ts::kStages, Tn::kSeg>(p, p.batch / (Ts::kThreads * Ts::kMpt), stream)
                 : launch_tile<Op, Tn::kThreads, Tn::kMpt, Tn::kStages, Tn::kSeg>(p, p.batch / TILE, stream);
    } else {
      rc = launch_tile<Op, Tn::kThreads, Tn::kMpt, Tn::kStages, Tn::kSeg>(p, p.batch / TILE, stream);
    }
    if (rc == 0) {
      t_last_path_tma = 1;
      return NFM_OK;
    }
    if (rc > 0) {
      set_error("tile kernel launch failed: %s", cudaGetErrorString(cudaError_t(rc)));
      return rc;
    }
  }
  int rc = launch_strided<Op>(p, stream);
  if (rc != 0) set_error("strided kernel launch failed: %s", cudaGetErrorString(cudaError_t(rc)));
  return rc;
}

}  // namespace nfm
