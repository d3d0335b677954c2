// tune_main.cu -- development tool (not part of the library): sweeps the tile
// geometry <THREADS, MPT, STAGES, SEG> of tile_kernel for the headline ops and
// prints achieved GB/s per configuration.   make tune && ./nfm_tune
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "nfm_dense_ops.cuh"
#include "nfm_pipeline.cuh"
#include "nfm_sym_ops.cuh"

namespace nfm {
std::atomic<unsigned long long> g_launch_count{0};
thread_local int t_last_path_tma = 0;
void set_error(const char*, ...) {}
const DeviceInfo& device_info() {
  static DeviceInfo d{};
  if (d.sm_count == 0) {
    cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0);
  }
  return d;
}
int current_device() { return 0; }
bool pdl_enabled() { return true; }
}  // namespace nfm

using namespace nfm;

template <typename T>
__global__ void fill_kernel(T* p, i64 n, int rec, int ndiag, T diag, T off) {
  for (i64 i = i64(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += i64(gridDim.x) * blockDim.x) {
    const int k = int(i % rec);
    // dense records: diagonal at k % (ndiag+1) == 0 when ndiag > 0 means row-major n x n
    bool is_diag = ndiag < 0 ? (k % (-ndiag + 1) == 0) : (k < ndiag);
    p[i] = (is_diag ? diag : off) + T(1e-3) * T(i % 7);
  }
}

struct Buffers {
  void *in0, *in1, *out;
  i64 batch;
};

template <class Op, int THREADS, int MPT, int STAGES, bool SEG>
void run_config(const char* name, const Buffers& b, int alg_bytes) {
  using T = typename Op::scalar;
  KParams p{};
  p.in[0].ptr = b.in0;
  p.in[0].stride = Op::kLen0;
  p.present = 1;
  if (Op::kUse & 2) {
    p.in[1].ptr = b.in1;
    p.in[1].stride = Op::kLen1;
    p.present |= 2;
  }
  p.out = b.out;
  p.out_stride = Op::kOut;
  constexpr int TILE = THREADS * MPT;
  const i64 ntiles = b.batch / TILE;
  p.batch = b.batch;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int rc = 0;
  for (int i = 0; i < 3 && rc == 0; ++i) rc = launch_tile<Op, THREADS, MPT, STAGES, SEG>(p, ntiles, 0);
  if (rc != 0) {
    printf("%-22s T=%4d thr=%3d st=%d seg=%d : launch failed rc=%d\n", name, TILE, THREADS, STAGES, int(SEG), rc);
    cudaGetLastError();
    return;
  }
  cudaDeviceSynchronize();
  const int reps = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) launch_tile<Op, THREADS, MPT, STAGES, SEG>(p, ntiles, 0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms * 1e3 / reps;
  const double gbs = double(p.batch) * alg_bytes / (us * 1e-6) / 1e9;
  printf("%-22s T=%4d thr=%3d st=%d seg=%d : %8.1f us  %7.1f GB/s  %6.2f Gmat/s\n", name, TILE, THREADS, STAGES, int(SEG), us,
         gbs, p.batch / (us * 1e-6) / 1e9);
  fflush(stdout);
}

// small-batch variant: the big buffers are cut into `slices` sub-batches that are
// visited round-robin, so that consecutive launches never find their data in L2
template <class Op, int THREADS, int MPT, int STAGES, bool SEG>
void run_sliced(const char* name, const Buffers& b, int alg_bytes, i64 sub) {
  using T = typename Op::scalar;
  const int slices = int(b.batch / sub);
  std::vector<KParams> ps(slices);
  for (int s = 0; s < slices; ++s) {
    KParams p{};
    p.in[0].ptr = static_cast<const T*>(b.in0) + i64(s) * sub * Op::kLen0;
    p.in[0].stride = Op::kLen0;
    p.present = 1;
    if (Op::kUse & 2) {
      p.in[1].ptr = static_cast<const T*>(b.in1) + i64(s) * sub * Op::kLen1;
      p.in[1].stride = Op::kLen1;
      p.present |= 2;
    }
    p.out = static_cast<T*>(b.out) + i64(s) * sub * Op::kOut;
    p.out_stride = Op::kOut;
    p.batch = sub;
    ps[s] = p;
  }
  constexpr int TILE = THREADS * MPT;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < slices; ++i) launch_tile<Op, THREADS, MPT, STAGES, SEG>(ps[i], sub / TILE, 0);
  cudaDeviceSynchronize();
  const int reps = 10 * slices;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) launch_tile<Op, THREADS, MPT, STAGES, SEG>(ps[i % slices], sub / TILE, 0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms * 1e3 / reps;
  printf("%-18s batch %8lld T=%4d thr=%3d st=%d : %7.2f us  %7.1f GB/s\n", name, sub, TILE, THREADS, STAGES, us,
         double(sub) * alg_bytes / (us * 1e-6) / 1e9);
  fflush(stdout);
}

template <typename T>
Buffers make(i64 batch, int len0, int ndiag0, int len1, int lout) {
  Buffers b{};
  b.batch = batch;
  cudaMalloc(&b.in0, size_t(batch) * len0 * sizeof(T));
  cudaMalloc(&b.in1, size_t(batch) * (len1 > 0 ? len1 : 1) * sizeof(T));
  cudaMalloc(&b.out, size_t(batch) * lout * sizeof(T));
  fill_kernel<T><<<1184, 256>>>(static_cast<T*>(b.in0), batch * len0, len0, ndiag0, T(8), T(0.25));
  if (len1 > 0) fill_kernel<T><<<1184, 256>>>(static_cast<T*>(b.in1), batch * len1, len1, 0, T(1), T(1));
  cudaDeviceSynchronize();
  return b;
}

void release(Buffers& b) {
  cudaFree(b.in0);
  cudaFree(b.in1);
  cudaFree(b.out);
}

#define CFG(OP, NAME, THR, MPT, ST, SEG, BYTES) run_config<OP, THR, MPT, ST, SEG>(NAME, buf, BYTES)

int main(int argc, char** argv) {
  // usage: nfm_tune [substring]   -- run only the blocks whose name contains it
  const char* only = argc > 1 ? argv[1] : "";
  auto want = [&](const char* n) { return only[0] == 0 || strstr(n, only) != nullptr; };

  // The three sweeps that produced the Tune<> rule are recorded in
  // profiles/r1_tile_geometry_sweep.txt; this list is the regression check:
  // the geometry the rule picks for each headline op next to its neighbours.
  if (want("solve3")) {
    using Op = SymSolveOp<float, 3, NFM_LAYOUT_SYM, 0>;
    Buffers buf = make<float>(256ll * 256 * 256, 6, 3, 3, 3);
    CFG(Op, "sym_solve3", 512, 2, 3, false, 48);   // rule
    CFG(Op, "sym_solve3", 256, 2, 3, false, 48);
    CFG(Op, "sym_solve3", 256, 4, 4, false, 48);
    CFG(Op, "sym_solve3", 1024, 1, 3, false, 48);
    release(buf);
  }
  if (want("small")) {
    using Op = SymSolveOp<float, 3, NFM_LAYOUT_SYM, 0>;
    Buffers buf = make<float>(256ll * 256 * 256, 6, 3, 3, 3);
    for (i64 sub : {i64(1) << 19, i64(1) << 20, i64(1) << 21, i64(1) << 22}) {
      run_sliced<Op, 512, 2, 3, false>("solve3", buf, 48, sub);
      run_sliced<Op, 256, 2, 3, false>("solve3", buf, 48, sub);
      run_sliced<Op, 512, 1, 3, false>("solve3", buf, 48, sub);
      run_sliced<Op, 256, 1, 4, false>("solve3", buf, 48, sub);
      run_sliced<Op, 128, 2, 4, false>("solve3", buf, 48, sub);
      run_sliced<Op, 256, 1, 3, false>("solve3", buf, 48, sub);
    }
    release(buf);
  }
  if (want("solve6")) {
    using Op = SymSolveOp<float, 6, NFM_LAYOUT_SYM, NFM_ALGO_AUTO>;
    using Ldl = SymSolveOp<float, 6, NFM_LAYOUT_SYM, NFM_ALGO_LDL>;
    Buffers buf = make<float>(192ll * 192 * 192, 21, 6, 6, 6);
    CFG(Op, "sym_solve6 auto", 512, 1, 2, false, 132);   // rule
    CFG(Ldl, "sym_solve6 ldl", 512, 1, 2, false, 132);
    CFG(Op, "sym_solve6 auto", 256, 2, 2, false, 132);
    CFG(Op, "sym_solve6 auto", 384, 1, 3, false, 132);
    release(buf);
  }
  if (want("invert6")) {
    using Op = SymInvertOp<float, 6, NFM_ALGO_AUTO, false>;
    Buffers buf = make<float>(192ll * 192 * 192, 21, 6, 0, 21);
    CFG(Op, "sym_invert6 auto", 384, 1, 2, false, 168);  // rule
    CFG(Op, "sym_invert6 auto", 256, 1, 2, false, 168);
    CFG(Op, "sym_invert6 auto", 512, 1, 2, false, 168);
    release(buf);
  }
  if (want("solve10")) {
    using Op = SymSolveOp<float, 10, NFM_LAYOUT_SYM, NFM_ALGO_AUTO>;
    Buffers buf = make<float>(160ll * 160 * 160, 55, 10, 10, 10);
    CFG(Op, "sym_solve10 auto", 256, 1, 2, false, 300);  // rule
    CFG(Op, "sym_solve10 auto", 128, 1, 2, false, 300);
    CFG(Op, "sym_solve10 auto", 128, 1, 3, false, 300);
    release(buf);
  }
  if (want("inv4d")) {
    using Op = BatchInvOp<double, 4, NFM_ALGO_AUTO>;
    Buffers buf = make<double>(16ll << 20, 16, -4, 0, 16);
    CFG(Op, "dense_inv4d", 256, 1, 3, true, 256);   // pinned (TuneFixed)
    CFG(Op, "dense_inv4d", 256, 1, 3, false, 256);  // dense layout: 8-way bank conflicts
    CFG(Op, "dense_inv4d", 128, 1, 3, true, 256);
    release(buf);
  }
  if (want("solve4d")) {
    using Op = BatchSolveOp<double, 4, NFM_ALGO_LU>;
    Buffers buf = make<double>(16ll << 20, 16, -4, 4, 4);
    CFG(Op, "dense_solve4d", 128, 1, 3, true, 192);  // pinned
    CFG(Op, "dense_solve4d", 256, 1, 3, true, 192);
    CFG(Op, "dense_solve4d", 512, 1, 2, true, 192);
    release(buf);
  }
  return 0;
}
