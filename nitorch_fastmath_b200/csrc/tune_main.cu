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
  const char* only = argc > 1 ? argv[1] : "";

  auto want = [&](const char* n) { return only[0] == 0 || strstr(n, only) != nullptr; };

  {
    Buffers buf = make<float>(192ll * 192 * 192, 21, 6, 6, 6);
    using A = SymSolveOp<float, 6, NFM_LAYOUT_SYM, NFM_ALGO_LDL>;
    using B = SymSolveOp<float, 6, NFM_LAYOUT_SYM, NFM_ALGO_AUTO>;
    CFG(A, "solve6 f32 ldl", 512, 1, 2, false, 132);
    CFG(B, "solve6 f32 auto", 512, 1, 2, false, 132);
    release(buf);
  }
  {
    Buffers buf = make<float>(160ll * 160 * 160, 55, 10, 10, 10);
    using A = SymSolveOp<float, 10, NFM_LAYOUT_SYM, NFM_ALGO_LDL>;
    using B = SymSolveOp<float, 10, NFM_LAYOUT_SYM, NFM_ALGO_AUTO>;
    CFG(A, "solve10 f32 ldl", 256, 1, 2, false, 300);
    CFG(B, "solve10 f32 auto", 256, 1, 2, false, 300);
    release(buf);
  }
  {
    Buffers buf = make<double>(128ll * 128 * 256, 55, 10, 10, 10);
    using A = SymSolveOp<double, 10, NFM_LAYOUT_SYM, NFM_ALGO_LDL>;
    using B = SymSolveOp<double, 10, NFM_LAYOUT_SYM, NFM_ALGO_AUTO>;
    CFG(A, "solve10 f64 ldl", 128, 1, 2, false, 600);
    CFG(B, "solve10 f64 auto", 128, 1, 2, false, 600);
    release(buf);
  }
  {
    Buffers buf = make<double>(128ll * 128 * 256, 45, 9, 9, 9);
    using A = SymSolveOp<double, 9, NFM_LAYOUT_SYM, NFM_ALGO_LDL>;
    using B = SymSolveOp<double, 9, NFM_LAYOUT_SYM, NFM_ALGO_AUTO>;
    CFG(A, "solve9 f64 ldl", 128, 1, 2, false, 504);
    CFG(B, "solve9 f64 auto", 128, 1, 2, false, 504);
    release(buf);
  }
  return 0;
}
