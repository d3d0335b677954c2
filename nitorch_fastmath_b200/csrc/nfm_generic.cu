// nfm_generic.cu -- run-time-sized kernels for the shapes the templated
// register kernels do not cover: rectangular m x n matvec (batchmatvec's
// "other sizes" branch, _impl/batched.py:175-176) and solves with several
// right-hand sides (sugar.lmdiv, sugar.py:75-137).  Correct for any stride;
// not on the measured hot path.
#include "nfm_pipeline.cuh"
#include "nfm_sym_math.cuh"

namespace nfm {

template <typename T>
__global__ void __launch_bounds__(128) matvec_rt_kernel(const T* __restrict__ mat, i64 ms, const T* __restrict__ vec,
                                                        i64 vs, T* __restrict__ out, i64 os, int m, int n, i64 batch) {
  const i64 total = batch * m;
  for (i64 t = i64(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += i64(gridDim.x) * blockDim.x) {
    const i64 b = t / m;
    const int i = int(t - b * m);
    const T* a = mat + b * ms + i64(i) * n;
    const T* v = vec + b * vs;
    T s = a[0] * v[0];
    for (int j = 1; j < n; ++j) s += a[j] * v[j];
    out[b * os + i] = s;
  }
}

// LU with partial pivoting (LDL^T when chol) in local memory, then nrhs substitutions
template <typename T>
__global__ void __launch_bounds__(128) solve_rt_kernel(const T* mat, i64 as, const T* rhs, i64 bs, T* out, i64 os, int n,
                                                       int nrhs, int chol, i64 batch) {
  for (i64 b = i64(blockIdx.x) * blockDim.x + threadIdx.x; b < batch; b += i64(gridDim.x) * blockDim.x) {
    T a[NFM_MAX_N][NFM_MAX_N];
    int perm[NFM_MAX_N];
    const T* src = mat + b * as;
    for (int i = 0; i < n; ++i) {
      perm[i] = i;
      for (int j = 0; j < n; ++j) a[i][j] = chol ? src[(i > j ? i : j) * n + (i > j ? j : i)] : src[i * n + j];
    }
    if (!chol) {
      for (int k = 0; k < n; ++k) {
        int p = k;
        T best = tabs(a[k][k]);
        for (int i = k + 1; i < n; ++i) {
          const T c = tabs(a[i][k]);
          if (c > best) { best = c; p = i; }
        }
        if (p != k) {
          for (int j = 0; j < n; ++j) { const T t = a[k][j]; a[k][j] = a[p][j]; a[p][j] = t; }
          const int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
        }
        const T rp = T(1) / a[k][k];
        for (int i = k + 1; i < n; ++i) {
          const T f = a[i][k] * rp;
          a[i][k] = f;
          for (int j = k + 1; j < n; ++j) a[i][j] -= f * a[k][j];
        }
      }
    } else {
      // LDL^T: strictly-lower part holds L, diagonal holds D
      for (int k = 0; k < n; ++k) {
        const T rp = T(1) / a[k][k];
        for (int i = k + 1; i < n; ++i)
          for (int j = k + 1; j <= i; ++j) a[i][j] -= a[i][k] * a[j][k] * rp;
        for (int i = k + 1; i < n; ++i) a[i][k] *= rp;
      }
    }
    const T* bb = rhs + b * bs;
    T* xx = out + b * os;
    for (int c = 0; c < nrhs; ++c) {
      T x[NFM_MAX_N];
      for (int i = 0; i < n; ++i) x[i] = bb[perm[i] * nrhs + c];
      for (int i = 0; i < n; ++i)  // L y = P b
        for (int j = 0; j < i; ++j) x[i] -= a[i][j] * x[j];
      if (chol) {
        for (int i = 0; i < n; ++i) x[i] /= a[i][i];
        for (int i = n - 1; i >= 0; --i)
          for (int j = i + 1; j < n; ++j) x[i] -= a[j][i] * x[j];
      } else {
        for (int i = n - 1; i >= 0; --i) {
          for (int j = i + 1; j < n; ++j) x[i] -= a[i][j] * x[j];
          x[i] /= a[i][i];
        }
      }
      for (int i = 0; i < n; ++i) xx[i * nrhs + c] = x[i];
    }
  }
}

// J^T H J (mode 0) or J H J^T (mode 1, k == d) for any 1 <= k, d <= 10:
// run-time-sized fallback of the templated SymMatmulOp (k, d <= 4)
template <typename T>
__global__ void __launch_bounds__(128) matmul_rt_kernel(const T* __restrict__ jac, i64 js, const T* __restrict__ hess, i64 hs,
                                                        T* __restrict__ out, i64 os, int k, int d, int mode, i64 batch) {
  const int hn = mode == 0 ? k : d;  // order of H
  const int on = mode == 0 ? d : k;  // order of the result
  for (i64 b = i64(blockIdx.x) * blockDim.x + threadIdx.x; b < batch; b += i64(gridDim.x) * blockDim.x) {
    const T* j = jac + b * js;
    const T* h = hess + b * hs;
    T hj[NFM_MAX_N][NFM_MAX_N];
    for (int a = 0; a < hn; ++a)
      for (int o = 0; o < on; ++o) {
        T s = T(0);
        for (int c = 0; c < hn; ++c) {
          const int lo = a < c ? a : c, hi = a < c ? c : a;
          const int idx = lo == hi ? lo : hn + lo * hn - (lo * (lo + 1)) / 2 + (hi - lo - 1);
          s += h[idx] * (mode == 0 ? j[c * d + o] : j[o * d + c]);
        }
        hj[a][o] = s;
      }
    T* dst = out + b * os;
    for (int o = 0; o < on; ++o)
      for (int q = o; q < on; ++q) {
        T s = T(0);
        for (int a = 0; a < hn; ++a) s += (mode == 0 ? j[a * d + o] : j[o * d + a]) * hj[a][q];
        dst[o == q ? o : on + o * on - (o * (o + 1)) / 2 + (q - o - 1)] = s;
      }
  }
}

static unsigned grid_for(i64 work);

template <typename T>
int sym_matmul_rt(int k, int d, int mode, i64 batch, const void* jac, i64 js, const void* hess, i64 hs, void* out, i64 os,
                  cudaStream_t s) {
  if (batch == 0) return 0;
  matmul_rt_kernel<T><<<grid_for(batch), 128, 0, s>>>(static_cast<const T*>(jac), js, static_cast<const T*>(hess), hs,
                                                      static_cast<T*>(out), os, k, d, mode, batch);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}
template int sym_matmul_rt<float>(int, int, int, i64, const void*, i64, const void*, i64, void*, i64, cudaStream_t);
template int sym_matmul_rt<double>(int, int, int, i64, const void*, i64, const void*, i64, void*, i64, cudaStream_t);

static unsigned grid_for(i64 work) {
  i64 blocks = (work + 127) / 128;
  const i64 cap = i64(device_info().sm_count) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return unsigned(blocks);
}

template <typename T>
int batch_matvec_rt(int m, int n, i64 batch, const void* mat, i64 ms, const void* vec, i64 vs, void* out, i64 os,
                    cudaStream_t s) {
  if (batch == 0) return 0;
  matvec_rt_kernel<T><<<grid_for(batch * m), 128, 0, s>>>(static_cast<const T*>(mat), ms, static_cast<const T*>(vec), vs,
                                                          static_cast<T*>(out), os, m, n, batch);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}

template <typename T>
int batch_solve_rt(int n, int nrhs, int chol, i64 batch, const void* a, i64 as, const void* b, i64 bs, void* out, i64 os,
                   cudaStream_t s) {
  if (batch == 0) return 0;
  solve_rt_kernel<T><<<grid_for(batch), 128, 0, s>>>(static_cast<const T*>(a), as, static_cast<const T*>(b), bs,
                                                     static_cast<T*>(out), os, n, nrhs, chol, batch);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return int(cudaGetLastError());
}

template int batch_matvec_rt<float>(int, int, i64, const void*, i64, const void*, i64, void*, i64, cudaStream_t);
template int batch_matvec_rt<double>(int, int, i64, const void*, i64, const void*, i64, void*, i64, cudaStream_t);
template int batch_solve_rt<float>(int, int, int, i64, const void*, i64, const void*, i64, void*, i64, cudaStream_t);
template int batch_solve_rt<double>(int, int, int, i64, const void*, i64, const void*, i64, void*, i64, cudaStream_t);

}  // namespace nfm
