// nfm_host.cu -- host-buffer entry points: chunked H2D -> kernel -> D2H
// pipelines over caller-provided streams and device workspace.  This is the
// end-to-end path (bench.py "e2e"): operands start and end in host memory.
#include "nfm_pipeline.cuh"
#include "nfm_sym_math.cuh"

namespace nfm {
namespace {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t x) { return (x + kAlign - 1) & ~(kAlign - 1); }

struct HostOperand {
  const void* host;  // nullable
  int elems;         // elements per matrix
};

inline size_t esize(int dtype) { return dtype == NFM_F64 ? 8 : 4; }

// bytes of one pipeline buffer (all operand chunks + the output chunk)
size_t buffer_bytes(int dtype, i64 chunk, const int* in_elems, int n_in, int out_elems) {
  size_t b = 0;
  for (int i = 0; i < n_in; ++i) b += align_up(size_t(chunk) * in_elems[i] * esize(dtype));
  b += align_up(size_t(chunk) * out_elems * esize(dtype));
  return b;
}

// Launch(d_in[3], d_out, count, stream) -> rc
template <class Launch>
int pipeline(int dtype, i64 batch, const HostOperand (&ops)[3], void* h_out, int out_elems, void* ws, size_t ws_bytes,
             i64 chunk, int nbuf, void** streams, Launch launch) {
  if (batch < 0 || chunk < 1 || nbuf < 1 || streams == nullptr || ws == nullptr || h_out == nullptr) {
    set_error("host pipeline: bad argument");
    return NFM_E_BADARG;
  }
  int elems[3];
  for (int i = 0; i < 3; ++i) elems[i] = ops[i].host ? ops[i].elems : 0;
  const size_t per_buf = buffer_bytes(dtype, chunk, elems, 3, out_elems);
  if (per_buf * size_t(nbuf) > ws_bytes) {
    set_error("host pipeline: workspace too small (%zu needed, %zu given)", per_buf * size_t(nbuf), ws_bytes);
    return NFM_E_WORKSPACE;
  }
  const size_t es = esize(dtype);
  int rc = 0;
  i64 done = 0;
  for (i64 c = 0; done < batch && rc == 0; ++c) {
    const i64 cnt = (batch - done < chunk) ? (batch - done) : chunk;
    const int slot = int(c % nbuf);
    auto s = static_cast<cudaStream_t>(streams[slot]);
    unsigned char* base = static_cast<unsigned char*>(ws) + size_t(slot) * per_buf;
    const void* d_in[3] = {nullptr, nullptr, nullptr};
    size_t off = 0;
    for (int i = 0; i < 3 && rc == 0; ++i) {
      if (!ops[i].host) continue;
      d_in[i] = base + off;
      const size_t bytes = size_t(cnt) * ops[i].elems * es;
      rc = int(cudaMemcpyAsync(base + off, static_cast<const unsigned char*>(ops[i].host) + size_t(done) * ops[i].elems * es,
                               bytes, cudaMemcpyHostToDevice, s));
      off += align_up(size_t(chunk) * ops[i].elems * es);
    }
    if (rc) break;
    void* d_out = base + off;
    rc = launch(d_in, d_out, cnt, s);
    if (rc) break;
    rc = int(cudaMemcpyAsync(static_cast<unsigned char*>(h_out) + size_t(done) * out_elems * es, d_out,
                             size_t(cnt) * out_elems * es, cudaMemcpyDeviceToHost, s));
    done += cnt;
  }
  for (int i = 0; i < nbuf; ++i) {
    const int e = int(cudaStreamSynchronize(static_cast<cudaStream_t>(streams[i])));
    if (rc == 0 && e != 0) rc = e;
  }
  if (rc > 0) set_error("host pipeline: %s", cudaGetErrorString(cudaError_t(rc)));
  return rc;
}

}  // namespace
}  // namespace nfm

using namespace nfm;

extern "C" {

size_t nfm_host_workspace_bytes(int dtype, int64_t chunk, int nbuf, int in_elems, int out_elems) {
  // upper bound: up to 3 separately aligned input chunks + the output chunk
  const size_t es = esize(dtype);
  return size_t(nbuf) * (align_up(size_t(chunk) * in_elems * es) + 3 * kAlign + align_up(size_t(chunk) * out_elems * es));
}

int nfm_sym_solve_host(int dtype, int n, int algo, int64_t batch, const void* h_mat, const void* h_vec, const void* h_diag,
                       void* h_out, void* d_workspace, size_t workspace_bytes, int64_t chunk, int nbuf, void** streams) {
  if (n < 1 || n > NFM_MAX_N || h_mat == nullptr || h_vec == nullptr) {
    set_error("sym_solve_host: bad argument");
    return NFM_E_BADARG;
  }
  const int nn = packed_len(n);
  const HostOperand ops[3] = {{h_mat, nn}, {h_vec, n}, {h_diag, n}};
  return pipeline(dtype, batch, ops, h_out, n, d_workspace, workspace_bytes, chunk, nbuf, streams,
                  [=](const void* const* d_in, void* d_out, i64 cnt, cudaStream_t s) {
                    return nfm_sym_solve(dtype, n, NFM_LAYOUT_SYM, algo, cnt, d_in[0], nn, d_in[1], n, d_in[2], n, d_out, n, s);
                  });
}

int nfm_sym_invert_host(int dtype, int n, int algo, int diag_only, int64_t batch, const void* h_mat, void* h_out,
                        void* d_workspace, size_t workspace_bytes, int64_t chunk, int nbuf, void** streams) {
  if (n < 1 || n > NFM_MAX_N || h_mat == nullptr) {
    set_error("sym_invert_host: bad argument");
    return NFM_E_BADARG;
  }
  const int nn = packed_len(n);
  const int no = diag_only ? n : nn;
  const HostOperand ops[3] = {{h_mat, nn}, {nullptr, 0}, {nullptr, 0}};
  return pipeline(dtype, batch, ops, h_out, no, d_workspace, workspace_bytes, chunk, nbuf, streams,
                  [=](const void* const* d_in, void* d_out, i64 cnt, cudaStream_t s) {
                    return nfm_sym_invert(dtype, n, algo, diag_only, cnt, d_in[0], nn, d_out, no, s);
                  });
}

int nfm_sym_matvec_host(int dtype, int n, int64_t batch, const void* h_mat, const void* h_vec, const void* h_inp, int sign,
                        void* h_out, void* d_workspace, size_t workspace_bytes, int64_t chunk, int nbuf, void** streams) {
  if (n < 1 || n > NFM_MAX_N || h_mat == nullptr || h_vec == nullptr) {
    set_error("sym_matvec_host: bad argument");
    return NFM_E_BADARG;
  }
  const int nn = packed_len(n);
  const HostOperand ops[3] = {{h_mat, nn}, {h_vec, n}, {h_inp, n}};
  return pipeline(dtype, batch, ops, h_out, n, d_workspace, workspace_bytes, chunk, nbuf, streams,
                  [=](const void* const* d_in, void* d_out, i64 cnt, cudaStream_t s) {
                    return nfm_sym_matvec(dtype, n, NFM_LAYOUT_SYM, cnt, d_in[0], nn, d_in[1], n, d_in[2], n, sign, d_out, n, s);
                  });
}

int nfm_batch_inv_host(int dtype, int n, int algo, int closed_form_reg, int64_t batch, const void* h_a, void* h_out,
                       void* d_workspace, size_t workspace_bytes, int64_t chunk, int nbuf, void** streams) {
  if (n < 1 || n > NFM_MAX_N || h_a == nullptr) {
    set_error("batch_inv_host: bad argument");
    return NFM_E_BADARG;
  }
  const int nn = n * n;
  const HostOperand ops[3] = {{h_a, nn}, {nullptr, 0}, {nullptr, 0}};
  return pipeline(dtype, batch, ops, h_out, nn, d_workspace, workspace_bytes, chunk, nbuf, streams,
                  [=](const void* const* d_in, void* d_out, i64 cnt, cudaStream_t s) {
                    return nfm_batch_inv(dtype, n, algo, closed_form_reg, cnt, d_in[0], nn, d_out, nn, s);
                  });
}

int nfm_batch_det_host(int dtype, int n, int64_t batch, const void* h_a, void* h_out, void* d_workspace,
                       size_t workspace_bytes, int64_t chunk, int nbuf, void** streams) {
  if (n < 1 || n > NFM_MAX_N || h_a == nullptr) {
    set_error("batch_det_host: bad argument");
    return NFM_E_BADARG;
  }
  const int nn = n * n;
  const HostOperand ops[3] = {{h_a, nn}, {nullptr, 0}, {nullptr, 0}};
  return pipeline(dtype, batch, ops, h_out, 1, d_workspace, workspace_bytes, chunk, nbuf, streams,
                  [=](const void* const* d_in, void* d_out, i64 cnt, cudaStream_t s) {
                    return nfm_batch_det(dtype, n, cnt, d_in[0], nn, d_out, 1, s);
                  });
}

int nfm_batch_solve_host(int dtype, int n, int nrhs, int algo, int64_t batch, const void* h_a, const void* h_b, void* h_out,
                         void* d_workspace, size_t workspace_bytes, int64_t chunk, int nbuf, void** streams) {
  if (n < 1 || n > NFM_MAX_N || nrhs < 1 || h_a == nullptr || h_b == nullptr) {
    set_error("batch_solve_host: bad argument");
    return NFM_E_BADARG;
  }
  const int nn = n * n, nk = n * nrhs;
  const HostOperand ops[3] = {{h_a, nn}, {h_b, nk}, {nullptr, 0}};
  return pipeline(dtype, batch, ops, h_out, nk, d_workspace, workspace_bytes, chunk, nbuf, streams,
                  [=](const void* const* d_in, void* d_out, i64 cnt, cudaStream_t s) {
                    return nfm_batch_solve(dtype, n, nrhs, algo, cnt, d_in[0], nn, d_in[1], nk, d_out, nk, s);
                  });
}

}  // extern "C"
