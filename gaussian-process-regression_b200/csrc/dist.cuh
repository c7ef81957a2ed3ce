// dist.cuh -- multi-GPU Cholesky + solve for matrices that do not fit one GPU (BASELINE config 5: n = 200 000,
// K = 320 GB in FP64).  One process per GPU; NCCL over NVLink/NVSwitch moves the panels.
//
// Layout: 1-D block-cyclic by OUTER PANEL (512 columns): panel p lives on rank p mod N as a full-height column slab
// (n_pad x 512, leading dimension n_pad).  Right-looking with one-panel lookahead:
//   owner(p):  factor panel p (the single-GPU panel sequence on the slab) -> pack rows >= 512 p -> ncclBroadcast
//   everyone:  update every owned panel right of p with the broadcast panel (DMMA GEMM, K = 512)
//   owner(p+1) updates its panel p + 1 first ("LA") on a high-priority stream, factors it and broadcasts it while the
//   other ranks are still busy with the rest of update p (two panel buffers, events order the reuse).
// Every panel crosses the links once: n^2/2 doubles received per GPU in total (160 GB at n = 200k, ~0.3 s at NVLink
// rates) against n^3/(3 N) flops per GPU (>= 9 s), so the factorisation stays bound by the FP64 tensor pipe.
// The two triangular solves walk the panels in order and pass the (small) right-hand side with a broadcast per panel.
//
// NCCL is bound at run time with dlopen (the process may already hold torch's bundled libnccl.so.2; the library has no
// link-time dependency on it).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "potrf.cuh"
#include "trsv.cuh"

namespace gprc {

constexpr int PANEL = OUTER_BLOCKS * NB;  // 512

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

inline int nccl_load(NcclApi& api, const char* path) {
  if (api.handle) return 0;
  const char* candidates[] = {path, "libnccl.so.2", "libnccl.so"};
  for (const char* cnd : candidates) {
    if (!cnd || !*cnd) continue;
    api.handle = dlopen(cnd, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return set_error(-4, __FILE__, __LINE__, "cannot dlopen libnccl.so.2");
#define GPRC_SYM(field, name)                                                             \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name));             \
  if (!api.field) return set_error(-4, __FILE__, __LINE__, "libnccl lacks " name)
  GPRC_SYM(GetUniqueId, "ncclGetUniqueId");
  GPRC_SYM(CommInitRank, "ncclCommInitRank");
  GPRC_SYM(CommDestroy, "ncclCommDestroy");
  GPRC_SYM(Broadcast, "ncclBroadcast");
  GPRC_SYM(AllReduce, "ncclAllReduce");
  GPRC_SYM(GetErrorString, "ncclGetErrorString");
#undef GPRC_SYM
  return 0;
}

#define GPRC_NCCL(api, call)                                                                        \
  do {                                                                                              \
    ncclResult_t r__ = (call);                                                                      \
    if (r__ != ncclSuccess) return gprc::set_error(-5, __FILE__, __LINE__, (api).GetErrorString(r__)); \
  } while (0)

// Trailing update of owned slabs with a broadcast panel:
//   C(ti, gj) -= P(ti, :) P(gj, :)^T   for every owned panel c = first_panel + world * blockIdx.y, its 4 block columns
//   gj = 4 c + tj and the row tiles ti >= gj.   P holds rows >= 128 prow0_tile of the factored panel (512 columns).
struct DistUpdatePolicy {
  static constexpr bool B_KMAJOR = false;
  double* Aloc;  // owned slabs, slab q at Aloc + q * slab_elems, element (row, local col) at [row + col * ld]
  long ld, slab_elems;
  const double* P;
  long ldp;
  int prow0_tile, nt, q0, first_panel, world;
  struct Tile {
    double* C;
  };
  __device__ __forceinline__ bool setup(TileWork& w, Tile& t) const {
    const int q = q0 + blockIdx.y;
    const int c = first_panel + world * (int)blockIdx.y;
    int idx = blockIdx.x, tj = 0;
    for (; tj < OUTER_BLOCKS; ++tj) {
      const int cnt = nt - (c * OUTER_BLOCKS + tj);
      if (idx < cnt) break;
      idx -= cnt;
    }
    if (tj == OUTER_BLOCKS) return false;
    const int gj = c * OUTER_BLOCKS + tj, ti = gj + idx;
    w.A = P + (long)(ti - prow0_tile) * NB;
    w.lda = ldp;
    w.B = P + (long)(gj - prow0_tile) * NB;
    w.ldb = ldp;
    w.k_begin = 0;
    w.k_end = PANEL;
    t.C = Aloc + (long)q * slab_elems + (long)ti * NB + (long)tj * NB * ld;
    return true;
  }
  __device__ __forceinline__ void prefetch(const Tile& t) const {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int line = threadIdx.x + q * GEMM_CONSUMERS;
      prefetch_l2(t.C + (long)(line >> 3) * ld + (line & 7) * 16);
    }
  }
  __device__ __forceinline__ void epilogue(const Tile& t, Acc& acc, double*) const {
    const WarpCoord wc;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      double* cp0 = t.C + (long)wc.col(nb, 0) * ld;
      double* cp1 = cp0 + ld;
      double c0[8], c1[8];
#pragma unroll
      for (int mb = 0; mb < 8; ++mb) {
        c0[mb] = cp0[wc.row(mb)];
        c1[mb] = cp1[wc.row(mb)];
      }
#pragma unroll
      for (int mb = 0; mb < 8; ++mb) {
        cp0[wc.row(mb)] = c0[mb] - acc[mb][nb][0];
        cp1[wc.row(mb)] = c1[mb] - acc[mb][nb][1];
      }
    }
  }
};

// Backward substitution, left-looking on a column slab:  partial[cta][c] = sum_{i in chunk} L[i, 128 j + c] x[i]
// over the rows i >= 128 (j + 1);  warp per column group, lanes along the (contiguous) rows.
constexpr int BWD_CHUNK = 2048;
__global__ void __launch_bounds__(256) trsv_bwd_left_partial_kernel(const double* __restrict__ L, long ld, int j, long n,
                                                                    const double* __restrict__ x,
                                                                    double* __restrict__ partial) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long r0 = (long)(j + 1) * NB + (long)blockIdx.x * BWD_CHUNK;
  const long r1 = (r0 + BWD_CHUNK < n) ? r0 + BWD_CHUNK : n;
  for (int c = warp; c < NB; c += 8) {
    const double* Lp = L + ((long)j * NB + c) * ld;
    double s = 0.0;
    for (long i = r0 + lane; i < r1; i += 32) s = fma(Lp[i], x[i], s);
    s = warp_sum(s);
    if (lane == 0) partial[(long)blockIdx.x * NB + c] = s;
  }
}
// x_j = Linv_j^T (b_j - sum_cta partial[cta][.])
__global__ void __launch_bounds__(256) trsv_bwd_left_diag_kernel(const double* __restrict__ dinv, int j,
                                                                 const double* __restrict__ partial, int nparts,
                                                                 const double* __restrict__ b, double* __restrict__ x) {
  __shared__ double rhs[NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < NB) {
    double t = 0.0;
    for (int q = 0; q < nparts; ++q) t += partial[(long)q * NB + tid];
    rhs[tid] = b[(long)j * NB + tid] - t;
  }
  __syncthreads();
  const double* Li = dinv + (long)j * NB * NB;
  for (int r = warp; r < NB; r += 8) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = lane + 32 * q;
      s = fma(Li[c + r * NB], rhs[c], s);
    }
    s = warp_sum(s);
    if (lane == 0) x[(long)j * NB + r] = s;
  }
}

}  // namespace gprc

struct gprc_dist {
  gprc_ctx* ctx = nullptr;
  gprc::NcclApi api;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  cudaStream_t s_comm = nullptr;
  cudaEvent_t ev_packed = nullptr, ev_bcast[2] = {nullptr, nullptr}, ev_rest = nullptr, ev_begin = nullptr;
  // optional per-panel timeline of the factorisation (gprc_dist_set_timeline): 6 timing events per panel --
  // [0,1] look-ahead update + factorisation of panel p on its owner (high-priority stream), [2,3] broadcast of panel p
  // (communication stream), [4,5] trailing update with panel p (main stream) -- read back relative to the start event
  bool timeline_on = false;
  int timeline_npan = 0;
  std::vector<cudaEvent_t> timeline;  // npan * 6 (+ 1 origin)
  std::vector<double> timeline_ms;    // resolved after the fit: npan * 6, NaN where this rank recorded nothing
};
