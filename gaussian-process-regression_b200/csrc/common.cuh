// common.cuh -- context, error plumbing and the sm_100a PTX wrappers (DMMA, mbarrier, bulk async copy) shared by
// every kernel of libgprc.  FP64 on Blackwell has no tcgen05/UMMA kind (SURVEY.md section 7 "hard parts"): the tensor
// path is warp-level mma.sync m8n8k4 (SASS DMMA.8x8x4) with register accumulators, fed from shared memory that the
// TMA engine fills with cp.async.bulk (SASS UBLKCP) completing on mbarriers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/gprc.h"

namespace gprc {

constexpr int NB = 128;  // block size of every blocked algorithm; all device matrices are padded to multiples of NB

// ---------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------
extern thread_local std::string g_last_error;

inline int set_error(int code, const char* file, int line, const char* what) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s:%d: %s", file, line, what);
  g_last_error = buf;
  return code;
}

#define GPRC_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) return gprc::set_error(-2, __FILE__, __LINE__, cudaGetErrorString(e__)); \
  } while (0)

#define GPRC_CHECK(call)                 \
  do {                                   \
    int r__ = (call);                    \
    if (r__ != 0) return r__;            \
  } while (0)

#define GPRC_ARG(cond)                                                               \
  do {                                                                               \
    if (!(cond)) return gprc::set_error(-1, __FILE__, __LINE__, "bad argument: " #cond); \
  } while (0)

inline long round_up(long x, long m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------------
}  // namespace gprc

struct gprc_ctx {
  std::recursive_mutex mutex;  // held by DeviceGuard for the duration of every entry point (gprc.cu)
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream_hi = nullptr;  // high priority: the latency-bound panel factorisation (Cholesky lookahead)
  cudaEvent_t ev_start = nullptr, ev_panel = nullptr, ev_rest = nullptr;
  int sm_count = 148;
  int opt_gram_dmma = 1;
  int opt_predict_path = 0;
  int opt_ozaki_digits = 7;
  int opt_int8_auto = 1;
  int opt_int8_tile = 2;   // kernel of the INT8 pass: 2 stacked planes on cluster pairs, 1 stacked planes, 64 / 128 round 1
  int opt_int8_test_shrink = 0;
  int opt_chol_tiles = 128;  // largest n / 128 factored by the persistent tile kernel (0: never); needs d_sched
  int opt_trsv = 1;        // 1 dataflow kernel (round 2, default), 0 cooperative sweeps with grid barriers (round 1)
  int last_predict_path = 0;
  long last_predict_chunks = 0;  // chunks of test points the most recent pointwise predict was cut into
  gprc_interrupt_fn interrupt_fn = nullptr;  // polled between chunks of a long predict
  void* interrupt_user = nullptr;
  long launches = 0;
  double timers[GPRC_T_COUNT] = {0};
  // pending (start, stop, phase) events; resolved lazily in gprc_ctx_get_timers so that timing never adds a sync
  struct Pending {
    cudaEvent_t a, b;
    int phase;
  };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> event_pool;
  cudaEvent_t marks[8] = {nullptr};
  // caching device allocator: fit()/predict allocate n^2-sized buffers on every call (an optimiser calls gprc_logml
  // hundreds of times) and cudaMalloc/cudaFree are slow and synchronising; freed blocks are kept for reuse
  std::multimap<size_t, void*> pool_free;
  std::unordered_map<void*, size_t> pool_live;
  size_t pool_cached_bytes = 0;
  int* d_sched = nullptr;     // device scratch (4 KB) for the schedulers of persistent kernels: counter, error, progress[]
  long* d_info = nullptr;     // device scratch for LAPACK-style info
  double* d_scalars = nullptr;  // device scratch for small reductions (64 doubles)
  double* h_scalars = nullptr;  // pinned mirror
  long* h_info = nullptr;       // pinned
};

namespace gprc {

// scoped phase timer: records two events on the context's stream
struct PhaseTimer {
  gprc_ctx* ctx;
  int phase;
  cudaEvent_t a = nullptr, b = nullptr;
  static cudaEvent_t get_event(gprc_ctx* c) {
    if (!c->event_pool.empty()) {
      cudaEvent_t e = c->event_pool.back();
      c->event_pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  PhaseTimer(gprc_ctx* c, int ph) : ctx(c), phase(ph) {
    a = get_event(c);
    b = get_event(c);
    cudaEventRecord(a, c->stream);
  }
  ~PhaseTimer() {
    cudaEventRecord(b, ctx->stream);
    ctx->pending.push_back({a, b, phase});
  }
};

// ---------------------------------------------------------------------------------------------------------------
// device-side PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// D(8x8) += A(8x4, row) * B(4x8, col); lane holds A[lane/4][lane%4], B[lane%4][lane/4], D[lane/4][2*(lane%4)+{0,1}]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// every thread arrives once per phase, announcing the bytes its own bulk copies will deliver (0 is allowed)
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// TMA engine 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace gprc
