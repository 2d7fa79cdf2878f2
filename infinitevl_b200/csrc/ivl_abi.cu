// extern "C" entry points of libivl_b200.so (declared in include/ivl_b200.h).
// Argument validation + launch only; all math lives in the kernel files.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>

#include "../../include/ivl_b200.h"
#include "gdn_layout.cuh"

namespace ivl {
cudaError_t launch_gdn_prep(const void* q, const void* k, const void* v, const float* g, const void* beta,
                            const GdnWorkspace& ws, const GdnVarlen& vl, int num_chunks, int B, int T, int H,
                            float scale, int l2norm, int scan_ctas_per_head, int transposed, cudaStream_t stream,
                            const GdnPrepFused* fused = nullptr);
cudaError_t configure_gdn_prep();
cudaError_t launch_gdn_scan_t(const void* v, const GdnWorkspace& ws, const GdnVarlen& vl, int ntrow, int nseq, int B,
                              const void* h0, int h0_dtype, void* o, void* ht, int ht_dtype, int T, int H, int lag,
                              cudaStream_t stream);
cudaError_t launch_gdn_scan(const GdnWorkspace& ws, const GdnVarlen& vl, int ntrow, int nseq, const void* h0,
                            int h0_dtype, void* o, void* ht, int ht_dtype, int T, int H, int bv, cudaStream_t stream);
cudaError_t launch_gdn_decode_step(const void* q_in, const void* k_in, const void* v_in, const void* a_in,
                                   const void* b_in, const void* gate_in, const void* wq, const void* wk,
                                   const void* wv, const float* A_log, const float* dt_bias, const void* norm_w,
                                   void* conv_q, void* conv_k, void* conv_v, void* state, int state_dtype, void* out,
                                   int B, int H, float scale, float eps, cudaStream_t stream);
cudaError_t launch_gdn_recurrent(const void* q, const void* k, const void* v, const float* g, const void* beta,
                                 const void* h0, int h0_dtype, void* o, void* ht, int ht_dtype, int B, int T, int H,
                                 float scale, int l2norm, cudaStream_t stream);
cudaError_t launch_swa_fwd(const void* q, const long long* qs, const void* k, const long long* ks, const void* v,
                           const long long* vs, void* o, const long long* os, int B, int Tq, int Tk, int Hq, int Hkv,
                           int window, float scale, const int* ring_state, int ring_R, long long key_pos0,
                           cudaStream_t stream, const int* vt_tok0 = nullptr, const int* vt_lo = nullptr,
                           const int* vt_hi = nullptr, int vt_tiles = 0);
cudaError_t launch_peer_put(void* dst, const void* src, size_t bytes, uint32_t* flag, uint32_t value, uint32_t* counter,
                            cudaStream_t stream);
size_t swa_ring_decode_workspace_bytes(int B, int Hq, int window);
cudaError_t launch_swa_ring_decode(const void* q, const void* knew, long long kn_sb, long long kn_sh, const void* vnew,
                                   long long vn_sb, long long vn_sh, void* ring_k, void* ring_v, int* state,
                                   void* workspace, void* o, int B, int Hq, int Hkv, int R, int window, float scale,
                                   cudaStream_t stream);
cudaError_t launch_swa_ring_append(const void* k, const long long* ks, const void* v, const long long* vs, void* ring_k,
                                   void* ring_v, int* state, int B, int Tq, int Hkv, int R, cudaStream_t stream);
size_t gdn_bwd_workspace_bytes(int B, int T, int H);
cudaError_t launch_gdn_bwd(const float* qn, const float* kn, const void* v, const float* g, const float* beta,
                           const void* dout, const float* h0, const float* dht, float* dqn, float* dkn, float* dv,
                           float* dg, float* dbeta, float* dh0, float* workspace, int B, int T, int H, float scale,
                           cudaStream_t stream);
cudaError_t launch_short_conv(const void* x, const void* w, const void* cache_in, void* y, void* cache_out, int B,
                              int T, int D, int act, cudaStream_t stream, const uint8_t* left_ctx = nullptr);
cudaError_t launch_gdn_gate(const void* a, const void* b, const float* A_log, const float* dt_bias, float* g,
                            void* beta, long long n, int H, cudaStream_t stream);
cudaError_t launch_rmsnorm_gated(const void* x, const void* gate, const void* w, void* y, long long rows, float eps,
                                 cudaStream_t stream);
size_t swa_decode_workspace_bytes(int B, int Tk, int Hq);
cudaError_t launch_swa_decode(const void* q, const void* k, const long long* ks, const void* v, const long long* vs,
                              void* o, int B, int Tk, int Hq, int Hkv, int window, float scale, void* workspace,
                              cudaStream_t stream);
cudaError_t launch_mrope(void* x, const long long* xs, const void* cosr, const void* sinr, int B, int T, int Hn,
                         cudaStream_t stream);
}  // namespace ivl

namespace {
// Text of the last CUDA runtime error seen by an entry point on this thread (ivl_last_cuda_error()).
thread_local char g_last_cuda_error[256] = "";
inline bool cuda_failed(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return false;
  snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  cudaGetLastError();  // clear the non-sticky error state
  return true;
}
#define IVL_CUDA(call) do { if (cuda_failed((call), #call)) return IVL_ERR_LAUNCH; } while (0)

inline bool bad_dtype(int d) { return d != IVL_DTYPE_F32 && d != IVL_DTYPE_BF16; }
inline int check_gdn_shape(int B, int T, int H, int K, int V) {
  if (B <= 0 || T <= 0 || H <= 0 || K != ivl::GDN_K || V != ivl::GDN_V) return IVL_ERR_BAD_SHAPE;
  if (B > 65535 || H > 65535) return IVL_ERR_BAD_SHAPE;
  return IVL_OK;
}
inline float default_scale(float scale, int K) { return scale > 0.f ? scale : 1.0f / sqrtf((float)K); }

// Developer knobs (read at every call, so a test can flip them):
//   IVL_GDN_BV    value columns per scan CTA: 32, 64 or 128 (default: 32 stand-alone, 64 overlapped)
//   IVL_GDN_PIPE  0 = prep then scan on the caller's stream; 1 = overlapped (default for T >= 2048)
inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
//   IVL_GDN_TSCAN 1 = transposed scan (gdn_scan_t.cu: two CTAs per head, state and v_new as TMEM A operands),
//                 0 = row-major scan (gdn_scan.cu: 32/64/128-column slices)
//                 2 = transposed scan, lag form (shortened serial chain, N = 128 MMA chains over 128B-swizzled stacked
//                     operands; prep also emits R = Wg Kt_prev^T).  Correct and deterministic, but measured slower
//                     than form 1 on B200 (1207 vs 977 ns per chunk: tensor-memory reads and ~80-cycle TS-mode MMAs
//                     bound both; profiles/r02_summary.md), so it is not the default.
//                 3 = transposed scan, pipelined form (default): the operands of form 1, but the state / v_new hand-offs
//                     to the tensor pipe are pipelined over the contraction dimension and the output products leave
//                     the serial chain (gdn_scan_t3_kernel)
// Chunks per head of the L2 image ring of the overlapped operator (0 = off).  24 measured best at 128K tokens: DRAM
// traffic of a call 7.97 -> 3.68 GB (1.14 x algorithmic), operator 2.15 -> 2.16 ms alone, the 36-layer step 134.2 ->
// 131.2 ms (less HBM power under the 1000 W cap); 20 stalls prep (2.22 ms), 32 starts to spill (4.94 GB).
constexpr int GDN_RING_DEFAULT = 24;
inline int tscan_mode() { const int m = env_int("IVL_GDN_TSCAN", 3); return (m >= 0 && m <= 3) ? m : 3; }
inline int prep_mode(int tscan) { return tscan == 3 ? 1 : tscan; }   // form 3 reads the images of form 1
inline bool tscan() { return tscan_mode() != 0; }
inline int scan_bv(int dflt) {
  const int bv = env_int("IVL_GDN_BV", dflt);
  return (bv == 32 || bv == 64 || bv == 128) ? bv : dflt;
}

// True when an NVIDIA tool's injection library is mapped into this process (checked once).
inline bool tool_attached() {
  static int cached = -1;
  if (cached < 0) {
    cached = 0;
    const char* inj = getenv("CUDA_INJECTION64_PATH");
    if (inj && *inj) cached = 1;
    if (FILE* f = fopen("/proc/self/maps", "r")) {
      char line[1024];
      while (!cached && fgets(line, sizeof(line), f))
        if (strstr(line, "InjectionTarget") || strstr(line, "cuda-injection") || strstr(line, "libsanitizer") ||
            strstr(line, "compute-sanitizer") || strstr(line, "nsight"))
          cached = 1;
      fclose(f);
    }
  }
  return cached == 1;
}

// sm_100 check, cached per device: every kernel in this library is compiled for sm_100a only, so on any other
// device a launch would fail with "no kernel image" at best.  0 = unknown, 1 = ok, 2 = wrong architecture.
inline int check_arch() {
  static std::atomic<int> state[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return IVL_ERR_LAUNCH; }
  if (dev < 0 || dev >= 64) return IVL_ERR_ARCH;
  int s = state[dev].load(std::memory_order_relaxed);
  if (s == 0) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      cudaGetLastError();
      return IVL_ERR_LAUNCH;
    }
    s = (major == 10) ? 1 : 2;
    state[dev].store(s, std::memory_order_relaxed);
  }
  return s == 1 ? IVL_OK : IVL_ERR_ARCH;
}
#define IVL_ARCH() do { if (int e_ = check_arch()) return e_; } while (0)

// Second stream + fork/join events of the overlapped chunk operator, one set per (device, caller stream), created
// on first use (event record / wait across streams is also how a capturing stream forks, so the operator stays
// graph-safe).  Callers on different streams -- or different host threads -- never share a set.
struct ForkJoin {
  cudaStream_t aux = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool used = false;   // an overlapped call has recorded `join` outside of stream capture
};
std::mutex g_fj_mu;
std::map<std::pair<int, cudaStream_t>, ForkJoin> g_fj_sets;

// *other_busy: an overlapped call issued from ANOTHER stream of this device has not finished yet.  Its scan CTAs
// each hold a whole SM and wait for flags; a second overlapped call next to it could leave no SM for either prep
// (the caller then takes the back-to-back form, which cannot starve).
inline ForkJoin* fork_join(cudaStream_t caller, bool* other_busy) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(g_fj_mu);
  if (other_busy) {
    *other_busy = false;
    for (auto& kv : g_fj_sets)
      if (kv.first.first == dev && kv.first.second != caller && kv.second.used &&
          cudaEventQuery(kv.second.join) == cudaErrorNotReady)
        *other_busy = true;
    cudaGetLastError();
  }
  ForkJoin& f = g_fj_sets[std::make_pair(dev, caller)];
  if (!f.aux) {
    // lowest priority: when SMs come free, the block scheduler should place the scan's few whole-SM CTAs (caller's
    // stream) before prep's many small ones, so that the scan is resident early and the two really overlap
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithPriority(&f.aux, cudaStreamNonBlocking, least) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&f.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &f;
}
}  // namespace

extern "C" {

int ivl_abi_version(void) { return 3; }

int ivl_stream_init(void* stream) {
  IVL_ARCH();
  if (!fork_join(static_cast<cudaStream_t>(stream), nullptr)) {
    cuda_failed(cudaGetLastError(), "fork_join stream/event creation");
    return IVL_ERR_LAUNCH;
  }
  return IVL_OK;
}

int ivl_stream_release(void* stream) {
  int dev = 0;
  IVL_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_fj_mu);
  auto it = g_fj_sets.find(std::make_pair(dev, static_cast<cudaStream_t>(stream)));
  if (it == g_fj_sets.end()) return IVL_OK;
  if (it->second.aux) {
    cudaStreamSynchronize(it->second.aux);
    cudaEventDestroy(it->second.fork);
    cudaEventDestroy(it->second.join);
    cudaStreamDestroy(it->second.aux);
  }
  g_fj_sets.erase(it);
  cudaGetLastError();
  return IVL_OK;
}

namespace {
typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
// cuStreamWaitValue32 through the runtime's driver entry point (no link-time dependency on libcuda); nullptr when the
// driver does not offer it -- callers that would rely on it (the image ring) then do without
WaitValue32Fn wait_value32_fn() {
  static std::atomic<WaitValue32Fn> fn{nullptr};
  static std::atomic<bool> tried{false};
  WaitValue32Fn f = fn.load(std::memory_order_acquire);
  if (!f && !tried.load(std::memory_order_acquire)) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess && p) {
      f = reinterpret_cast<WaitValue32Fn>(p);
      fn.store(f, std::memory_order_release);
    } else {
      cudaGetLastError();
    }
    tried.store(true, std::memory_order_release);
  }
  return f;
}
}  // namespace

int ivl_stream_wait_value32(void* stream, const void* addr, uint32_t value) {
  if (!addr || (reinterpret_cast<uintptr_t>(addr) & 3)) return IVL_ERR_NULL;
  WaitValue32Fn f = wait_value32_fn();
  if (!f) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "cuStreamWaitValue32 is not available from this driver");
    return IVL_ERR_LAUNCH;
  }
  const CUresult r = f(static_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(addr), value, CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "cuStreamWaitValue32: CUresult %d", (int)r);
    return IVL_ERR_LAUNCH;
  }
  return IVL_OK;
}

int ivl_peer_put(void* dst, const void* src, size_t bytes, uint32_t* flag, uint32_t value, uint32_t* counter,
                 void* stream) {
  if (!flag || !counter || (bytes && (!dst || !src))) return IVL_ERR_NULL;
  if ((bytes & 15) || ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) ||
      (reinterpret_cast<uintptr_t>(flag) & 3) || (reinterpret_cast<uintptr_t>(counter) & 3))
    return IVL_ERR_BAD_SHAPE;
  IVL_CUDA(ivl::launch_peer_put(dst, src, bytes, flag, value, counter, static_cast<cudaStream_t>(stream)));
  return IVL_OK;
}

int ivl_ipc_export(const void* ptr, void* handle64_out, uint64_t* offset_out) {
  if (!ptr || !handle64_out || !offset_out) return IVL_ERR_NULL;
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cuda_failed(cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q), "cudaGetDriverEntryPoint") ||
      q != cudaDriverEntryPointSuccess || !p)
    return IVL_ERR_LAUNCH;
  CUdeviceptr base = 0;
  size_t size = 0;
  const CUresult r = reinterpret_cast<RangeFn>(p)(&base, &size, reinterpret_cast<CUdeviceptr>(ptr));
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "cuMemGetAddressRange: CUresult %d", (int)r);
    return IVL_ERR_LAUNCH;
  }
  cudaIpcMemHandle_t h;
  IVL_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  memcpy(handle64_out, &h, sizeof(h));
  *offset_out = (uint64_t)(reinterpret_cast<CUdeviceptr>(ptr) - base);
  return IVL_OK;
}

int ivl_ipc_open(const void* handle64, void** base_ptr) {
  if (!handle64 || !base_ptr) return IVL_ERR_NULL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  IVL_CUDA(cudaIpcOpenMemHandle(base_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return IVL_OK;
}

int ivl_ipc_close(void* base_ptr) {
  if (!base_ptr) return IVL_ERR_NULL;
  IVL_CUDA(cudaIpcCloseMemHandle(base_ptr));
  return IVL_OK;
}

const char* ivl_last_cuda_error(void) { return g_last_cuda_error; }

const char* ivl_strerror(int code) {
  switch (code) {
    case IVL_OK: return "ok";
    case IVL_ERR_BAD_SHAPE: return "unsupported shape (need K=128, V=256, positive sizes)";
    case IVL_ERR_NULL: return "required pointer is NULL";
    case IVL_ERR_WORKSPACE: return "workspace too small or misaligned";
    case IVL_ERR_DTYPE: return "unknown dtype code";
    case IVL_ERR_LAUNCH: return "CUDA launch failed";
    case IVL_ERR_ARCH: return "device is not sm_100";
    default: return "unknown error";
  }
}

size_t ivl_gdn_chunk_workspace_bytes(int B, int T, int H) {
  if (B <= 0 || T <= 0 || H <= 0) return 0;
  return ivl::gdn_workspace_bytes(B, T, H);
}

int ivl_gdn_chunk_prep(const void* q, const void* k, const void* v, const float* g, const void* beta, int B, int T,
                       int H, float scale, int l2norm_qk, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_gdn_shape(B, T, H, ivl::GDN_K, ivl::GDN_V)) return e;
  if (!q || !k || !v || !g || !beta || !workspace) return IVL_ERR_NULL;
  if (workspace_bytes < ivl::gdn_workspace_bytes(B, T, H) || (reinterpret_cast<uintptr_t>(workspace) & 1023))
    return IVL_ERR_WORKSPACE;
  IVL_ARCH();
  ivl::GdnWorkspace ws = ivl::gdn_carve(workspace, B, T, H, /*ring=*/0);  // one slot per chunk
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  IVL_CUDA(cudaMemsetAsync(ws.ready, 0, ivl::gdn_sync_bytes(B, T, H), st));
  IVL_CUDA(ivl::launch_gdn_prep(q, k, v, g, beta, ws, ivl::GdnVarlen{}, ivl::gdn_num_chunks(T), B, T, H,
                                default_scale(scale, ivl::GDN_K), l2norm_qk, 0, prep_mode(tscan_mode()), st));
  return IVL_OK;
}

int ivl_gdn_chunk_scan(const void* v, const void* h0, int h0_dtype, void* o, void* ht, int ht_dtype, int B, int T, int H,
                       void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_gdn_shape(B, T, H, ivl::GDN_K, ivl::GDN_V)) return e;
  if (!v || !o || !workspace) return IVL_ERR_NULL;
  if ((h0 && bad_dtype(h0_dtype)) || (ht && bad_dtype(ht_dtype))) return IVL_ERR_DTYPE;
  if (workspace_bytes < ivl::gdn_workspace_bytes(B, T, H) || (reinterpret_cast<uintptr_t>(workspace) & 1023))
    return IVL_ERR_WORKSPACE;
  IVL_ARCH();
  ivl::GdnWorkspace ws = ivl::gdn_carve(workspace, B, T, H, /*ring=*/0);
  if (tscan())
    IVL_CUDA(ivl::launch_gdn_scan_t(v, ws, ivl::GdnVarlen{}, ivl::gdn_num_chunks(T), B, B, h0, h0_dtype, o, ht, ht_dtype,
                                    T, H, tscan_mode(), static_cast<cudaStream_t>(stream)));
  else
    IVL_CUDA(ivl::launch_gdn_scan(ws, ivl::GdnVarlen{}, ivl::gdn_num_chunks(T), B, h0, h0_dtype, o, ht, ht_dtype, T, H,
                                  scan_bv(32), static_cast<cudaStream_t>(stream)));
  return IVL_OK;
}

namespace {
// Shared body of the dense and the packed (variable-length) chunk operator.  nseq: batch rows (dense) or sequences
// (packed, B = 1); num_chunks: chunks per workspace row.
int gdn_chunk_fwd_impl(const void* q, const void* k, const void* v, const float* g, const void* beta, const void* h0,
                       int h0_dtype, void* o, void* ht, int ht_dtype, int B, int T, int H, float scale, int l2norm_qk,
                       const ivl::GdnVarlen& vl, int num_chunks, int nseq, void* workspace, size_t workspace_bytes,
                       void* stream, const ivl::GdnPrepFused* fused = nullptr) {
  if (!q || !k || !v || !o || !workspace) return IVL_ERR_NULL;
  if (!fused && (!g || !beta)) return IVL_ERR_NULL;
  if (fused && tscan_mode() != 3 && tscan_mode() != 1) return IVL_ERR_BAD_SHAPE;   // images of prep mode 1 only
  if ((h0 && bad_dtype(h0_dtype)) || (ht && bad_dtype(ht_dtype))) return IVL_ERR_DTYPE;
  const int T_ws = num_chunks * ivl::GDN_C;   // the workspace is sized by chunks
  if (workspace_bytes < ivl::gdn_workspace_bytes(B, T_ws, H) || (reinterpret_cast<uintptr_t>(workspace) & 1023))
    return IVL_ERR_WORKSPACE;
  IVL_ARCH();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float sc = default_scale(scale, ivl::GDN_K);
  // The first call on a device runs the two kernels back to back: a first launch may have to load the kernel
  // or grow the context's local-memory pool, both of which wait for running kernels -- and in the overlapped
  // form the running scan waits for prep.  (Not needed under stream capture: nothing runs at capture time.)
  static std::atomic<bool> warmed[64];
  int dev = 0;
  IVL_CUDA(cudaGetDevice(&dev));
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  IVL_CUDA(cudaStreamIsCapturing(st, &cap));
  // exchange(): exactly one of several racing first callers sees `false`; the others may overlap while that one
  // is still loading the kernels, which is why the warm-up call below also synchronises nothing but merely runs
  // back to back -- a racing overlapped call then finds the module loaded or blocks in its own launch, not in a
  // running scan
  const bool first = dev >= 0 && dev < 64 && cap == cudaStreamCaptureStatusNone &&
                     !warmed[dev].load(std::memory_order_acquire) && !warmed[dev].exchange(true);
  // Profilers and sanitizers (ncu, compute-sanitizer: their injection library is mapped into the process) and
  // CUDA_LAUNCH_BLOCKING=1 run one kernel at a time; a scan that waits for a prep that cannot start would only
  // hit its time-out trap, so those runs get the back-to-back form unless IVL_GDN_PIPE is set explicitly.
  const bool serialised = tool_attached() || env_int("CUDA_LAUNCH_BLOCKING", 0) != 0;
  const int overlap_default = (num_chunks >= 32 && !serialised) ? 1 : 0;
  // The overlapped form needs the scan's CTAs resident AND at least half of the SMs left for prep: scan CTAs
  // that fill the GPU would spin on flags that prep could then never publish.  Big batches take wider slices
  // or, failing that, the back-to-back form.
  int sms = 0;
  IVL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int trm = tscan_mode();
  const bool tr = trm != 0;
  int bv_overlap = tr ? 128 : scan_bv(64);   // the transposed scan always owns 128 value columns per CTA
  if ((long long)nseq * H * (ivl::GDN_V / bv_overlap) > sms / 2) bv_overlap = 128;
  const bool fits = (long long)nseq * H * (ivl::GDN_V / bv_overlap) <= sms / 2;
  if (first || !fits || env_int("IVL_GDN_PIPE", overlap_default) == 0) {
    ivl::GdnWorkspace ws = ivl::gdn_carve(workspace, B, T_ws, H, /*ring=*/0);
    IVL_CUDA(cudaMemsetAsync(ws.ready, 0, ivl::gdn_sync_bytes(B, T_ws, H), st));
    IVL_CUDA(ivl::launch_gdn_prep(q, k, v, g, beta, ws, vl, num_chunks, B, T, H, sc, l2norm_qk, 0, prep_mode(trm), st, fused));
    if (tr)
      IVL_CUDA(ivl::launch_gdn_scan_t(v, ws, vl, num_chunks, nseq, B, h0, h0_dtype, o, ht, ht_dtype, T, H, trm, st));
    else
      IVL_CUDA(ivl::launch_gdn_scan(ws, vl, num_chunks, nseq, h0, h0_dtype, o, ht, ht_dtype, T, H, scan_bv(32), st));
    return IVL_OK;
  }
  // Overlapped form.  The scan goes first, on the caller's stream, with 64-column slices: its few CTAs (64 at
  // B = 1, H = 16) take their SMs and then follow the ready flags; prep runs on the second stream on the SMs
  // that are left and publishes chunk after chunk, so the operand images are consumed while they are still
  // in L2.  At 128K tokens both sides then take ~2.3 ms (scan alone on 64 SMs 2.32 ms, prep alone on 84 SMs
  // 1.34 x 148 / 84 = 2.37 ms), against 1.34 + 1.59 ms back to back.
  bool other_busy = false;
  ForkJoin* fj = fork_join(st, cap == cudaStreamCaptureStatusNone ? &other_busy : nullptr);
  if (!fj) { cuda_failed(cudaGetLastError(), "fork_join stream/event creation"); return IVL_ERR_LAUNCH; }
  if (other_busy) {
    ivl::GdnWorkspace ws = ivl::gdn_carve(workspace, B, T_ws, H, /*ring=*/0);
    IVL_CUDA(cudaMemsetAsync(ws.ready, 0, ivl::gdn_sync_bytes(B, T_ws, H), st));
    IVL_CUDA(ivl::launch_gdn_prep(q, k, v, g, beta, ws, vl, num_chunks, B, T, H, sc, l2norm_qk, 0, prep_mode(trm), st, fused));
    if (tr)
      IVL_CUDA(ivl::launch_gdn_scan_t(v, ws, vl, num_chunks, nseq, B, h0, h0_dtype, o, ht, ht_dtype, T, H, trm, st));
    else
      IVL_CUDA(ivl::launch_gdn_scan(ws, vl, num_chunks, nseq, h0, h0_dtype, o, ht, ht_dtype, T, H, scan_bv(32), st));
    return IVL_OK;
  }
  // Image ring (gdn_layout.cuh): prep reuses a short ring of chunk slots per head and waits for the scan's progress
  // before it overwrites one, so the operand images are produced and consumed inside L2 and never travel to HBM
  // (the round trip is 2 x 2.1 GB of the call's 7.97 GB of DRAM traffic at 128K tokens).  Round 1 found two problems:
  // prep needs ~16-22 chunks in flight per head, so a ring shorter than that makes the two kernels wait on each other,
  // and -- a real hang -- prep CTAs could become resident before the scan's (which need a whole SM each); once every
  // SM held a prep CTA waiting for scan progress the scan could never start.  Residency is now established first:
  // every scan CTA checks in (ws.checkin) and the prep launch waits for the full count with a stream memory operation
  // on the helper stream.  After that the oldest unfinished prep CTA only ever waits for a scan that only waits for
  // older chunks: no cycle.  Transposed scans, dense batches, not under stream capture (IVL_GDN_RING chunks; 0 = off).
  const int bv = bv_overlap;
  int ring = env_int("IVL_GDN_RING", GDN_RING_DEFAULT);
  if (!tr || vl.chunk_tok0 != nullptr || cap != cudaStreamCaptureStatusNone || ring < 0 || ring >= num_chunks) ring = 0;
  if (ring > 0 && ring < 4) ring = 4;
  if (ring > 0 && !tool_attached() && wait_value32_fn() == nullptr) ring = 0;   // no handshake, no ring
  ivl::GdnWorkspace ws = ivl::gdn_carve(workspace, B, T_ws, H, ring);
  IVL_CUDA(ivl::configure_gdn_prep());  // prep must be loaded before a scan that waits for it is running
  IVL_CUDA(cudaMemsetAsync(ws.ready, 0, ivl::gdn_sync_bytes(B, T_ws, H), st));
  IVL_CUDA(cudaEventRecord(fj->fork, st));
  IVL_CUDA(cudaStreamWaitEvent(fj->aux, fj->fork, 0));
  if (tr)
    IVL_CUDA(ivl::launch_gdn_scan_t(v, ws, vl, num_chunks, nseq, B, h0, h0_dtype, o, ht, ht_dtype, T, H, trm, st));
  else
    IVL_CUDA(ivl::launch_gdn_scan(ws, vl, num_chunks, nseq, h0, h0_dtype, o, ht, ht_dtype, T, H, bv, st));
  if (ring > 0) {
    const int scan_ctas = nseq * H * (ivl::GDN_V / bv);
    // (ncu range replay refuses stream memory operations; a profiled run of a single call goes without the handshake --
    // the hazard it removes needs a previous call still draining -- and both wait loops trap instead of hanging)
    // (a failing wait is reported but must not stop here: the scan is already waiting for prep's flags)
    int wait_err = IVL_OK;
    if (!tool_attached()) wait_err = ivl_stream_wait_value32(fj->aux, ws.checkin, (uint32_t)scan_ctas);
    (void)wait_err;
  }
  // (should prep fail to launch, the scan traps after its time-out instead of hanging)
  IVL_CUDA(ivl::launch_gdn_prep(q, k, v, g, beta, ws, vl, num_chunks, B, T, H, sc, l2norm_qk, ivl::GDN_V / bv,
                                prep_mode(trm), fj->aux, fused));
  IVL_CUDA(cudaEventRecord(fj->join, fj->aux));
  IVL_CUDA(cudaStreamWaitEvent(st, fj->join, 0));
  if (cap == cudaStreamCaptureStatusNone) {
    std::lock_guard<std::mutex> lock(g_fj_mu);
    fj->used = true;
  }
  return IVL_OK;
}
}  // namespace

int ivl_gdn_chunk_fwd(const void* q, const void* k, const void* v, const float* g, const void* beta, const void* h0,
                      int h0_dtype, void* o, void* ht, int ht_dtype, int B, int T, int H, int K, int V, float scale,
                      int l2norm_qk, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_gdn_shape(B, T, H, K, V)) return e;
  return gdn_chunk_fwd_impl(q, k, v, g, beta, h0, h0_dtype, o, ht, ht_dtype, B, T, H, scale, l2norm_qk, ivl::GdnVarlen{},
                            ivl::gdn_num_chunks(T), B, workspace, workspace_bytes, stream);
}

int ivl_gdn_chunk_fwd_fused(const void* xq, const void* xk, const void* v, const void* a, const void* b,
                            const void* conv_wq, const void* conv_wk, const void* conv_q_in, const void* conv_k_in,
                            void* conv_q_out, void* conv_k_out, const float* A_log, const float* dt_bias, const void* h0,
                            int h0_dtype, void* o, void* ht, int ht_dtype, int B, int T, int H, int K, int V, float scale,
                            void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_gdn_shape(B, T, H, K, V)) return e;
  if (!xq || !xk || !a || !b || !conv_wq || !conv_wk || !A_log || !dt_bias) return IVL_ERR_NULL;
  if ((conv_q_out == nullptr) != (conv_k_out == nullptr)) return IVL_ERR_NULL;
  ivl::GdnPrepFused fz;
  fz.wq = conv_wq; fz.wk = conv_wk; fz.cq_in = conv_q_in; fz.ck_in = conv_k_in; fz.cq_out = conv_q_out;
  fz.ck_out = conv_k_out; fz.a = a; fz.b = b; fz.A_log = A_log; fz.dt_bias = dt_bias;
  return gdn_chunk_fwd_impl(xq, xk, v, nullptr, nullptr, h0, h0_dtype, o, ht, ht_dtype, B, T, H, scale, /*l2norm=*/1,
                            ivl::GdnVarlen{}, ivl::gdn_num_chunks(T), B, workspace, workspace_bytes, stream, &fz);
}

int ivl_gdn_chunk_fwd_varlen(const void* q, const void* k, const void* v, const float* g, const void* beta,
                             const void* h0, int h0_dtype, void* o, void* ht, int ht_dtype, int T, int H, int K, int V,
                             float scale, int l2norm_qk, const int32_t* chunk_tok0, const int32_t* chunk_valid,
                             int num_chunks, const int32_t* seq_chunk_begin, int num_seqs, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (int e = check_gdn_shape(1, T, H, K, V)) return e;
  if (num_chunks <= 0 || num_seqs <= 0 || num_seqs > 65535) return IVL_ERR_BAD_SHAPE;
  if (!chunk_tok0 || !chunk_valid || !seq_chunk_begin) return IVL_ERR_NULL;
  ivl::GdnVarlen vl{chunk_tok0, chunk_valid, seq_chunk_begin, num_seqs};
  return gdn_chunk_fwd_impl(q, k, v, g, beta, h0, h0_dtype, o, ht, ht_dtype, 1, T, H, scale, l2norm_qk, vl, num_chunks,
                            num_seqs, workspace, workspace_bytes, stream);
}

int ivl_gdn_recurrent_fwd(const void* q, const void* k, const void* v, const float* g, const void* beta,
                          const void* h0, int h0_dtype, void* o, void* ht, int ht_dtype, int B, int T, int H, int K,
                          int V, float scale, int l2norm_qk, void* stream) {
  if (int e = check_gdn_shape(B, T, H, K, V)) return e;
  if (!q || !k || !v || !g || !beta || !o) return IVL_ERR_NULL;
  if ((h0 && bad_dtype(h0_dtype)) || (ht && bad_dtype(ht_dtype))) return IVL_ERR_DTYPE;
  IVL_ARCH();
  cudaError_t e = ivl::launch_gdn_recurrent(q, k, v, g, beta, h0, h0 ? h0_dtype : 0, o, ht, ht ? ht_dtype : 0, B, T, H,
                                            default_scale(scale, K), l2norm_qk, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? IVL_OK : IVL_ERR_LAUNCH;
}

size_t ivl_gdn_bwd_workspace_bytes(int B, int T, int H) {
  if (B <= 0 || T <= 0 || H <= 0) return 0;
  return ivl::gdn_bwd_workspace_bytes(B, T, H);
}

int ivl_gdn_bwd(const float* qn, const float* kn, const void* v, const float* g, const float* beta, const void* d_o,
                const float* h0, const float* d_ht, float* d_qn, float* d_kn, float* d_v, float* d_g, float* d_beta,
                float* d_h0, int B, int T, int H, int K, int V, float scale, void* workspace, size_t workspace_bytes,
                void* stream) {
  if (int e = check_gdn_shape(B, T, H, K, V)) return e;
  if (!qn || !kn || !v || !g || !beta || !d_o || !d_qn || !d_kn || !d_v || !d_g || !d_beta || !workspace)
    return IVL_ERR_NULL;
  if (workspace_bytes < ivl::gdn_bwd_workspace_bytes(B, T, H) || (reinterpret_cast<uintptr_t>(workspace) & 15))
    return IVL_ERR_WORKSPACE;
  IVL_ARCH();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = (size_t)B * T * H;
  // dq, dk, dg, dbeta are accumulated across the value slices with atomics
  IVL_CUDA(cudaMemsetAsync(d_qn, 0, n * K * sizeof(float), st));
  IVL_CUDA(cudaMemsetAsync(d_kn, 0, n * K * sizeof(float), st));
  IVL_CUDA(cudaMemsetAsync(d_g, 0, n * sizeof(float), st));
  IVL_CUDA(cudaMemsetAsync(d_beta, 0, n * sizeof(float), st));
  IVL_CUDA(ivl::launch_gdn_bwd(qn, kn, v, g, beta, d_o, h0, d_ht, d_qn, d_kn, d_v, d_g, d_beta, d_h0,
                               static_cast<float*>(workspace), B, T, H, default_scale(scale, K), st));
  return IVL_OK;
}

int ivl_gdn_decode_step(const void* q_in, const void* k_in, const void* v_in, const void* a_in, const void* b_in,
                        const void* gate_in, const void* conv_weight_q, const void* conv_weight_k,
                        const void* conv_weight_v, const float* A_log, const float* dt_bias, const void* norm_weight,
                        void* conv_state_q, void* conv_state_k, void* conv_state_v, void* state, int state_dtype,
                        void* out, int B, int H, int K, int V, float scale, float norm_eps, void* stream) {
  if (int e = check_gdn_shape(B, 1, H, K, V)) return e;
  if (!q_in || !k_in || !v_in || !a_in || !b_in || !gate_in || !conv_weight_q || !conv_weight_k || !conv_weight_v ||
      !A_log || !dt_bias || !norm_weight || !conv_state_q || !conv_state_k || !conv_state_v || !state || !out)
    return IVL_ERR_NULL;
  if (bad_dtype(state_dtype)) return IVL_ERR_DTYPE;
  IVL_ARCH();
  IVL_CUDA(ivl::launch_gdn_decode_step(q_in, k_in, v_in, a_in, b_in, gate_in, conv_weight_q, conv_weight_k,
                                       conv_weight_v, A_log, dt_bias, norm_weight, conv_state_q, conv_state_k,
                                       conv_state_v, state, state_dtype, out, B, H, default_scale(scale, K), norm_eps,
                                       static_cast<cudaStream_t>(stream)));
  return IVL_OK;
}

int ivl_swa_fwd(const void* q, const int64_t* q_strides, const void* k, const int64_t* k_strides, const void* v,
                const int64_t* v_strides, void* o, const int64_t* o_strides, int B, int Tq, int Tk, int Hq, int Hkv,
                int D, int window, float scale, void* stream) {
  return ivl_swa_fwd_pos(q, q_strides, k, k_strides, v, v_strides, o, o_strides, B, Tq, Tk, Hq, Hkv, D, window, scale,
                         0, stream);
}

int ivl_swa_fwd_pos(const void* q, const int64_t* q_strides, const void* k, const int64_t* k_strides, const void* v,
                    const int64_t* v_strides, void* o, const int64_t* o_strides, int B, int Tq, int Tk, int Hq, int Hkv,
                    int D, int window, float scale, int64_t key_pos0, void* stream) {
  if (key_pos0 < 0) return IVL_ERR_BAD_SHAPE;
  if (B <= 0 || Tq <= 0 || Tk < Tq || Hq <= 0 || Hkv <= 0 || Hq % Hkv != 0 || D != 128 || B > 65535)
    return IVL_ERR_BAD_SHAPE;
  if ((Tq + 127) / 128 > 65535) return IVL_ERR_BAD_SHAPE;
  if (!q || !k || !v || !o || !q_strides || !k_strides || !v_strides || !o_strides) return IVL_ERR_NULL;
  long long qs[3], ks[3], vs[3], os[3];
  for (int i = 0; i < 3; ++i) {
    qs[i] = q_strides[i]; ks[i] = k_strides[i]; vs[i] = v_strides[i]; os[i] = o_strides[i];
    if ((qs[i] | ks[i] | vs[i] | os[i]) & 7) return IVL_ERR_BAD_SHAPE;  // TMA / 16-byte vector alignment
  }
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(o)) & 15)
    return IVL_ERR_BAD_SHAPE;
  const float sc = scale > 0.f ? scale : 1.0f / sqrtf((float)D);
  IVL_ARCH();
  cudaError_t e = ivl::launch_swa_fwd(q, qs, k, ks, v, vs, o, os, B, Tq, Tk, Hq, Hkv, window, sc, nullptr, 0,
                                      (long long)key_pos0, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? IVL_OK : IVL_ERR_LAUNCH;
}

int ivl_swa_fwd_varlen(const void* q, const int64_t* q_strides, const void* k, const int64_t* k_strides, const void* v,
                       const int64_t* v_strides, void* o, const int64_t* o_strides, int T, int Hq, int Hkv, int D,
                       int window, float scale, const int32_t* tile_tok0, const int32_t* tile_seq_lo,
                       const int32_t* tile_seq_hi, int num_tiles, void* stream) {
  if (T <= 0 || Hq <= 0 || Hkv <= 0 || Hq % Hkv != 0 || D != 128 || num_tiles <= 0 || num_tiles > 65535)
    return IVL_ERR_BAD_SHAPE;
  if (!q || !k || !v || !o || !q_strides || !k_strides || !v_strides || !o_strides || !tile_tok0 || !tile_seq_lo ||
      !tile_seq_hi)
    return IVL_ERR_NULL;
  long long qs[3], ks[3], vs[3], os[3];
  for (int i = 0; i < 3; ++i) {
    qs[i] = q_strides[i]; ks[i] = k_strides[i]; vs[i] = v_strides[i]; os[i] = o_strides[i];
    if ((qs[i] | ks[i] | vs[i] | os[i]) & 7) return IVL_ERR_BAD_SHAPE;
  }
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(o)) & 15)
    return IVL_ERR_BAD_SHAPE;
  const float sc = scale > 0.f ? scale : 1.0f / sqrtf((float)D);
  IVL_ARCH();
  cudaError_t e = ivl::launch_swa_fwd(q, qs, k, ks, v, vs, o, os, 1, T, T, Hq, Hkv, window, sc, nullptr, 0, 0,
                                      static_cast<cudaStream_t>(stream), tile_tok0, tile_seq_lo, tile_seq_hi, num_tiles);
  return e == cudaSuccess ? IVL_OK : IVL_ERR_LAUNCH;
}

// ---- ring-buffer window cache -------------------------------------------------------------------------------
size_t ivl_swa_ring_workspace_bytes(int B, int Hq, int window) {
  if (B <= 0 || Hq <= 0 || window <= 0) return 0;
  return ivl::swa_ring_decode_workspace_bytes(B, Hq, window);
}

size_t ivl_swa_ring_state_bytes(int B, int Hkv) {
  if (B <= 0 || Hkv <= 0) return 0;
  return (size_t)(2 + B * Hkv) * sizeof(int32_t);
}

static inline int check_ring(int B, int Hq, int Hkv, int D, int window, int R) {
  if (B <= 0 || Hq <= 0 || Hkv <= 0 || Hq % Hkv != 0 || Hq / Hkv > 8 || D != 128 || B > 65535) return IVL_ERR_BAD_SHAPE;
  if (window <= 0 || R < window) return IVL_ERR_BAD_SHAPE;
  return IVL_OK;
}

int ivl_swa_ring_append(const void* k, const int64_t* k_strides, const void* v, const int64_t* v_strides, void* ring_k,
                        void* ring_v, int32_t* state, int B, int Tq, int Hkv, int D, int R, void* stream) {
  if (B <= 0 || Tq <= 0 || Hkv <= 0 || D != 128 || R <= 0) return IVL_ERR_BAD_SHAPE;
  if (!k || !v || !k_strides || !v_strides || !ring_k || !ring_v || !state) return IVL_ERR_NULL;
  long long ks[3], vs[3];
  for (int i = 0; i < 3; ++i) {
    ks[i] = k_strides[i]; vs[i] = v_strides[i];
    if ((ks[i] | vs[i]) & 7) return IVL_ERR_BAD_SHAPE;
  }
  IVL_ARCH();
  IVL_CUDA(ivl::launch_swa_ring_append(k, ks, v, vs, ring_k, ring_v, state, B, Tq, Hkv, R, static_cast<cudaStream_t>(stream)));
  return IVL_OK;
}

int ivl_swa_ring_decode(const void* q, const void* k_new, const int64_t* k_new_strides, const void* v_new,
                        const int64_t* v_new_strides, void* ring_k, void* ring_v, int32_t* state, void* o, int B, int Hq,
                        int Hkv, int D, int window, int R, float scale, void* workspace, size_t workspace_bytes,
                        void* stream) {
  if (int e = check_ring(B, Hq, Hkv, D, window, R)) return e;
  if (!q || !k_new || !v_new || !k_new_strides || !v_new_strides || !ring_k || !ring_v || !state || !o || !workspace)
    return IVL_ERR_NULL;
  if (workspace_bytes < ivl::swa_ring_decode_workspace_bytes(B, Hq, window)) return IVL_ERR_WORKSPACE;
  if (window > 8192) return IVL_ERR_BAD_SHAPE;   // the fused combine holds at most 64 slices of 128 keys
  if ((k_new_strides[0] | k_new_strides[1] | v_new_strides[0] | v_new_strides[1]) & 7) return IVL_ERR_BAD_SHAPE;
  IVL_ARCH();
  const float sc = scale > 0.f ? scale : 1.0f / sqrtf((float)D);
  IVL_CUDA(ivl::launch_swa_ring_decode(q, k_new, k_new_strides[0], k_new_strides[1], v_new, v_new_strides[0],
                                       v_new_strides[1], ring_k, ring_v, state, workspace, o, B, Hq, Hkv, R, window, sc,
                                       static_cast<cudaStream_t>(stream)));
  return IVL_OK;
}

int ivl_swa_ring_fwd(const void* q, const int64_t* q_strides, const void* ring_k, const void* ring_v,
                     const int32_t* state, void* o, const int64_t* o_strides, int B, int Tq, int Hq, int Hkv, int D,
                     int window, int R, float scale, void* stream) {
  if (int e = check_ring(B, Hq, Hkv, D, window, R)) return e;
  if (Tq <= 0 || (Tq + 127) / 128 > 65535 || R < window - 1 + Tq) return IVL_ERR_BAD_SHAPE;
  if (!q || !q_strides || !ring_k || !ring_v || !state || !o || !o_strides) return IVL_ERR_NULL;
  long long qs[3], os[3];
  for (int i = 0; i < 3; ++i) {
    qs[i] = q_strides[i]; os[i] = o_strides[i];
    if ((qs[i] | os[i]) & 7) return IVL_ERR_BAD_SHAPE;
  }
  const long long rs[3] = {2ll * R * Hkv * D, (long long)Hkv * D, (long long)D};
  IVL_ARCH();
  const float sc = scale > 0.f ? scale : 1.0f / sqrtf((float)D);
  cudaError_t e = ivl::launch_swa_fwd(q, qs, ring_k, rs, ring_v, rs, o, os, B, Tq, 2 * R, Hq, Hkv, window, sc, state, R,
                                      0, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? IVL_OK : IVL_ERR_LAUNCH;
}

size_t ivl_swa_decode_workspace_bytes(int B, int Tk, int Hq) {
  if (B <= 0 || Tk <= 0 || Hq <= 0) return 0;
  return ivl::swa_decode_workspace_bytes(B, Tk, Hq);
}

int ivl_swa_decode_fwd(const void* q, const void* k, const int64_t* k_strides, const void* v, const int64_t* v_strides,
                       void* o, int B, int Tk, int Hq, int Hkv, int D, int window, float scale, void* workspace,
                       size_t workspace_bytes, void* stream) {
  if (B <= 0 || Tk <= 0 || Hq <= 0 || Hkv <= 0 || Hq % Hkv != 0 || Hq / Hkv > 8 || D != 128 || B > 65535)
    return IVL_ERR_BAD_SHAPE;
  if (!q || !k || !v || !o || !k_strides || !v_strides || !workspace) return IVL_ERR_NULL;
  if (workspace_bytes < ivl::swa_decode_workspace_bytes(B, Tk, Hq)) return IVL_ERR_WORKSPACE;
  long long ks[3], vs[3];
  for (int i = 0; i < 3; ++i) {
    ks[i] = k_strides[i]; vs[i] = v_strides[i];
    if ((ks[i] | vs[i]) & 7) return IVL_ERR_BAD_SHAPE;
  }
  const float sc = scale > 0.f ? scale : 1.0f / sqrtf((float)D);
  IVL_ARCH();
  cudaError_t e = ivl::launch_swa_decode(q, k, ks, v, vs, o, B, Tk, Hq, Hkv, window, sc, workspace,
                                         static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? IVL_OK : IVL_ERR_LAUNCH;
}

int ivl_short_conv_fwd(const void* x, const void* w, const void* cache_in, void* y, void* cache_out, int B, int T,
                       int D, int activation_silu, void* stream) {
  if (B <= 0 || T <= 0 || D <= 0 || (D & 7) || B > 65535 || (T + 31) / 32 > 65535 * 32) return IVL_ERR_BAD_SHAPE;
  if (!x || !w || !y) return IVL_ERR_NULL;
  if (cache_out && cache_out == cache_in) return IVL_ERR_BAD_SHAPE;
  if ((T + 31) / 32 > 65535) return IVL_ERR_BAD_SHAPE;
  IVL_ARCH();
  cudaError_t e = ivl::launch_short_conv(x, w, cache_in, y, cache_out, B, T, D, activation_silu,
                                         static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? IVL_OK : IVL_ERR_LAUNCH;
}

int ivl_short_conv_fwd_varlen(const void* x, const void* w, void* y, const uint8_t* left_ctx, int T, int D,
                              int activation_silu, void* stream) {
  if (T <= 0 || D <= 0 || (D & 7) || (T + 31) / 32 > 65535) return IVL_ERR_BAD_SHAPE;
  if (!x || !w || !y || !left_ctx) return IVL_ERR_NULL;
  IVL_ARCH();
  cudaError_t e = ivl::launch_short_conv(x, w, nullptr, y, nullptr, 1, T, D, activation_silu,
                                         static_cast<cudaStream_t>(stream), left_ctx);
  return e == cudaSuccess ? IVL_OK : IVL_ERR_LAUNCH;
}

int ivl_gdn_gate_fwd(const void* a, const void* b, const float* A_log, const float* dt_bias, float* g, void* beta,
                     int64_t n_tokens, int H, void* stream) {
  if (n_tokens <= 0 || H <= 0) return IVL_ERR_BAD_SHAPE;
  if (!a || !b || !A_log || !dt_bias || !g || !beta) return IVL_ERR_NULL;
  IVL_ARCH();
  cudaError_t e = ivl::launch_gdn_gate(a, b, A_log, dt_bias, g, beta, (long long)n_tokens * H, H,
                                       static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? IVL_OK : IVL_ERR_LAUNCH;
}

int ivl_rmsnorm_gated_fwd(const void* x, const void* gate, const void* w, void* y, int64_t rows, int dim, float eps,
                          void* stream) {
  if (rows <= 0 || dim != 256) return IVL_ERR_BAD_SHAPE;
  if (!x || !gate || !w || !y) return IVL_ERR_NULL;
  IVL_ARCH();
  cudaError_t e = ivl::launch_rmsnorm_gated(x, gate, w, y, rows, eps, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? IVL_OK : IVL_ERR_LAUNCH;
}

int ivl_mrope_apply(void* x, const int64_t* x_strides, const void* cos, const void* sin, int B, int T, int Hn, int D,
                    void* stream) {
  if (B <= 0 || T <= 0 || Hn <= 0 || D != 128 || B > 65535) return IVL_ERR_BAD_SHAPE;
  if (!x || !x_strides || !cos || !sin) return IVL_ERR_NULL;
  long long xs[3] = {x_strides[0], x_strides[1], x_strides[2]};
  if ((xs[0] | xs[1] | xs[2]) & 7) return IVL_ERR_BAD_SHAPE;
  IVL_ARCH();
  cudaError_t e = ivl::launch_mrope(x, xs, cos, sin, B, T, Hn, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? IVL_OK : IVL_ERR_LAUNCH;
}

}  // extern "C"
