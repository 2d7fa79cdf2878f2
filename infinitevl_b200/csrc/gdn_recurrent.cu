// Gated DeltaNet token recurrence (decode and short prefill, q_len <= 64 in the model).
//
// Replaces fused_recurrent_gated_delta_rule_fwd_kernel
// (src/llamafactory/model/fla/ops/gated_delta_rule/fused_recurrent.py:21-112): all math in
// fp32, q/k normalised in-kernel without an intermediate bf16 rounding, exactly as there.
//
//   S <- exp(g_t) S ;  v' = beta_t (v_t - S^T k_t) ;  S <- S + k_t v'^T ;  o_t = scale S^T q_t
//
// One CTA per (batch, head, 32 value columns); thread (kq, col) keeps the 32 state
// entries S[32 kq .. 32 kq + 31][col] in registers for the whole call, so a decode step
// touches the state exactly once in each direction (HBM-bound: 2 x 128 x 256 x elt bytes
// per head).  The state may be updated in place (h0 == ht) for CUDA-graph replay.
#include "gdn_layout.cuh"
#include "sm100.cuh"

namespace ivl {

namespace {

constexpr int REC_THREADS = 128;

template <typename T>
__device__ __forceinline__ float ld_state(const T* p);
template <>
__device__ __forceinline__ float ld_state<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_state<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_state(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_state(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(REC_THREADS)
gdn_recurrent_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                     const __nv_bfloat16* __restrict__ v, const float* __restrict__ g,
                     const __nv_bfloat16* __restrict__ beta, const TIn* h0, __nv_bfloat16* __restrict__ o, TOut* ht,
                     int T, int H, float scale, int l2norm) {
  __shared__ float qn[GDN_K], kn[GDN_K];
  __shared__ float red[4][GDN_BV];
  __shared__ float ssq[2][4];
  const int tid = threadIdx.x, kq = tid >> 5, col = tid & 31;
  const int slice = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const size_t sbase = (((size_t)b * H + h) * GDN_K + kq * 32) * GDN_V + slice * GDN_BV + col;
  float S[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) S[j] = h0 ? ld_state<TIn>(h0 + sbase + (size_t)j * GDN_V) : 0.f;

  for (int t = 0; t < T; ++t) {
    const size_t tok = (size_t)b * T + t;
    // q, k rows: one element per thread, block-wide sum of squares
    float qv = __bfloat162float(q[(tok * H + h) * GDN_K + tid]);
    float kv = __bfloat162float(k[(tok * H + h) * GDN_K + tid]);
    if (l2norm) {
      float sq = qv * qv, sk = kv * kv;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, d);
        sk += __shfl_xor_sync(0xffffffffu, sk, d);
      }
      if (col == 0) { ssq[0][kq] = sq; ssq[1][kq] = sk; }
      __syncthreads();
      sq = ssq[0][0] + ssq[0][1] + ssq[0][2] + ssq[0][3];
      sk = ssq[1][0] + ssq[1][1] + ssq[1][2] + ssq[1][3];
      qv *= 1.0f / sqrtf(sq + 1e-6f);
      kv *= 1.0f / sqrtf(sk + 1e-6f);
    }
    qn[tid] = qv * scale;
    kn[tid] = kv;
    const float alpha = __expf(g[tok * H + h]);
    const float bt = __bfloat162float(beta[tok * H + h]);
    const float vt = __bfloat162float(v[(tok * H + h) * GDN_V + slice * GDN_BV + col]);
    __syncthreads();
    // prediction of the decayed memory for k_t
    float pred = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      S[j] *= alpha;
      pred = fmaf(S[j], kn[kq * 32 + j], pred);
    }
    red[kq][col] = pred;
    __syncthreads();
    pred = red[0][col] + red[1][col] + red[2][col] + red[3][col];
    const float vp = (vt - pred) * bt;
    float out = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      S[j] = fmaf(kn[kq * 32 + j], vp, S[j]);
      out = fmaf(S[j], qn[kq * 32 + j], out);
    }
    __syncthreads();  // everyone has consumed red[] (pred) before it is reused
    red[kq][col] = out;
    __syncthreads();
    if (kq == 0)
      o[(tok * H + h) * GDN_V + slice * GDN_BV + col] =
          __float2bfloat16(red[0][col] + red[1][col] + red[2][col] + red[3][col]);
    // next iteration's first __syncthreads (or the ssq one) orders the reuse of qn/kn/red
    __syncthreads();
  }
  if (ht) {
#pragma unroll
    for (int j = 0; j < 32; ++j) st_state(ht + sbase + (size_t)j * GDN_V, S[j]);
  }
}

}  // namespace

cudaError_t launch_gdn_recurrent(const void* q, const void* k, const void* v, const float* g, const void* beta,
                                 const void* h0, int h0_dtype, void* o, void* ht, int ht_dtype, int B, int T, int H,
                                 float scale, int l2norm, cudaStream_t stream) {
  dim3 grid(GDN_NS, H, B);
  auto Q = static_cast<const __nv_bfloat16*>(q);
  auto K = static_cast<const __nv_bfloat16*>(k);
  auto V = static_cast<const __nv_bfloat16*>(v);
  auto Bt = static_cast<const __nv_bfloat16*>(beta);
  auto O = static_cast<__nv_bfloat16*>(o);
#define IVL_REC(TI, TO)                                                                                   \
  gdn_recurrent_kernel<TI, TO><<<grid, REC_THREADS, 0, stream>>>(Q, K, V, g, Bt, static_cast<const TI*>(h0), O, \
                                                                 static_cast<TO*>(ht), T, H, scale, l2norm)
  if (h0_dtype == 0 && ht_dtype == 0) IVL_REC(float, float);
  else if (h0_dtype == 0) IVL_REC(float, __nv_bfloat16);
  else if (ht_dtype == 0) IVL_REC(__nv_bfloat16, float);
  else IVL_REC(__nv_bfloat16, __nv_bfloat16);
#undef IVL_REC
  return cudaGetLastError();
}

}  // namespace ivl
