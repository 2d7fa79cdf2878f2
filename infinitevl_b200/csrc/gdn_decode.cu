// One decode step of a whole Gated DeltaNet mixer core in ONE launch: the three short convolutions
// (k = 4, SiLU, carried tails), the gate math, the token recurrence on the cached state and the
// gated RMS norm of the output -- everything between the input projections and o_proj of
// GatedDeltaNet.forward for q_len == 1 (infinitevl_standard/modeling_infinitevl.py:1263-1342).
//
// The unfused path costs seven launches per layer and step (3 x short_conv_kernel, gdn_gate_kernel,
// gdn_recurrent_kernel, rmsnorm_gated_kernel) plus up to four cache copies; at ~4 us per launch that
// is most of a decode step's mixer time.  All caches (conv tails, recurrent state) are updated in place,
// so the call is CUDA-graph friendly.
//
// One CTA per (head, batch row), 1024 threads: thread (kq, col) keeps the 32 state entries
// S[32 kq .. 32 kq + 31][col] of its head in registers (as gdn_recurrent_kernel does for a 32-column
// slice); the head's 128 + 128 + 256 conv channels are computed once by the first 512 threads.  Every
// arithmetic step is done in the same order and with the same roundings as the unfused kernels, so the
// results are bit-identical to them (tests/test_modules_gpu.py).
#include "gdn_layout.cuh"
#include "sm100.cuh"

namespace ivl {

namespace {

constexpr int DS_THREADS = 1024;

template <typename T>
__device__ __forceinline__ float ds_ld(const T* p);
template <>
__device__ __forceinline__ float ds_ld<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ds_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void ds_st(float* p, float v) { *p = v; }
__device__ __forceinline__ void ds_st(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }
__device__ __forceinline__ float ds_silu(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float ds_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

// depthwise conv step for one channel: cache holds the last four inputs (newest at [3]); returns SiLU(conv)
// rounded to bf16 (the unfused kernel writes bf16) and rolls the cache in place
__device__ __forceinline__ float conv_step(const __nv_bfloat16* x, const __nv_bfloat16* w, __nv_bfloat16* cache,
                                           size_t ch) {
  const uint2 cw = *reinterpret_cast<const uint2*>(cache + ch * 4);
  const uint2 ww = __ldg(reinterpret_cast<const uint2*>(w + ch * 4));
  const __nv_bfloat16 xb = x[ch];
  const float h0 = bf16_hi(cw.x), h1 = bf16_lo(cw.y), h2 = bf16_hi(cw.y), cur = __bfloat162float(xb);
  float a = bf16_lo(ww.x) * h0;
  a = fmaf(bf16_hi(ww.x), h1, a);
  a = fmaf(bf16_lo(ww.y), h2, a);
  a = fmaf(bf16_hi(ww.y), cur, a);
  uint2 nw;
  nw.x = (cw.x >> 16) | (cw.y << 16);                                  // old [1], [2]
  nw.y = (cw.y >> 16) | ((uint32_t)__bfloat16_as_ushort(xb) << 16);    // old [3], new input
  *reinterpret_cast<uint2*>(cache + ch * 4) = nw;
  return ds_round(ds_silu(a));
}

template <typename TS>
__global__ void __launch_bounds__(DS_THREADS, 1)
gdn_decode_step_kernel(const __nv_bfloat16* __restrict__ q_in, const __nv_bfloat16* __restrict__ k_in,
                       const __nv_bfloat16* __restrict__ v_in, const __nv_bfloat16* __restrict__ a_in,
                       const __nv_bfloat16* __restrict__ b_in, const __nv_bfloat16* __restrict__ gate_in,
                       const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk,
                       const __nv_bfloat16* __restrict__ wv, const float* __restrict__ A_log,
                       const float* __restrict__ dt_bias, const __nv_bfloat16* __restrict__ norm_w,
                       __nv_bfloat16* conv_q, __nv_bfloat16* conv_k, __nv_bfloat16* conv_v, TS* state,
                       __nv_bfloat16* __restrict__ out, int H, float scale, float eps) {
  __shared__ float qn[GDN_K], kn[GDN_K], vs[GDN_V];
  __shared__ float red[4][GDN_V];
  __shared__ float ssq[2][4];
  __shared__ float os[GDN_V];
  const int tid = threadIdx.x, kq = tid >> 8, col = tid & 255, lane = tid & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  // the state slice first: its latency hides behind the convolutions
  TS* sp = state + (((size_t)b * H + h) * GDN_K + kq * 32) * GDN_V + col;
  float S[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) S[j] = ds_ld<TS>(sp + (size_t)j * GDN_V);

  // ---- short convolutions of the head's channels (q: threads 0..127, k: 128..255, v: 256..511) ----
  float qv = 0.f, kv = 0.f;
  if (tid < 128) {
    const size_t ch = (size_t)h * GDN_K + tid;
    qv = conv_step(q_in + (size_t)b * H * GDN_K, wq, conv_q + (size_t)b * H * GDN_K * 4, ch);
  } else if (tid < 256) {
    const size_t ch = (size_t)h * GDN_K + (tid - 128);
    kv = conv_step(k_in + (size_t)b * H * GDN_K, wk, conv_k + (size_t)b * H * GDN_K * 4, ch);
  } else if (tid < 512) {
    const size_t ch = (size_t)h * GDN_V + (tid - 256);
    vs[tid - 256] = conv_step(v_in + (size_t)b * H * GDN_V, wv, conv_v + (size_t)b * H * GDN_V * 4, ch);
  }
  // ---- L2 norm of q and k (fp32, same reduction tree as gdn_recurrent_kernel) ----
  if (tid < 256) {
    float sq = tid < 128 ? qv * qv : kv * kv;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
    if (lane == 0) ssq[tid >> 7][(tid >> 5) & 3] = sq;
  }
  __syncthreads();
  if (tid < 128) {
    const float sq = ssq[0][0] + ssq[0][1] + ssq[0][2] + ssq[0][3];
    qn[tid] = qv * (1.0f / sqrtf(sq + 1e-6f)) * scale;
  } else if (tid < 256) {
    const float sk = ssq[1][0] + ssq[1][1] + ssq[1][2] + ssq[1][3];
    kn[tid - 128] = kv * (1.0f / sqrtf(sk + 1e-6f));
  }
  // ---- gates (gdn_gate_kernel): g = -exp(A_log) softplus(a + dt_bias), beta = bf16(sigmoid(b)) ----
  const float z = __bfloat162float(a_in[(size_t)b * H + h]) + dt_bias[h];
  const float spl = z > 20.f ? z : log1pf(expf(z));
  const float alpha = __expf(-expf(A_log[h]) * spl);
  const float bt = ds_round(1.0f / (1.0f + expf(-__bfloat162float(b_in[(size_t)b * H + h]))));
  __syncthreads();
  // ---- the recurrence (gdn_recurrent_kernel, one token) ----
  float pred = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    S[j] *= alpha;
    pred = fmaf(S[j], kn[kq * 32 + j], pred);
  }
  red[kq][col] = pred;
  __syncthreads();
  pred = red[0][col] + red[1][col] + red[2][col] + red[3][col];
  const float vp = (vs[col] - pred) * bt;
  float o = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    S[j] = fmaf(kn[kq * 32 + j], vp, S[j]);
    o = fmaf(S[j], qn[kq * 32 + j], o);
  }
  __syncthreads();  // everyone has consumed red[] (pred) before it is reused
  red[kq][col] = o;
#pragma unroll
  for (int j = 0; j < 32; ++j) ds_st(sp + (size_t)j * GDN_V, S[j]);
  __syncthreads();
  if (kq == 0) os[col] = ds_round(red[0][col] + red[1][col] + red[2][col] + red[3][col]);
  __syncthreads();
  // ---- gated RMS norm of the head's 256 outputs (rmsnorm_gated_kernel: one warp, 8 elements per lane) ----
  if (tid < 32) {
    const size_t row = (size_t)b * H + h;
    const uint4 gu = __ldg(reinterpret_cast<const uint4*>(gate_in + row * GDN_V) + lane);
    const uint4 wu = __ldg(reinterpret_cast<const uint4*>(norm_w) + lane);
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(&gu);
    const uint32_t* ww = reinterpret_cast<const uint32_t*>(&wu);
    float xv[8], gv[8], wv8[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      gv[2 * e] = bf16_lo(gw[e]); gv[2 * e + 1] = bf16_hi(gw[e]);
      wv8[2 * e] = bf16_lo(ww[e]); wv8[2 * e + 1] = bf16_hi(ww[e]);
    }
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      xv[e] = os[lane * 8 + e];
      ss = fmaf(xv[e], xv[e], ss);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    const float rstd = 1.0f / sqrtf(ss * (1.0f / 256.0f) + eps);
    uint4 y;
    uint32_t* yw = reinterpret_cast<uint32_t*>(&y);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float y0 = xv[2 * e] * rstd * wv8[2 * e] * gv[2 * e] / (1.0f + __expf(-gv[2 * e]));
      const float y1 = xv[2 * e + 1] * rstd * wv8[2 * e + 1] * gv[2 * e + 1] / (1.0f + __expf(-gv[2 * e + 1]));
      yw[e] = pack_bf16(y0, y1);
    }
    *(reinterpret_cast<uint4*>(out + row * GDN_V) + lane) = y;
  }
}

}  // namespace

cudaError_t launch_gdn_decode_step(const void* q_in, const void* k_in, const void* v_in, const void* a_in,
                                   const void* b_in, const void* gate_in, const void* wq, const void* wk,
                                   const void* wv, const float* A_log, const float* dt_bias, const void* norm_w,
                                   void* conv_q, void* conv_k, void* conv_v, void* state, int state_dtype, void* out,
                                   int B, int H, float scale, float eps, cudaStream_t stream) {
  dim3 grid(H, B);
  auto bf = [](const void* p) { return static_cast<const __nv_bfloat16*>(p); };
  auto bfm = [](void* p) { return static_cast<__nv_bfloat16*>(p); };
  if (state_dtype == 0)
    gdn_decode_step_kernel<float><<<grid, DS_THREADS, 0, stream>>>(
        bf(q_in), bf(k_in), bf(v_in), bf(a_in), bf(b_in), bf(gate_in), bf(wq), bf(wk), bf(wv), A_log, dt_bias, bf(norm_w),
        bfm(conv_q), bfm(conv_k), bfm(conv_v), static_cast<float*>(state), bfm(out), H, scale, eps);
  else
    gdn_decode_step_kernel<__nv_bfloat16><<<grid, DS_THREADS, 0, stream>>>(
        bf(q_in), bf(k_in), bf(v_in), bf(a_in), bf(b_in), bf(gate_in), bf(wq), bf(wk), bf(wv), A_log, dt_bias, bf(norm_w),
        bfm(conv_q), bfm(conv_k), bfm(conv_v), static_cast<__nv_bfloat16*>(state), bfm(out), H, scale, eps);
  return cudaGetLastError();
}

}  // namespace ivl
