// HBM-bound element-wise pieces of the Gated DeltaNet mixer, one launch each:
//   short_conv_silu : depthwise causal conv (k = 4) + SiLU with a carried input tail
//                     (replaces causal_conv1d_fn / causal_conv1d_update, called from
//                      src/llamafactory/model/fla/modules/convolution.py:253-282; carried-state semantics of
//                      the pip package: site-packages/fla/modules/conv/triton/kernels.py:85-126)
//   gdn_gate        : g = -exp(A_log) softplus(a + dt_bias) (fp32), beta = sigmoid(b) (bf16)
//                     (infinitevl_standard/modeling_infinitevl.py:1293-1294)
//   rmsnorm_gated   : y = x rsqrt(mean x^2 + eps) w * gate sigmoid(gate) over 256-wide head vectors
//                     (fla/modules/fused_norm_gate.py:26-95)
// All are coalesced 16-byte-vector streaming kernels; fp32 math, bf16 I/O.
#include <stdint.h>

#include "sm100.cuh"

namespace ivl {

namespace {

constexpr int CONV_W = 4;
constexpr int CONV_TT = 32;       // tokens per thread
constexpr int CONV_THREADS = 128; // x 8 channels = 1024 channels per CTA

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) { f[2 * e] = bf16_lo(w[e]); f[2 * e + 1] = bf16_hi(w[e]); }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
  u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  return u;
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

__global__ void __launch_bounds__(CONV_THREADS)
short_conv_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                  const __nv_bfloat16* __restrict__ cache_in, __nv_bfloat16* __restrict__ y,
                  __nv_bfloat16* __restrict__ cache_out, int T, int D, int act,
                  const uint8_t* __restrict__ left_ctx) {
  const int c0 = (blockIdx.x * CONV_THREADS + threadIdx.x) * 8;  // first of 8 channels
  if (c0 >= D) return;
  const int b = blockIdx.z;
  const int t0 = blockIdx.y * CONV_TT;
  const int t1 = min(T, t0 + CONV_TT);
  // weights [D][4]: 8 channels x 4 taps = 64 contiguous bytes
  float wt[8][CONV_W];
  {
    const uint4* wp = reinterpret_cast<const uint4*>(w + (size_t)c0 * CONV_W);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float f[8];
      unpack8(__ldg(wp + i), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) wt[(i * 8 + e) / CONV_W][(i * 8 + e) % CONV_W] = f[e];
    }
  }
  const __nv_bfloat16* xb = x + (size_t)b * T * D + c0;
  // cache [B][D][4], newest input in column 3; value of "token" t < 0 is cache[..., 4 + t]
  auto load_row = [&](int t, float* f) {
    if (t >= 0) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(xb + (size_t)t * D)), f);
    } else if (cache_in != nullptr && t >= -CONV_W) {
      const __nv_bfloat16* cp = cache_in + ((size_t)b * D + c0) * CONV_W + (CONV_W + t);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __bfloat162float(cp[e * CONV_W]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = 0.f;
    }
  };
  float h0[8], h1[8], h2[8], cur[8];
  load_row(t0 - 3, h0);
  load_row(t0 - 2, h1);
  load_row(t0 - 1, h2);
  for (int t = t0; t < t1; ++t) {
    load_row(t, cur);
    float out[8];
    // packed sequences (cu_seqlens; fla/modules/convolution.py:224-251): left_ctx[t] = min(3, tokens of t's own
    // sequence before t) -- inputs of an earlier sequence never enter the window
    const int m = left_ctx ? (int)left_ctx[(size_t)b * T + t] : 3;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float a = wt[e][0] * (m >= 3 ? h0[e] : 0.f);
      a = fmaf(wt[e][1], m >= 2 ? h1[e] : 0.f, a);
      a = fmaf(wt[e][2], m >= 1 ? h2[e] : 0.f, a);
      a = fmaf(wt[e][3], cur[e], a);
      out[e] = act ? silu(a) : a;
      h0[e] = h1[e]; h1[e] = h2[e]; h2[e] = cur[e];
    }
    *reinterpret_cast<uint4*>(y + ((size_t)b * T + t) * D + c0) = pack8(out);
  }
  if (cache_out != nullptr && t1 == T) {
    // the last 4 inputs of [cache_in | x]; this CTA holds x[T-3..T-1] in h0..h2 only if it loaded them,
    // so re-read (cheap, once per call)
    float tail[CONV_W][8];
#pragma unroll
    for (int j = 0; j < CONV_W; ++j) load_row(T - CONV_W + j, tail[j]);
    __nv_bfloat16* cp = cache_out + ((size_t)b * D + c0) * CONV_W;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      uint2 v;
      v.x = pack_bf16(tail[0][e], tail[1][e]);
      v.y = pack_bf16(tail[2][e], tail[3][e]);
      *reinterpret_cast<uint2*>(cp + e * CONV_W) = v;
    }
  }
}

__global__ void gdn_gate_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ bproj,
                                const float* __restrict__ A_log, const float* __restrict__ dt_bias,
                                float* __restrict__ g, __nv_bfloat16* __restrict__ beta, long long n, int H) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int h = (int)(i % H);
  const float z = __bfloat162float(a[i]) + dt_bias[h];
  const float sp = z > 20.f ? z : log1pf(expf(z));  // torch softplus (threshold 20)
  g[i] = -expf(A_log[h]) * sp;
  const float bb = __bfloat162float(bproj[i]);
  beta[i] = __float2bfloat16(1.0f / (1.0f + expf(-bb)));
}

// one warp per 256-wide row
__global__ void __launch_bounds__(256)
rmsnorm_gated_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gate,
                     const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ y, long long rows, float eps) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float xv[8], gv[8], wv[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(x + row * 256) + lane), xv);
  unpack8(__ldg(reinterpret_cast<const uint4*>(gate + row * 256) + lane), gv);
  unpack8(__ldg(reinterpret_cast<const uint4*>(w) + lane), wv);
  float ss = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) ss = fmaf(xv[e], xv[e], ss);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
  const float rstd = 1.0f / sqrtf(ss * (1.0f / 256.0f) + eps);
  float out[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) out[e] = xv[e] * rstd * wv[e] * gv[e] / (1.0f + __expf(-gv[e]));
  *(reinterpret_cast<uint4*>(y + row * 256) + lane) = pack8(out);
}

}  // namespace

cudaError_t launch_short_conv(const void* x, const void* w, const void* cache_in, void* y, void* cache_out, int B,
                              int T, int D, int act, cudaStream_t stream, const uint8_t* left_ctx) {
  dim3 grid((D / 8 + CONV_THREADS - 1) / CONV_THREADS, (T + CONV_TT - 1) / CONV_TT, B);
  short_conv_kernel<<<grid, CONV_THREADS, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w),
      static_cast<const __nv_bfloat16*>(cache_in), static_cast<__nv_bfloat16*>(y),
      static_cast<__nv_bfloat16*>(cache_out), T, D, act, left_ctx);
  return cudaGetLastError();
}

cudaError_t launch_gdn_gate(const void* a, const void* b, const float* A_log, const float* dt_bias, float* g,
                            void* beta, long long n, int H, cudaStream_t stream) {
  const int threads = 256;
  gdn_gate_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), A_log, dt_bias, g,
      static_cast<__nv_bfloat16*>(beta), n, H);
  return cudaGetLastError();
}

cudaError_t launch_rmsnorm_gated(const void* x, const void* gate, const void* w, void* y, long long rows, float eps,
                                 cudaStream_t stream) {
  rmsnorm_gated_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(gate),
      static_cast<const __nv_bfloat16*>(w), static_cast<__nv_bfloat16*>(y), rows, eps);
  return cudaGetLastError();
}

}  // namespace ivl
