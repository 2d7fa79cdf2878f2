// Small kernels around the sliding-window attention mixer.
//
//   mrope_apply : rotate q and k in place with per-token cos/sin rows (already merged over the
//                 three M-RoPE position rows).  Replaces apply_multimodal_rotary_pos_emb +
//                 rotate_half (infinitevl_standard/modeling_infinitevl.py:521-525,949-984).  The
//                 reference runs this in bf16 with every intermediate rounded (std:930 casts
//                 cos/sin to bf16 first), so the kernel rounds at the same three points and is
//                 bit-exact against the torch expression  q*cos + rotate_half(q)*sin.
//   swa_decode  : attention of a few new tokens against the cached window (split over the
//                 key axis, combined with a log-sum-exp reduction).
#include "sm100.cuh"

namespace ivl {

namespace {

__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16(x)); }

// x [B, T, Hn, 128] with element strides (sb, st, sh); cos/sin [B, T, 128] contiguous.
__global__ void mrope_kernel(__nv_bfloat16* __restrict__ x, long long sb, long long st, long long sh,
                             const __nv_bfloat16* __restrict__ cosr, const __nv_bfloat16* __restrict__ sinr, int T,
                             int Hn) {
  // one thread = 8 channels of the first half + the matching 8 of the second half
  const int piece = threadIdx.x & 7;        // 8 pieces x 8 = 64 channels
  const int h = blockIdx.y * (blockDim.x >> 3) + (threadIdx.x >> 3);
  const int t = blockIdx.x, b = blockIdx.z;
  if (h >= Hn) return;
  __nv_bfloat16* p = x + b * sb + t * st + h * sh + piece * 8;
  const __nv_bfloat16* c = cosr + ((size_t)b * T + t) * 128 + piece * 8;
  const __nv_bfloat16* s = sinr + ((size_t)b * T + t) * 128 + piece * 8;
  uint4 lo = *reinterpret_cast<const uint4*>(p), hi = *reinterpret_cast<const uint4*>(p + 64);
  const uint4 clo = __ldg(reinterpret_cast<const uint4*>(c)), chi = __ldg(reinterpret_cast<const uint4*>(c + 64));
  const uint4 slo = __ldg(reinterpret_cast<const uint4*>(s)), shi = __ldg(reinterpret_cast<const uint4*>(s + 64));
  uint32_t* l = reinterpret_cast<uint32_t*>(&lo);
  uint32_t* u = reinterpret_cast<uint32_t*>(&hi);
  const uint32_t* cl = reinterpret_cast<const uint32_t*>(&clo);
  const uint32_t* cu = reinterpret_cast<const uint32_t*>(&chi);
  const uint32_t* sl = reinterpret_cast<const uint32_t*>(&slo);
  const uint32_t* su = reinterpret_cast<const uint32_t*>(&shi);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float a0 = bf16_lo(l[e]), a1 = bf16_hi(l[e]);   // x[i]
    const float b0 = bf16_lo(u[e]), b1 = bf16_hi(u[e]);   // x[i + 64]
    // out[i]      = bf16( bf16(x[i] cos[i])       + bf16(-x[i+64] sin[i]) )
    // out[i + 64] = bf16( bf16(x[i+64] cos[i+64]) + bf16( x[i]    sin[i+64]) )
    const float o0 = rbf(a0 * bf16_lo(cl[e])) + rbf(-b0 * bf16_lo(sl[e]));
    const float o1 = rbf(a1 * bf16_hi(cl[e])) + rbf(-b1 * bf16_hi(sl[e]));
    const float q0 = rbf(b0 * bf16_lo(cu[e])) + rbf(a0 * bf16_lo(su[e]));
    const float q1 = rbf(b1 * bf16_hi(cu[e])) + rbf(a1 * bf16_hi(su[e]));
    l[e] = pack_bf16(o0, o1);
    u[e] = pack_bf16(q0, q1);
  }
  *reinterpret_cast<uint4*>(p) = lo;
  *reinterpret_cast<uint4*>(p + 64) = hi;
}

}  // namespace

cudaError_t launch_mrope(void* x, const long long* xs, const void* cosr, const void* sinr, int B, int T, int Hn,
                         cudaStream_t stream) {
  const int heads_per_block = Hn < 16 ? Hn : 16;
  dim3 grid(T, (Hn + heads_per_block - 1) / heads_per_block, B);
  mrope_kernel<<<grid, heads_per_block * 8, 0, stream>>>(static_cast<__nv_bfloat16*>(x), xs[0], xs[1], xs[2],
                                                         static_cast<const __nv_bfloat16*>(cosr),
                                                         static_cast<const __nv_bfloat16*>(sinr), T, Hn);
  return cudaGetLastError();
}

}  // namespace ivl
