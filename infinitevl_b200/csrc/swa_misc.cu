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

// ---------------------------------------------------------------------------------------------
// Decode attention: one new token per sequence against the cached window.  HBM-bound (reads the
// K/V window once: 2 * Hkv * Tk * 128 * 2 bytes), so it is a split-KV CUDA-core kernel: CTA =
// (128-key slice, kv-head, batch) serving all `group` q-heads of that kv-head, followed by a
// log-sum-exp combine.  Replaces the flash-attn call for q_len == 1 (std:1092-1108) -- and the O(W)
// cache roll stays in the cache class.
// ---------------------------------------------------------------------------------------------
constexpr int DEC_KEYS = 128;     // keys per CTA
constexpr int DEC_MAXG = 8;       // q-heads per kv-head supported
constexpr int DEC_LD = 136;       // padded smem row (bf16 elements)

__global__ void __launch_bounds__(128)
swa_decode_partial_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, long long k_sb,
                          long long k_st, long long k_sh, const __nv_bfloat16* __restrict__ v, long long v_sb,
                          long long v_st, long long v_sh, float* __restrict__ part, int Tk, int Hq, int group,
                          int first_key, float scale_log2) {
  extern __shared__ __align__(16) uint8_t dsm[];
  __nv_bfloat16* sk = reinterpret_cast<__nv_bfloat16*>(dsm);              // [128][136]
  __nv_bfloat16* sv = sk + DEC_KEYS * DEC_LD;                             // [128][136]
  float* sq = reinterpret_cast<float*>(sv + DEC_KEYS * DEC_LD);           // [8][128]
  float* sp = sq + DEC_MAXG * 128;                                        // [8][128]
  __shared__ float red[DEC_MAXG][4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, hk = blockIdx.y, b = blockIdx.z, nsplit = gridDim.x;
  const int j0 = first_key + split * DEC_KEYS;
  const int nk = min(DEC_KEYS, Tk - j0);
  // stage K and V slices (coalesced 16-byte copies) and the group's queries
  for (int i = tid; i < DEC_KEYS * 16; i += 128) {
    const int r = i >> 4, c = i & 15;
    uint4 kk = make_uint4(0, 0, 0, 0), vv = kk;
    if (r < nk) {
      kk = __ldg(reinterpret_cast<const uint4*>(k + b * k_sb + (long long)(j0 + r) * k_st + hk * k_sh) + c);
      vv = __ldg(reinterpret_cast<const uint4*>(v + b * v_sb + (long long)(j0 + r) * v_st + hk * v_sh) + c);
    }
    *reinterpret_cast<uint4*>(sk + r * DEC_LD + c * 8) = kk;
    *reinterpret_cast<uint4*>(sv + r * DEC_LD + c * 8) = vv;
  }
  for (int i = tid; i < group * 128; i += 128)
    sq[i] = __bfloat162float(q[((long long)b * Hq + hk * group) * 128 + i]) * scale_log2;
  __syncthreads();
  // scores: thread = key
  float sc[DEC_MAXG];
#pragma unroll
  for (int g = 0; g < DEC_MAXG; ++g) sc[g] = 0.f;
  {
    const __nv_bfloat16* kr = sk + tid * DEC_LD;
#pragma unroll 4
    for (int c = 0; c < 16; ++c) {
      const uint4 u = *reinterpret_cast<const uint4*>(kr + c * 8);
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&u);
      float kf[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { kf[2 * e] = bf16_lo(w[e]); kf[2 * e + 1] = bf16_hi(w[e]); }
#pragma unroll
      for (int g = 0; g < DEC_MAXG; ++g) {
        if (g < group) {
#pragma unroll
          for (int e = 0; e < 8; ++e) sc[g] = fmaf(kf[e], sq[g * 128 + c * 8 + e], sc[g]);
        }
      }
    }
  }
  const bool valid = tid < nk;
  float mloc[DEC_MAXG], lloc[DEC_MAXG];
#pragma unroll
  for (int g = 0; g < DEC_MAXG; ++g) {
    float x = valid ? sc[g] : -INFINITY;
    float m = x;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (lane == 0) red[g][warp] = m;
    mloc[g] = x;
  }
  __syncthreads();
#pragma unroll
  for (int g = 0; g < DEC_MAXG; ++g) {
    const float m = fmaxf(fmaxf(red[g][0], red[g][1]), fmaxf(red[g][2], red[g][3]));
    const float p = (m == -INFINITY) ? 0.f : exp2f(mloc[g] - m);
    sp[g * 128 + tid] = p;
    float l = p;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) l += __shfl_xor_sync(0xffffffffu, l, d);
    mloc[g] = m;
    lloc[g] = l;
  }
  __syncthreads();  // sp complete; red[] about to be reused for the sums
#pragma unroll
  for (int g = 0; g < DEC_MAXG; ++g)
    if (lane == 0) red[g][warp] = lloc[g];
  __syncthreads();
  // PV: thread = output dim
  float acc[DEC_MAXG];
#pragma unroll
  for (int g = 0; g < DEC_MAXG; ++g) acc[g] = 0.f;
  for (int r = 0; r < nk; ++r) {
    const float vv = __bfloat162float(sv[r * DEC_LD + tid]);
#pragma unroll
    for (int g = 0; g < DEC_MAXG; ++g)
      if (g < group) acc[g] = fmaf(sp[g * 128 + r], vv, acc[g]);
  }
  // partial record per (b, q-head, split): [m, l, o[128]]
#pragma unroll
  for (int g = 0; g < DEC_MAXG; ++g) {
    if (g < group) {
      float* rec = part + (((long long)b * Hq + hk * group + g) * nsplit + split) * 130;
      rec[2 + tid] = acc[g];
      if (tid == 0) {
        rec[0] = mloc[g];
        rec[1] = red[g][0] + red[g][1] + red[g][2] + red[g][3];
      }
    }
  }
}

__global__ void __launch_bounds__(128)
swa_decode_combine_kernel(const float* __restrict__ part, __nv_bfloat16* __restrict__ o, int nsplit) {
  const long long bh = blockIdx.x;  // b * Hq + h
  const int tid = threadIdx.x;
  const float* rec = part + bh * nsplit * 130;
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, rec[s * 130]);
  float L = 0.f, acc = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float m = rec[s * 130];
    const float w = (m == -INFINITY) ? 0.f : exp2f(m - M);
    L = fmaf(w, rec[s * 130 + 1], L);
    acc = fmaf(w, rec[s * 130 + 2 + tid], acc);
  }
  o[bh * 128 + tid] = __float2bfloat16(L > 0.f ? acc / L : 0.f);
}

}  // namespace

size_t swa_decode_workspace_bytes(int B, int Tk, int Hq) {
  const int nsplit = (Tk + DEC_KEYS - 1) / DEC_KEYS;
  return (size_t)B * Hq * nsplit * 130 * sizeof(float);
}

// q, o [B,1,Hq,128] contiguous; k, v [B,Tk,Hkv,128] with element strides (batch, time, head)
cudaError_t launch_swa_decode(const void* q, const void* k, const long long* ks, const void* v, const long long* vs,
                              void* o, int B, int Tk, int Hq, int Hkv, int window, float scale, void* workspace,
                              cudaStream_t stream) {
  static bool configured = false;
  const int smem = 2 * DEC_KEYS * DEC_LD * 2 + 2 * DEC_MAXG * 128 * 4;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(swa_decode_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int first_key = (window > 0 && Tk > window) ? Tk - window : 0;
  const int nsplit = (Tk - first_key + DEC_KEYS - 1) / DEC_KEYS;
  float* part = static_cast<float*>(workspace);
  dim3 grid(nsplit, Hkv, B);
  swa_decode_partial_kernel<<<grid, 128, smem, stream>>>(
      static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k), ks[0], ks[1], ks[2],
      static_cast<const __nv_bfloat16*>(v), vs[0], vs[1], vs[2], part, Tk, Hq, Hq / Hkv, first_key,
      scale * 1.4426950408889634f);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  swa_decode_combine_kernel<<<B * Hq, 128, 0, stream>>>(part, static_cast<__nv_bfloat16*>(o), nsplit);
  return cudaGetLastError();
}

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
