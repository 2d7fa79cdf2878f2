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
#include <atomic>
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
// K/V window once: 2 * Hkv * Tk * 128 * 2 bytes), so it is a split-KV kernel: CTA = (128-key slice,
// kv-head, batch) serving all `group` q-heads of that kv-head, followed by a log-sum-exp combine.
// The slice is staged with cp.async and both products run on warp-level tensor-core MMAs (the 8
// q-heads of a kv-head are the -- half empty -- M = 16 rows of an m16n8k16 tile): scores = Q K^T
// with the four warps splitting the keys, O = P V with the warps splitting the head dim.  (The first
// version did these with scalar FMAs over shared memory and took ~20 us per layer for 8.4 MB.)
// Replaces the flash-attn call for q_len == 1 (std:1092-1108) -- the O(W) cache roll stays in the
// cache class.
// ---------------------------------------------------------------------------------------------
constexpr int DEC_KEYS = 128;     // keys per CTA
constexpr int DEC_MAXG = 8;       // q-heads per kv-head supported
constexpr int DEC_LD = 136;       // padded smem row (bf16 elements): 272 B, conflict-free ldmatrix

struct DecSmem {
  __nv_bfloat16 k[DEC_KEYS * DEC_LD];
  __nv_bfloat16 v[DEC_KEYS * DEC_LD];
  __nv_bfloat16 q[16 * DEC_LD];     // rows >= group are zero
  __nv_bfloat16 p[16 * DEC_LD];     // probabilities, bf16 (as in the prefill kernel)
  float red_m[DEC_MAXG][4], red_l[DEC_MAXG][4];
};

__device__ __forceinline__ void dec_ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void dec_ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void dec_mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128)
swa_decode_partial_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, long long k_sb,
                          long long k_st, long long k_sh, const __nv_bfloat16* __restrict__ v, long long v_sb,
                          long long v_st, long long v_sh, float* __restrict__ part, int Tk, int Hq, int group,
                          int first_key, float scale_log2) {
  extern __shared__ __align__(16) uint8_t dsm[];
  DecSmem& s = *reinterpret_cast<DecSmem*>(dsm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;   // mma fragment coordinates: row (= q-head) and column pair
  const int split = blockIdx.x, hk = blockIdx.y, b = blockIdx.z, nsplit = gridDim.x;
  const int j0 = first_key + split * DEC_KEYS;
  const int nk = min(DEC_KEYS, Tk - j0);
  // stage the K and V slices (16-byte async copies, rows past the end zeroed) and the group's queries
  for (int i = tid; i < DEC_KEYS * 16; i += 128) {
    const int r = i >> 4, c = i & 15;
    __nv_bfloat16* kd = s.k + r * DEC_LD + c * 8;
    __nv_bfloat16* vd = s.v + r * DEC_LD + c * 8;
    if (r < nk) {
      const __nv_bfloat16* kg = k + b * k_sb + (long long)(j0 + r) * k_st + hk * k_sh + c * 8;
      const __nv_bfloat16* vg = v + b * v_sb + (long long)(j0 + r) * v_st + hk * v_sh + c * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(kd)), "l"(kg));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(vd)), "l"(vg));
    } else {
      *reinterpret_cast<uint4*>(kd) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(vd) = make_uint4(0, 0, 0, 0);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = tid; i < 16 * 16; i += 128) {
    const int r = i >> 4, c = i & 15;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (r < group) u = __ldg(reinterpret_cast<const uint4*>(q + ((long long)b * Hq + hk * group + r) * 128) + c);
    *reinterpret_cast<uint4*>(s.q + r * DEC_LD + c * 8) = u;
    *reinterpret_cast<uint4*>(s.p + r * DEC_LD + c * 8) = make_uint4(0, 0, 0, 0);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  // ---- scores: warp w owns keys 32w .. 32w+31 (four n-tiles of 8 keys) ----
  float sc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) sc[i][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    uint32_t a[4];
    dec_ldsm_x4(smem_u32(s.q + (lane & 15) * DEC_LD + ks * 16 + (lane >> 4) * 8), a[0], a[1], a[2], a[3]);
#pragma unroll
    for (int ntp = 0; ntp < 2; ++ntp) {
      uint32_t b0, b1, b2, b3;
      const int n = warp * 32 + ntp * 16 + (lane & 7) + (lane >> 4) * 8, kk = ks * 16 + ((lane >> 3) & 1) * 8;
      dec_ldsm_x4(smem_u32(s.k + n * DEC_LD + kk), b0, b1, b2, b3);
      dec_mma16816(sc[2 * ntp], a, b0, b1);
      dec_mma16816(sc[2 * ntp + 1], a, b2, b3);
    }
  }
  // this thread holds, for q-head gq (accumulator rows gq; rows gq + 8 are padding), keys 32w + 8nt + 2tq + {0,1}
  float m = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int key = warp * 32 + nt * 8 + 2 * tq + e;
      const float x = key < nk ? sc[nt][e] * scale_log2 : -INFINITY;
      sc[nt][e] = x;
      m = fmaxf(m, x);
    }
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  if (tq == 0) s.red_m[gq][warp] = m;
  __syncthreads();
  m = fmaxf(fmaxf(s.red_m[gq][0], s.red_m[gq][1]), fmaxf(s.red_m[gq][2], s.red_m[gq][3]));
  float l = 0.f;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float p0 = (m == -INFINITY) ? 0.f : exp2f(sc[nt][0] - m);
    const float p1 = (m == -INFINITY) ? 0.f : exp2f(sc[nt][1] - m);
    l += p0 + p1;
    *reinterpret_cast<uint32_t*>(s.p + gq * DEC_LD + warp * 32 + nt * 8 + 2 * tq) = pack_bf16(p0, p1);
  }
  l += __shfl_xor_sync(0xffffffffu, l, 1);
  l += __shfl_xor_sync(0xffffffffu, l, 2);
  if (tq == 0) s.red_l[gq][warp] = l;
  __syncthreads();

  // ---- O = P V: warp w owns head dims 32w .. 32w+31 (four n-tiles of 8 dims), all 128 keys ----
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    uint32_t a[4];
    dec_ldsm_x4(smem_u32(s.p + (lane & 15) * DEC_LD + ks * 16 + (lane >> 4) * 8), a[0], a[1], a[2], a[3]);
    const int kk = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ntp = 0; ntp < 2; ++ntp) {
      uint32_t b0, b1, b2, b3;
      dec_ldsm_x4_t(smem_u32(s.v + kk * DEC_LD + warp * 32 + ntp * 16 + (lane >> 4) * 8), b0, b1, b2, b3);
      dec_mma16816(acc[2 * ntp], a, b0, b1);
      dec_mma16816(acc[2 * ntp + 1], a, b2, b3);
    }
  }
  // partial record per (b, q-head, split): [m, l, o[128]]
  if (gq < group) {
    float* rec = part + (((long long)b * Hq + hk * group + gq) * nsplit + split) * 130;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
      *reinterpret_cast<float2*>(rec + 2 + warp * 32 + nt * 8 + 2 * tq) = make_float2(acc[nt][0], acc[nt][1]);
    if (warp == 0 && tq == 0) {
      rec[0] = m;
      rec[1] = s.red_l[gq][0] + s.red_l[gq][1] + s.red_l[gq][2] + s.red_l[gq][3];
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
  static std::atomic<bool> configured_dev[64];   // function attributes are per device
  int dev_ = 0;
  if (cudaError_t e = cudaGetDevice(&dev_)) return e;
  if (dev_ < 0 || dev_ >= 64) return cudaErrorInvalidDevice;
  std::atomic<bool>& configured = configured_dev[dev_];
  const int smem = (int)sizeof(DecSmem);
  if (!configured.load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(swa_decode_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured.store(true, std::memory_order_release);
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
