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
constexpr int DEC_REC = 132;      // floats per partial record: [m, l, -, -, o[128]] (o 16-byte aligned)
constexpr int DEC_O = 4;

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

// One 128-key slice of the split-KV decode attention for the `group` q-heads of kv-head hk: stages the slice
// (rows(r, &kg, &vg) yields the global K / V row of slice row r), computes scores, probabilities and the partial
// output and writes the record [m, l, o[128]] of (b, q-head, split).  nk = valid rows of the slice (<= 0: none).
template <class Rows>
__device__ __forceinline__ void dec_partial_body(DecSmem& s, const __nv_bfloat16* __restrict__ q, Rows rows,
                                                 float* __restrict__ part, int nk, int Hq, int group, int split,
                                                 int nsplit, int hk, int b, float scale_log2) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;   // mma fragment coordinates: row (= q-head) and column pair
  // stage the K and V slices (16-byte async copies, rows past the end zeroed) and the group's queries
  for (int i = tid; i < DEC_KEYS * 16; i += 128) {
    const int r = i >> 4, c = i & 15;
    __nv_bfloat16* kd = s.k + r * DEC_LD + c * 8;
    __nv_bfloat16* vd = s.v + r * DEC_LD + c * 8;
    if (r < nk) {
      const __nv_bfloat16 *kg, *vg;
      rows(r, kg, vg);
      kg += c * 8;
      vg += c * 8;
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
  // partial record per (b, q-head, split): [m, l, -, -, o[128]]
  if (gq < group) {
    float* rec = part + (((long long)b * Hq + hk * group + gq) * nsplit + split) * DEC_REC;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
      *reinterpret_cast<float2*>(rec + DEC_O + warp * 32 + nt * 8 + 2 * tq) = make_float2(acc[nt][0], acc[nt][1]);
    if (warp == 0 && tq == 0) {
      rec[0] = m;
      rec[1] = s.red_l[gq][0] + s.red_l[gq][1] + s.red_l[gq][2] + s.red_l[gq][3];
    }
  }
}

__global__ void __launch_bounds__(128)
swa_decode_partial_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, long long k_sb,
                          long long k_st, long long k_sh, const __nv_bfloat16* __restrict__ v, long long v_sb,
                          long long v_st, long long v_sh, float* __restrict__ part, int Tk, int Hq, int group,
                          int first_key, float scale_log2) {
  extern __shared__ __align__(16) uint8_t dsm[];
  DecSmem& s = *reinterpret_cast<DecSmem*>(dsm);
  const int split = blockIdx.x, hk = blockIdx.y, b = blockIdx.z, nsplit = gridDim.x;
  const int j0 = first_key + split * DEC_KEYS;
  const int nk = min(DEC_KEYS, Tk - j0);
  dec_partial_body(s, q, [&](int r, const __nv_bfloat16*& kg, const __nv_bfloat16*& vg) {
    kg = k + b * k_sb + (long long)(j0 + r) * k_st + hk * k_sh;
    vg = v + b * v_sb + (long long)(j0 + r) * v_st + hk * v_sh;
  }, part, nk, Hq, group, split, nsplit, hk, b, scale_log2);
}

// log-sum-exp combine of the partial records of one (b, q-head); thread tid owns head dim tid
__device__ __forceinline__ void dec_combine(const float* rec, __nv_bfloat16* o_row, int nsplit, int tid) {
  float M = -INFINITY;
  for (int sp = 0; sp < nsplit; ++sp) M = fmaxf(M, __ldcg(rec + sp * DEC_REC));
  float L = 0.f, acc = 0.f;
  for (int sp = 0; sp < nsplit; ++sp) {
    const float m = __ldcg(rec + sp * DEC_REC);
    const float w = (m == -INFINITY) ? 0.f : exp2f(m - M);
    L = fmaf(w, __ldcg(rec + sp * DEC_REC + 1), L);
    acc = fmaf(w, __ldcg(rec + sp * DEC_REC + DEC_O + tid), acc);
  }
  o_row[tid] = __float2bfloat16(L > 0.f ? acc / L : 0.f);
}

// ---------------------------------------------------------------------------------------------
// Ring-buffer window cache (replaces the O(W) roll of
// StaticSlidingWindowLayerPrealloc.update, std:142-173).  Storage [B, 2R, Hkv, 128]: token t lives in slot t % R AND
// in slot t % R + R, so the last n <= R tokens are always one contiguous run starting at ((cum - n) % R).  The
// token counter lives in DEVICE memory (state[0]), so a captured CUDA graph replays correctly -- the reference
// keeps it in Python integers, which a replay cannot advance.
//   state (int32): [0] tokens appended so far, [1] blocks-done counter of the running launch,
//                  [2 + b * Hkv + hk] slice-done counters of the decode kernel.
// ---------------------------------------------------------------------------------------------

// Decode step in ONE launch: append the new token's K/V to the ring, attend over the last min(cum + 1, W) tokens
// (split-KV partials), combine (the last slice block of a (b, kv-head) to finish) and advance the counter (the
// last block of the launch).  knew / vnew: [B, 1, Hkv, 128] projection outputs with (batch, head) strides.
__global__ void __launch_bounds__(128)
swa_ring_decode_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ knew, long long kn_sb,
                       long long kn_sh, const __nv_bfloat16* __restrict__ vnew, long long vn_sb, long long vn_sh,
                       __nv_bfloat16* __restrict__ ring_k, __nv_bfloat16* __restrict__ ring_v, int* __restrict__ state,
                       float* __restrict__ part, __nv_bfloat16* __restrict__ o, int Hq, int Hkv, int group, int R, int W,
                       float scale_log2) {
  extern __shared__ __align__(16) uint8_t dsm[];
  DecSmem& s = *reinterpret_cast<DecSmem*>(dsm);
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int split = blockIdx.x, hk = blockIdx.y, b = blockIdx.z, nsplit = gridDim.x;
  const int cum = *reinterpret_cast<volatile int*>(state);      // tokens in the ring before this step
  const int n = min(cum + 1, W);                                 // window incl. the new token
  const int base_tok = cum + 1 - n;                              // oldest visible token
  const int phys0 = base_tok % R;
  const int j0 = split * DEC_KEYS;
  const int nk = min(DEC_KEYS, n - j0);
  const long long t_st = (long long)Hkv * 128, b_st = 2ll * R * t_st;
  __nv_bfloat16* rk = ring_k + b * b_st + hk * 128;
  __nv_bfloat16* rv = ring_v + b * b_st + hk * 128;
  const __nv_bfloat16* kn = knew + b * kn_sb + hk * kn_sh;
  const __nv_bfloat16* vn = vnew + b * vn_sb + hk * vn_sh;
  if (nk > 0 && j0 + nk == n && tid < 32) {
    // this slice ends with the new token: store it into both of its ring slots for the steps to come
    const int slot = cum % R, c = tid & 15;
    const uint4 val = __ldg(reinterpret_cast<const uint4*>((tid < 16 ? kn : vn) + c * 8));
    __nv_bfloat16* dst = (tid < 16 ? rk : rv) + c * 8;
    *reinterpret_cast<uint4*>(dst + slot * t_st) = val;
    *reinterpret_cast<uint4*>(dst + (slot + R) * t_st) = val;
  }
  dec_partial_body(s, q, [&](int r, const __nv_bfloat16*& kg, const __nv_bfloat16*& vg) {
    if (j0 + r == n - 1) { kg = kn; vg = vn; }                  // the new token comes from the projection output
    else { kg = rk + (long long)(phys0 + j0 + r) * t_st; vg = rv + (long long)(phys0 + j0 + r) * t_st; }
  }, part, nk, Hq, group, split, nsplit, hk, b, scale_log2);
  // ---- last slice block of this (b, kv-head): combine; last block of the launch: advance the counter ----
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(state + 2 + b * Hkv + hk, 1) == nsplit - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // Block-parallel log-sum-exp combine of the group's records (up to 8 heads x 64 slices x 528 B = 270 KB): the
  // slice maxima / sums first (one load per thread and record), then four heads at a time, one float4 of the output
  // per thread and slice with eight loads in flight.  (A loop of dependent loads per head, as in the stand-alone
  // combine kernel, costs ~70 us here: one block would walk 8 heads x 64 slices serially.)
  {
    const int warp = tid >> 5, lane = tid & 31;
    float* sm_m = reinterpret_cast<float*>(dsm);              // [8][64]  (the slice staging area is dead)
    float* sm_w = sm_m + DEC_MAXG * 64;                      // [8][64]
    const float* recs = part + ((long long)b * Hq + hk * group) * nsplit * DEC_REC;
    for (int i = tid; i < group * nsplit; i += 128) {
      sm_m[(i / nsplit) * 64 + i % nsplit] = __ldcg(recs + (long long)i * DEC_REC);
      sm_w[(i / nsplit) * 64 + i % nsplit] = __ldcg(recs + (long long)i * DEC_REC + 1);
    }
    __syncthreads();
    for (int pass = 0; pass < 2; ++pass) {
      const int gh = warp + 4 * pass;                        // warp w normalises heads w and w + 4
      if (gh < group) {
        const float m0 = lane < nsplit ? sm_m[gh * 64 + lane] : -INFINITY;
        const float m1 = lane + 32 < nsplit ? sm_m[gh * 64 + lane + 32] : -INFINITY;
        float M = fmaxf(m0, m1);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, d));
        const float w0 = m0 == -INFINITY ? 0.f : exp2f(m0 - M), w1 = m1 == -INFINITY ? 0.f : exp2f(m1 - M);
        float L = (lane < nsplit ? w0 * sm_w[gh * 64 + lane] : 0.f) + (lane + 32 < nsplit ? w1 * sm_w[gh * 64 + lane + 32] : 0.f);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) L += __shfl_xor_sync(0xffffffffu, L, d);
        if (lane < nsplit) sm_w[gh * 64 + lane] = w0;
        if (lane + 32 < nsplit) sm_w[gh * 64 + lane + 32] = w1;
        if (lane == 0) sm_m[gh * 64] = L > 0.f ? 1.f / L : 0.f;   // slot 0 of the maxima now holds 1 / L
      }
    }
    __syncthreads();
    for (int pass = 0; pass < 2; ++pass) {
      const int gh = warp + 4 * pass;                        // warp w sums heads w and w + 4: lane = 4 head dims
      if (gh < group) {
        const float* r = recs + (long long)gh * nsplit * DEC_REC + DEC_O + 4 * lane;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int sp = 0;
        for (; sp + 8 <= nsplit; sp += 8) {
          float4 x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = __ldcg(reinterpret_cast<const float4*>(r + (long long)(sp + j) * DEC_REC));
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float w = sm_w[gh * 64 + sp + j];
            acc.x = fmaf(w, x[j].x, acc.x); acc.y = fmaf(w, x[j].y, acc.y);
            acc.z = fmaf(w, x[j].z, acc.z); acc.w = fmaf(w, x[j].w, acc.w);
          }
        }
        for (; sp < nsplit; ++sp) {
          const float4 x = __ldcg(reinterpret_cast<const float4*>(r + (long long)sp * DEC_REC));
          const float w = sm_w[gh * 64 + sp];
          acc.x = fmaf(w, x.x, acc.x); acc.y = fmaf(w, x.y, acc.y); acc.z = fmaf(w, x.z, acc.z); acc.w = fmaf(w, x.w, acc.w);
        }
        const float inv = sm_m[gh * 64];
        uint2 out;
        out.x = pack_bf16(acc.x * inv, acc.y * inv);
        out.y = pack_bf16(acc.z * inv, acc.w * inv);
        *reinterpret_cast<uint2*>(o + ((long long)b * Hq + hk * group + gh) * 128 + 4 * lane) = out;
      }
    }
  }
  if (tid == 0) {
    state[2 + b * Hkv + hk] = 0;                                 // ready for the next launch / graph replay
    if (atomicAdd(state + 1, 1) == (int)(gridDim.y * gridDim.z) - 1) {
      state[1] = 0;
      __threadfence();
      *reinterpret_cast<volatile int*>(state) = cum + 1;
    }
  }
}

// Append Tq new tokens (k, v: [B, Tq, Hkv, 128] with (batch, time, head) strides) to the ring; the last block to
// finish advances the counter.  Tokens older than the last R of the append are skipped.
__global__ void __launch_bounds__(256)
swa_ring_append_kernel(const __nv_bfloat16* __restrict__ k, long long k_sb, long long k_st, long long k_sh,
                       const __nv_bfloat16* __restrict__ v, long long v_sb, long long v_st, long long v_sh,
                       __nv_bfloat16* __restrict__ ring_k, __nv_bfloat16* __restrict__ ring_v, int* __restrict__ state,
                       int B, int Tq, int Hkv, int R) {
  const int cum = *reinterpret_cast<volatile int*>(state);
  const long long total = (long long)B * Tq * Hkv * 16;
  const long long t_st = (long long)Hkv * 128, b_st = 2ll * R * t_st;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 15);
    const int hk = (int)((i >> 4) % Hkv);
    const int t = (int)((i >> 4) / Hkv % Tq);
    const int b = (int)((i >> 4) / Hkv / Tq);
    if (t < Tq - R) continue;
    const int slot = (int)(((long long)cum + t) % R);
    const uint4 kv = __ldg(reinterpret_cast<const uint4*>(k + b * k_sb + (long long)t * k_st + hk * k_sh + c * 8));
    const uint4 vv = __ldg(reinterpret_cast<const uint4*>(v + b * v_sb + (long long)t * v_st + hk * v_sh + c * 8));
    __nv_bfloat16* dk = ring_k + b * b_st + hk * 128 + c * 8;
    __nv_bfloat16* dv = ring_v + b * b_st + hk * 128 + c * 8;
    *reinterpret_cast<uint4*>(dk + slot * t_st) = kv;
    *reinterpret_cast<uint4*>(dk + (slot + R) * t_st) = kv;
    *reinterpret_cast<uint4*>(dv + slot * t_st) = vv;
    *reinterpret_cast<uint4*>(dv + (slot + R) * t_st) = vv;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(state + 1, 1) == (int)gridDim.x - 1) {
    state[1] = 0;
    __threadfence();
    *reinterpret_cast<volatile int*>(state) = cum + Tq;
  }
}

__global__ void __launch_bounds__(128)
swa_decode_combine_kernel(const float* __restrict__ part, __nv_bfloat16* __restrict__ o, int nsplit) {
  const long long bh = blockIdx.x;  // b * Hq + h
  dec_combine(part + bh * nsplit * DEC_REC, o + bh * 128, nsplit, threadIdx.x);
}

}  // namespace

size_t swa_decode_workspace_bytes(int B, int Tk, int Hq) {
  const int nsplit = (Tk + DEC_KEYS - 1) / DEC_KEYS;
  return (size_t)B * Hq * nsplit * DEC_REC * sizeof(float);
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

static cudaError_t configure_decode_kernels() {
  static std::atomic<bool> configured_dev[64];
  int dev_ = 0;
  if (cudaError_t e = cudaGetDevice(&dev_)) return e;
  if (dev_ < 0 || dev_ >= 64) return cudaErrorInvalidDevice;
  if (!configured_dev[dev_].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(swa_ring_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(DecSmem));
    if (e != cudaSuccess) return e;
    configured_dev[dev_].store(true, std::memory_order_release);
  }
  return cudaSuccess;
}

size_t swa_ring_decode_workspace_bytes(int B, int Hq, int window) {
  const int nsplit = (window + DEC_KEYS - 1) / DEC_KEYS;
  return (size_t)B * Hq * nsplit * DEC_REC * sizeof(float);
}

cudaError_t launch_swa_ring_decode(const void* q, const void* knew, long long kn_sb, long long kn_sh, const void* vnew,
                                   long long vn_sb, long long vn_sh, void* ring_k, void* ring_v, int* state,
                                   void* workspace, void* o, int B, int Hq, int Hkv, int R, int window, float scale,
                                   cudaStream_t stream) {
  if (cudaError_t e = configure_decode_kernels()) return e;
  const int nsplit = (window + DEC_KEYS - 1) / DEC_KEYS;
  dim3 grid(nsplit, Hkv, B);
  swa_ring_decode_kernel<<<grid, 128, (int)sizeof(DecSmem), stream>>>(
      static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(knew), kn_sb, kn_sh,
      static_cast<const __nv_bfloat16*>(vnew), vn_sb, vn_sh, static_cast<__nv_bfloat16*>(ring_k),
      static_cast<__nv_bfloat16*>(ring_v), state, static_cast<float*>(workspace), static_cast<__nv_bfloat16*>(o), Hq, Hkv,
      Hq / Hkv, R, window, scale * 1.4426950408889634f);
  return cudaGetLastError();
}

cudaError_t launch_swa_ring_append(const void* k, const long long* ks, const void* v, const long long* vs, void* ring_k,
                                   void* ring_v, int* state, int B, int Tq, int Hkv, int R, cudaStream_t stream) {
  const long long total = (long long)B * Tq * Hkv * 16;
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  swa_ring_append_kernel<<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(k), ks[0], ks[1], ks[2],
                                                     static_cast<const __nv_bfloat16*>(v), vs[0], vs[1], vs[2],
                                                     static_cast<__nv_bfloat16*>(ring_k),
                                                     static_cast<__nv_bfloat16*>(ring_v), state, B, Tq, Hkv, R);
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
