// Sliding-window causal GQA attention, prefill (any Tq >= 1): flash-attention forward on
// tcgen05 tensor cores with TMA-fed operands and TMEM accumulators.
//
// Replaces the flash-attn wheel reached through the HF attention interface
// (ALL_ATTENTION_FUNCTIONS["flash_attention_2"], infinitevl_standard/modeling_infinitevl.py:1092-1108;
//  window rule transformers/modeling_flash_attention_utils.py:627-632): key j is visible to
// query i iff 0 <= (i + Tk - Tq) - j <= window - 1 (window == 0: plain bottom-right causal).
//
// One CTA = 128 consecutive queries of one q-head; two CTAs per SM so that one CTA's softmax
// overlaps the other's MMAs.  Per 64-key tile:
//   MMA   S = Q K^T                  M128 N64 K128   (Q, K: K-major 128B-swizzled TMA tiles)
//   warps row softmax (one thread per query row, no shuffles), lazy rescale of O in TMEM,
//         P -> bf16 -> TMEM, over the first 32 columns of the S buffer it was computed from
//   MMA   O += P V                   M128 N128 K64   (P: TMEM A operand; V: MN-major 128B-swizzled TMA tile)
// S is double-buffered in TMEM so S(t+1) is issued before P(t) is ready; S(t+2) reuses the buffer of S(t) / P(t)
// and is issued after PV(t), which the tensor pipe executes in order.  P never touches shared memory: the
// kernel was at the shared-memory port's limit (operand reads + TMA fills + 16 KB of P stores per tile).  The softmax is instruction-issue
// bound (ncu: tensor and XU pipes ~50 % each, profiles/r01c_summary.md), so its inner loops use the packed
// FP32x2 instructions of sm_100 (scale and max subtraction in one FFMA2, row sum in FADD2).
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 softmax / epilogue.
#include <atomic>
#include <cuda.h>
#include <stdlib.h>

#include <initializer_list>

#include "sm100.cuh"

namespace ivl {

namespace {

constexpr int BM = 128, BN = 64, HD = 128;
constexpr int SWA_THREADS = 192;
constexpr uint32_t Q_BYTES = BM * HD * 2;        // 32 KiB: 2 panels [128][64]
constexpr uint32_t KT_BYTES_ = BN * HD * 2;      // 16 KiB: 2 panels [64][64]
constexpr uint32_t OFF_Q = 0;
constexpr uint32_t OFF_K = Q_BYTES;              // 2 stages
constexpr uint32_t OFF_V = OFF_K + 2 * KT_BYTES_;
constexpr uint32_t DATA_BYTES = OFF_V + 2 * KT_BYTES_;  // 96 KiB
constexpr uint32_t SWA_SMEM = DATA_BYTES + 1024;   // barriers live in the alignment slack (or the tail)
constexpr uint32_t TM_S = 0;      // 2 buffers x 64 columns
constexpr uint32_t TM_O = 128;    // 128 columns
constexpr uint32_t TM_COLS = 256;
// Measured (tools/exp_swa_poly.py, B200 under its 1000 W cap): alone and hot, POLY = 3 is 6 % faster than 0
// (8.24 vs 8.79 ms at 128K); inside the bench step the extra FMA work lowers the sustained clock (1597-1635 vs
// 1650-1672 MHz) and the step does not get faster (150.9 / 151.4 vs 150.3 / 150.2 ms), so the default stays 0.
constexpr int SWA_POLY_DEFAULT = 0;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units: O is only rescaled when the row max grows by > 2^8

struct Bars {
  // K and V tiles travel in separate two-slot rings: a K slot is free as soon as S = Q K^T of its tile has
  // retired (before that tile's softmax even starts), a V slot only after O += P V, so the K tile of tile t+2 is
  // already in flight while tile t is still in the softmax
  uint64_t fullK[2], emptyK[2], fullV[2], emptyV[2], q, s[2], sfree[2], p, pv;
  uint32_t tmem_base;
};

struct SwaArgs {
  __nv_bfloat16* o;
  long long o_sb, o_st, o_sh;  // element strides of the output [B, Tq, Hq, D]
  int Tq, Tk, Hq, group;       // group = Hq / Hkv
  int window;                  // 0 = none
  float scale_log2;            // softmax scale * log2(e)
  // ring-buffer window cache (swa_misc.cu): K / V maps cover the physical ring [B, 2R, Hkv, D]; the number of cached
  // tokens is read from DEVICE memory (after the append of this call's tokens), so a CUDA-graph replay follows the
  // stream: Tk = min(cum, W - 1 + Tq), the keys start at ring slot (cum - Tk) % R
  const int* ring_state;       // nullptr: plain [B, Tk, Hkv, D] keys
  int ring_R, ring_W;
  // Position of key 0 in its sequence, modulo the key tile (ring cache: derived from the device counter).  The key
  // tiles are anchored at ABSOLUTE multiples of 64, so a query sees the same tiles -- the same running maxima, the same
  // bf16 roundings of P, the same summation order -- whether its keys arrive as one long prefill, as cache + new
  // tokens, or as halo + local shard: chunked, streamed and sequence-sharded prefills are bit-identical to the
  // one-shot prefill.  The partial first tile starts before key 0 (TMA zero-fills, the mask hides it).
  int kalign;
  // packed variable-length batch (cu_seqlens of the training collator): B = 1, Tq == Tk, the
  // token axis holds several sequences back to back.  One query tile never straddles two sequences; per tile the
  // tables give its first token and the [lo, hi) token range of its sequence.  Null: dense batch.
  const int* vt_tok0;
  const int* vt_lo;
  const int* vt_hi;
};

// 2^x for a pair of values on the FMA pipe instead of the XU pipe (two MUFU.EX2): Cody-Waite range reduction
// with the round-to-nearest magic constant, a degree-3 minimax polynomial on [-0.5, 0.5] (relative error 7.7e-5,
// far below the bf16 rounding P gets anyway) and the integer part added to the exponent field.  x <= ~8 here
// (lazy rescaling); masked scores (-inf) are clamped and come out as 2^-125 instead of 0, which no sum notices.
__device__ __forceinline__ float2 exp2_fma_pair(float2 x) {
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.f, 12582912.f));        // 1.5 * 2^23: integer in the low bits
  const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);                  // x - round(x)
  float2 p = __ffma2_rn(make_float2(0.05508868f, 0.05508868f), f, make_float2(0.24260405f, 0.24260405f));
  p = __ffma2_rn(p, f, make_float2(0.69327623f, 0.69327623f));
  p = __ffma2_rn(p, f, make_float2(0.99992895f, 0.99992895f));
  p.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  p.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return p;
}

// POLY: of every 8 pairs of exponentials, POLY are computed on the FMA pipe (exp2_fma_pair), the rest by MUFU.EX2.
// With two CTAs per SM the two softmax warps that share a scheduler are XU-bound in their exponential phase
// (64 MUFU x 8 issue cycles per tile and warp); moving part of them to the half-idle FMA pipe balances the two.
template <int POLY>
__global__ void __launch_bounds__(SWA_THREADS, 2)
swa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, SwaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t slack = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + slack;
  Bars& bars = *reinterpret_cast<Bars*>(slack >= 128 ? smem_raw : smem + DATA_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x, mt = blockIdx.y, b = blockIdx.z;
  const int hk = h / a.group;
  const bool packed = a.vt_tok0 != nullptr;
  const int i0 = packed ? __ldg(a.vt_tok0 + mt) : mt * BM;
  int koff = 0, kal = a.kalign;
  if (a.ring_state != nullptr) {
    const int cum = *reinterpret_cast<const volatile int*>(a.ring_state);
    a.Tk = min(cum, a.ring_W - 1 + a.Tq);
    koff = (cum - a.Tk) % a.ring_R;
    kal = (cum - a.Tk) & (BN - 1);
    a.window = a.Tk > a.ring_W ? a.ring_W : 0;             // the window rule of the HF glue, evaluated on the device
  }
  // [jmin, jmax): the keys this tile's queries may see at all; org: index of an aligned tile boundary (position 0 of the
  // sequence modulo the key tile); i_end: end of the query rows that exist
  const int jmin = packed ? __ldg(a.vt_lo + mt) : 0;
  const int jmax = packed ? __ldg(a.vt_hi + mt) : a.Tk;
  const int i_end = packed ? jmax : a.Tq;
  const int org = packed ? jmin : -kal;
  if (packed) a.window = (a.window > 0 && jmax - jmin > a.window) ? a.window : 0;   // per sequence, like the dense rule
  const int shift = packed ? 0 : a.Tk - a.Tq;              // bottom-right alignment
  const int p_first = i0 + shift;
  const int p_last = min(i0 + BM - 1, i_end - 1) + shift;
  const int lo_key = a.window > 0 ? max(jmin, p_first - a.window + 1) : jmin;
  const int t_lo = (lo_key - org) / BN, t_hi = (min(p_last, jmax - 1) - org) / BN;   // tiles of the ALIGNED key axis
  const int n_tiles = t_hi - t_lo + 1;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.fullK[i], 1); mbar_init(&bars.emptyK[i], 1);
      mbar_init(&bars.fullV[i], 1); mbar_init(&bars.emptyV[i], 1);
      mbar_init(&bars.s[i], 1); mbar_init(&bars.sfree[i], 4);
    }
    mbar_init(&bars.q, 1); mbar_init(&bars.p, 4); mbar_init(&bars.pv, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
  }
  if (warp == 1) tmem_alloc<TM_COLS>(&bars.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp == 0) {
    // ---------------------------------- TMA producer ------------------------------------
    mbar_arrive_expect_tx_ws(&bars.q, Q_BYTES);
    tma_load_4d_ws(smem + OFF_Q, &tmQ, 0, h, i0, b, &bars.q);
    tma_load_4d_ws(smem + OFF_Q + Q_BYTES / 2, &tmQ, 64, h, i0, b, &bars.q);
    // order of the copies = order in which their slots come free: K(t) [after S(t-2)], then V(t-1) [after PV(t-3)]
    auto load_k = [&](int t) {
      const int s = t & 1;
      if (t >= 2) mbar_wait_relaxed(&bars.emptyK[s], ((t >> 1) - 1) & 1);
      const int j0 = org + (t_lo + t) * BN;   // may be negative for the first tile: zero-filled, masked
      uint8_t* kd = smem + OFF_K + s * KT_BYTES_;
      mbar_arrive_expect_tx_ws(&bars.fullK[s], KT_BYTES_);
      tma_load_4d_ws(kd, &tmK, 0, hk, koff + j0, b, &bars.fullK[s]);
      tma_load_4d_ws(kd + KT_BYTES_ / 2, &tmK, 64, hk, koff + j0, b, &bars.fullK[s]);
    };
    auto load_v = [&](int t) {
      const int s = t & 1;
      if (t >= 2) mbar_wait_relaxed(&bars.emptyV[s], ((t >> 1) - 1) & 1);
      const int j0 = org + (t_lo + t) * BN;
      uint8_t* vd = smem + OFF_V + s * KT_BYTES_;
      mbar_arrive_expect_tx_ws(&bars.fullV[s], KT_BYTES_);
      tma_load_4d_ws(vd, &tmV, 0, hk, koff + j0, b, &bars.fullV[s]);
      tma_load_4d_ws(vd + KT_BYTES_ / 2, &tmV, 64, hk, koff + j0, b, &bars.fullV[s]);
    };
    load_k(0);
    for (int t = 1; t < n_tiles; ++t) {
      load_k(t);
      load_v(t - 1);
    }
    load_v(n_tiles - 1);
  } else if (warp == 1) {
    // ---------------------------------- MMA issuer --------------------------------------
    // All 32 lanes run converged; the elected lane issues (umma_*_ws) so descriptors stay in uniform registers.
    constexpr uint32_t idescS = umma_idesc_bf16(BM, BN, 0, 0);
    constexpr uint32_t idescO = umma_idesc_bf16(BM, HD, 0, 1);
    const uint32_t sb = smem_u32(smem);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint64_t dQ = umma_desc(sb + OFF_Q, 16, 1024, SWZ_128B);
    auto issue_s = [&](int t) {
      const int s = t & 1;
      const uint64_t dK = umma_desc(sb + OFF_K + s * KT_BYTES_, 16, 1024, SWZ_128B);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t qoff = ((j >> 2) * (Q_BYTES / 2) + (j & 3) * 32) >> 4;
        const uint32_t koff = ((j >> 2) * (KT_BYTES_ / 2) + (j & 3) * 32) >> 4;
        umma_bf16_ws(tm + TM_S + s * BN, dQ + qoff, dK + koff, idescS, j > 0);
      }
      umma_commit_ws(&bars.s[s]);
      umma_commit_ws(&bars.emptyK[s]);
    };
    mbar_wait(&bars.q, 0);
    mbar_wait(&bars.fullK[0], 0);
    tc_fence_after();
    issue_s(0);
    for (int t = 0; t < n_tiles; ++t) {
      if (t + 1 < n_tiles) {
        const int s1 = (t + 1) & 1;
        mbar_wait(&bars.fullK[s1], ((t + 1) >> 1) & 1);
        if (t + 1 >= 2) mbar_wait(&bars.sfree[s1], (((t + 1) >> 1) - 1) & 1);
        tc_fence_after();
        issue_s(t + 1);
      }
      mbar_wait(&bars.p, t & 1);
      mbar_wait(&bars.fullV[t & 1], (t >> 1) & 1);
      tc_fence_after();
      {
        const int s = t & 1;
        // V tile: MN-major (d contiguous), two 64-wide d panels 8 KiB apart, 8-key groups 1 KiB apart
        const uint64_t dV = umma_desc(sb + OFF_V + s * KT_BYTES_, KT_BYTES_ / 2, 1024, SWZ_128B);
#pragma unroll
        for (int j = 0; j < BN / 16; ++j)
          umma_bf16_ts_ws(tm + TM_O, tm + TM_S + s * BN + j * 8, dV + j * 128, idescO, (t > 0 || j > 0) ? 1u : 0u);
        umma_commit_ws(&bars.pv);
        umma_commit_ws(&bars.emptyV[s]);
      }
    }
  } else {
    // ---------------------------------- softmax / epilogue ------------------------------
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const int pos = i0 + row + shift;  // absolute key position of this query
    float m = -INFINITY, l = 0.f;   // running max (log2 domain, scaled) and row sum
    uint32_t r[32], r2[32];          // raw scores of keys 0..31 / 32..63 of the tile
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t & 1;
      const int j0 = org + (t_lo + t) * BN;
      mbar_wait(&bars.s[s], (t >> 1) & 1);
      tc_fence_after();
      tmem_ld32(tlane + TM_S + s * BN, r);
      tmem_ld32(tlane + TM_S + s * BN + 32, r2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.sfree[s]);
      // masks only on boundary tiles (CTA-uniform test)
      const bool need_mask = (j0 + BN - 1 > p_first) || (a.window > 0 && j0 < p_last - a.window + 1) ||
                             (j0 + BN > jmax) || (j0 < jmin);
      if (need_mask) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int ja = j0 + i, jb = j0 + 32 + i;
          const bool va = (ja >= jmin) && (ja <= pos) && (ja < jmax) && (a.window <= 0 || pos - ja < a.window);
          const bool vb = (jb >= jmin) && (jb <= pos) && (jb < jmax) && (a.window <= 0 || pos - jb < a.window);
          r[i] = va ? r[i] : 0xff800000u;   // -inf
          r2[i] = vb ? r2[i] : 0xff800000u;
        }
      }
      float mx = fmaxf(__uint_as_float(r[0]), __uint_as_float(r2[0]));
#pragma unroll
      for (int i = 1; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(r[i]), __uint_as_float(r2[i])));
      mx *= a.scale_log2;  // scale > 0, so the max commutes with it
      // lazy rescale: the running max only moves when it would grow by more than 2^THRESHOLD
      const bool grow = mx > m + RESCALE_THRESHOLD;  // also true for the first finite max (m = -inf)
      const float m_old = m;
      m = grow ? mx : m;
      const float m_eff = (m == -INFINITY) ? 0.f : m;
      // p = exp2(score * scale - m): one packed FFMA per two keys, MUFU.EX2 each, packed row sum.  All of it
      // happens BEFORE waiting for PV(t-1): only the store of P (single buffer) and the rare rescale of O
      // depend on that MMA, so the exponentials overlap with it.
      const float2 sc2 = make_float2(a.scale_log2, a.scale_log2), nm2 = make_float2(-m_eff, -m_eff);
      float2 sum2 = make_float2(0.f, 0.f);
      uint32_t w[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const uint32_t* src = (i < 16) ? (r + 2 * i) : (r2 + 2 * (i - 16));
        float2 v = __ffma2_rn(make_float2(__uint_as_float(src[0]), __uint_as_float(src[1])), sc2, nm2);
        if ((i & 7) < POLY) {
          v = exp2_fma_pair(v);
        } else {
          v.x = exp2f(v.x);
          v.y = exp2f(v.y);
        }
        sum2 = __fadd2_rn(sum2, v);
        w[i] = pack_bf16(v.x, v.y);
      }
      if (__any_sync(0xffffffffu, grow) && t > 0) {
        mbar_wait(&bars.pv, (t - 1) & 1);  // O complete up to tile t-1 (only the rare rescale needs it)
        tc_fence_after();
        const float f = grow ? exp2f(m_old - m) : 1.f;  // exp2(-inf) = 0 wipes an all-masked prefix
        l *= f;
        uint32_t ro[32];
#pragma unroll
        for (int c = 0; c < HD; c += 32) {
          tmem_ld32(tlane + TM_O + c, ro);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * f);
          tmem_st32(tlane + TM_O + c, ro);
        }
      }
      l += sum2.x + sum2.y;
      // P(t) as a TMEM A operand: lane = query row, word i = keys 2i, 2i+1, over the S buffer this warp has
      // already read (its other 32 columns stay dead until S(t+2) overwrites the buffer)
      tmem_st32(tlane + TM_S + s * BN, w);
      tmem_st_wait();
      tc_fence_before();
      // Stay within one phase of the PV barrier (a parity wait cannot tell phase n from phase n - 2): PV(t-1)
      // was issued a whole tile ago, so this is normally free, and PV(t) cannot retire before the arrival below.
      if (t > 0) mbar_wait(&bars.pv, (t - 1) & 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.p);
    }
    // epilogue: O / l -> bf16 -> global
    mbar_wait(&bars.pv, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    const int i = i0 + row;
    __nv_bfloat16* dst = a.o + (long long)b * a.o_sb + (long long)i * a.o_st + (long long)h * a.o_sh;
#pragma unroll
    for (int c = 0; c < HD; c += 32) {
      tmem_ld32(tlane + TM_O + c, r);
      tmem_ld_wait();
      if (i < i_end) {
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(r[v4 * 8 + 0]) * inv, __uint_as_float(r[v4 * 8 + 1]) * inv);
          w.y = pack_bf16(__uint_as_float(r[v4 * 8 + 2]) * inv, __uint_as_float(r[v4 * 8 + 3]) * inv);
          w.z = pack_bf16(__uint_as_float(r[v4 * 8 + 4]) * inv, __uint_as_float(r[v4 * 8 + 5]) * inv);
          w.w = pack_bf16(__uint_as_float(r[v4 * 8 + 6]) * inv, __uint_as_float(r[v4 * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + c + v4 * 8) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TM_COLS>(tmem);
}

// ---------------------------------------------------------------------------------------
// host side: tensor maps through the driver entry point (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [B, T, Hn, 128] bf16 tensor with element strides (sb, st, sh), innermost contiguous; box = 64 x 1 x rows x 1
bool make_map(CUtensorMap* m, const void* ptr, int B, int T, int Hn, long long sb, long long st, long long sh,
              int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)HD, (cuuint64_t)Hn, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)sh * 2, (cuuint64_t)st * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace

// q [B,Tq,Hq,128], k/v [B,Tk,Hkv,128], o [B,Tq,Hq,128]; strides in elements (batch, time, head).
// ring_state != nullptr: k, v are the physical rings [B, Tk = 2R, Hkv, 128] of a ring-buffer window cache and the
// visible keys are derived on the device (SwaArgs::ring_state); window must be > 0.
cudaError_t launch_swa_fwd(const void* q, const long long* qs, const void* k, const long long* ks, const void* v,
                           const long long* vs, void* o, const long long* os, int B, int Tq, int Tk, int Hq, int Hkv,
                           int window, float scale, const int* ring_state, int ring_R, long long key_pos0,
                           cudaStream_t stream, const int* vt_tok0 = nullptr, const int* vt_lo = nullptr,
                           const int* vt_hi = nullptr, int vt_tiles = 0) {
  // IVL_SWA_POLY (developer knob): pairs out of 8 whose exponentials run on the FMA pipe
  int poly = SWA_POLY_DEFAULT;
  if (const char* e = getenv("IVL_SWA_POLY")) poly = atoi(e);
  void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, SwaArgs) =
      poly <= 0 ? swa_fwd_kernel<0> : poly <= 2 ? swa_fwd_kernel<2> : poly == 3 ? swa_fwd_kernel<3>
      : poly == 4 ? swa_fwd_kernel<4> : swa_fwd_kernel<5>;
  static std::atomic<bool> configured_dev[64];   // function attributes are per device
  int dev_ = 0;
  if (cudaError_t e = cudaGetDevice(&dev_)) return e;
  if (dev_ < 0 || dev_ >= 64) return cudaErrorInvalidDevice;
  std::atomic<bool>& configured = configured_dev[dev_];
  if (!configured.load(std::memory_order_acquire)) {
    for (auto f : {swa_fwd_kernel<0>, swa_fwd_kernel<2>, swa_fwd_kernel<3>, swa_fwd_kernel<4>, swa_fwd_kernel<5>}) {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, SWA_SMEM);
      if (e != cudaSuccess) return e;
    }
    configured.store(true, std::memory_order_release);
  }
  CUtensorMap tq, tk, tv;
  if (!make_map(&tq, q, B, Tq, Hq, qs[0], qs[1], qs[2], BM) || !make_map(&tk, k, B, Tk, Hkv, ks[0], ks[1], ks[2], BN) ||
      !make_map(&tv, v, B, Tk, Hkv, vs[0], vs[1], vs[2], BN))
    return cudaErrorInvalidValue;
  SwaArgs a;
  a.o = static_cast<__nv_bfloat16*>(o);
  a.o_sb = os[0]; a.o_st = os[1]; a.o_sh = os[2];
  a.Tq = Tq; a.Tk = Tk; a.Hq = Hq; a.group = Hq / Hkv;
  a.window = (window > 0 && (Tk > window || vt_tok0)) ? window : 0;  // HF glue passes the window only when key_len > W
  a.vt_tok0 = vt_tok0; a.vt_lo = vt_lo; a.vt_hi = vt_hi;
  a.ring_state = ring_state; a.ring_R = ring_R; a.ring_W = window;
  a.kalign = (int)(key_pos0 & (BN - 1));
  a.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid(Hq, vt_tok0 ? vt_tiles : (Tq + BM - 1) / BM, B);
  kern<<<grid, SWA_THREADS, SWA_SMEM, stream>>>(tq, tk, tv, a);
  return cudaGetLastError();
}

}  // namespace ivl
