// Gated DeltaNet inter-chunk scan, TRANSPOSED form: value columns on the TMEM lanes.
//
// Same job as gdn_scan.cu (the serial part of the chunked delta rule; replaces
// chunk_gated_delta_rule_fwd_h + chunk_fwd_o of the reference,
// src/llamafactory/model/fla/ops/common/chunk_delta_h.py:32-124,247-318 and ops/common/chunk_o.py:32-114,456-497),
// but one persistent CTA owns 128 value columns of one (batch, head) -- two CTAs per head instead of
// four or eight -- and keeps the state TRANSPOSED: S^T [128 value columns x 128 key dims].
//
// Why: in the row-major form the per-head operands ([-Wg;Qg], Kt, P: 56 KiB per chunk) are the A operands of
// every MMA and are re-read from shared memory once per 32-column slice; with four or eight slices per head
// the shared-memory port (128 B/cycle: TMA fills + operand reads) is the measured limit
// (profiles/r01e_summary.md).  Transposed, the state and v_new are the A operands and live in TENSOR MEMORY
// (tcgen05.mma with a TMEM A operand); the per-head operands are the B operands, read from shared memory once per
// 128 value columns.  Per chunk and CTA the port moves ~160 KiB instead of ~256 KiB per 64 columns, and the
// scan needs 2 H CTAs instead of 4 H or 8 H, which leaves the other SMs to the concurrently running pre-pass.
//
// Per chunk c (all operands are the images gdn_prep.cu writes; V arrives straight from the caller's tensor):
//   U part  DV    = V^T Au^T                    SS  M128 N64  K64    (A: value tile, MN-major, 128B-swizzled TMA tile)
//   W part  DV   += bf16(S^T) (-Wg)^T           TS  M128 N64  K128   -> v_new^T = U^T - S^T Wg^T
//   O part  DO    = bf16(S^T) Qg^T              TS  M128 N64  K128
//   epi V   v_new^T -> bf16 -> TMEM (A operand of the next two products)
//   B part  DS    = v_new^T Kt                  TS  M128 N128 K64
//   C part  DO   += v_new^T P^T                 TS  M128 N64  K64    -> O^T
//   epi S   S^T = gamma S^T + DS (fp32, in REGISTERS for the whole sequence) -> bf16 -> TMEM A operand
//   epi O   DO -> bf16 -> global
// The U part of chunk c+1 does not depend on the state and is issued right behind chunk c's products, so the
// serial chain per chunk is  W part -> epi V -> B part -> epi S.
// Warp roles: warp 0 copies (TMA engine; follows the pre-pass's ready flags), warp 1 issues the MMAs, warps 2..5
// are the v_new / output epilogue, warps 6..13 own the state (TMEM lane quadrant = warp % 4 in both groups; two
// state warps per quadrant split the 128 key dims).
#include <atomic>
#include <cuda.h>

#include "gdn_layout.cuh"
#include "sm100.cuh"

namespace ivl {

namespace {

struct TCfg {
  static constexpr int THREADS = 448;   // 14 warps: copy, MMA, 4 v_new/output, 8 state
  static constexpr int NA = 2, NK = 2;
  // A slot: operands that are dead once the W and O parts have retired (early in the step)
  static constexpr uint32_t A_BW = 0;                         // [-Wg ; Qg]  32 KiB, K-major, no swizzle
  static constexpr uint32_t A_V = A1_BYTES;                   // value tile: 2 panels [64 tok][64 val], 128B swizzle
  static constexpr uint32_t V_PANEL = 64 * 64 * 2;            // 8 KiB
  static constexpr uint32_t A_AU = A1_BYTES + 2 * V_PANEL;    // Au 8 KiB, K-major, no swizzle
  static constexpr uint32_t ASLOT = A_AU + AU_BYTES;          // 56 KiB
  static constexpr uint32_t A_TX = ASLOT;
  // K slot: operands of the B and C parts (dead at the end of the step): P | Kt | gamma, one copy
  static constexpr uint32_t K_P = 0;
  static constexpr uint32_t K_KT = P_BYTES;
  static constexpr uint32_t K_TAIL = P_BYTES + KT_BYTES;
  static constexpr uint32_t K_TX = P_BYTES + KT_BYTES + TAIL_BYTES;
  static constexpr uint32_t KSLOT = P_BYTES + KT_BYTES + 1024;   // 25 KiB
  static constexpr uint32_t OFF_A = 0;
  static constexpr uint32_t OFF_K = NA * ASLOT;
  static constexpr uint32_t OFF_BARS = OFF_K + NK * KSLOT;
  static constexpr uint32_t SMEM = OFF_BARS + 512 + 1024;        // + alignment slack
  // tensor memory (512 columns): fp32 accumulators and the two bf16 A operands
  static constexpr uint32_t TM_DS = 0;       // 128: v_new^T Kt          (state increment)
  static constexpr uint32_t TM_SB = 128;     //  64: bf16 S^T            (A operand, K = 128)
  static constexpr uint32_t TM_DV = 192;     // 2 x 64: v_new^T accumulators
  static constexpr uint32_t TM_VB = 320;     //  32: bf16 v_new^T        (A operand, K = 64)
  static constexpr uint32_t TM_DO = 384;     // 2 x 64: O^T accumulators
  static constexpr uint32_t TM_COLS = 512;
  static_assert(ASLOT % 1024 == 0 && KSLOT % 1024 == 0 && A_V % 1024 == 0, "swizzled tiles need 1 KiB alignment");
  static_assert(SMEM <= 232448, "exceeds 227 KiB");
};

// Developer-only timeline probe (compiled in with -DIVL_TRACE by tools/trace_tscan.py; never in the product build)
#ifdef IVL_TRACE
__device__ long long ivl_ttrace_buf[64 * 16];
#define TTR(slot)                                                                                        \
  do {                                                                                                   \
    if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && c >= 1000 && c < 1064)                        \
      ivl_ttrace_buf[(c - 1000) * 16 + (slot)] = clock64();                                              \
  } while (0)
#else
#define TTR(slot) do { } while (0)
#endif

struct TBars {
  uint64_t fullA[2], emptyA[2], fullK[2], emptyK[2];
  uint64_t sb, vb, ds, dv[2], dofull[2], dofree[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t ld_acquire_gpu_t(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// 448 threads: registers are granted per four warps, so the budget is 65536 / 512 = 128 per thread -- a state thread
// keeps 64 fp32 state entries (half a row of S^T) for the whole sequence plus one 32-column staging buffer
__global__ void __launch_bounds__(TCfg::THREADS, 1)
gdn_scan_t_kernel(const __grid_constant__ CUtensorMap tmV, GdnWorkspace ws, GdnVarlen vl, const void* __restrict__ h0,
                  int h0_dtype, __nv_bfloat16* __restrict__ o, void* __restrict__ ht, int ht_dtype, int T, int H,
                  int NTROW) {
  using C = TCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  TBars& bars = *reinterpret_cast<TBars*>(smem + C::OFF_BARS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int vh = blockIdx.x, h = blockIdx.y;               // value half, head
  const bool varlen = vl.chunk_tok0 != nullptr;
  const int b = varlen ? 0 : blockIdx.z;
  const int seq = blockIdx.z;
  const int cb = varlen ? __ldg(vl.seq_chunk_begin + seq) : 0;
  const int NT = varlen ? __ldg(vl.seq_chunk_begin + seq + 1) - cb : NTROW;
  const size_t ch0 = ((size_t)b * H + h) * NTROW + cb;
  const size_t slot0 = ((size_t)b * H + h) * ws.ring;
  const int ring = ws.ring;
  const int col0 = vh * 128;
  if (NT <= 0) {
    // empty sequence: the final state is the initial state
    if (ht != nullptr) {
      const size_t base = ((size_t)seq * H + h) * GDN_K * GDN_V;
      for (int i = tid; i < GDN_K * 128; i += C::THREADS) {
        const size_t off = base + (size_t)(i >> 7) * GDN_V + col0 + (i & 127);
        const float x = h0 == nullptr ? 0.f
                        : (h0_dtype == 0 ? static_cast<const float*>(h0)[off]
                                         : __bfloat162float(static_cast<const __nv_bfloat16*>(h0)[off]));
        if (ht_dtype == 0) static_cast<float*>(ht)[off] = x;
        else static_cast<__nv_bfloat16*>(ht)[off] = __float2bfloat16(x);
      }
    }
    return;
  }
  const uint8_t* blob = ws.blob + slot0 * BLOB_BYTES;
  const uint8_t* aublob = ws.ublob + slot0 * (GDN_NS * UBLOB_BYTES);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.fullA[s], 1); mbar_init(&bars.emptyA[s], 1);
      mbar_init(&bars.fullK[s], 1);
      mbar_init(&bars.emptyK[s], 1 + 8);   // the C part has retired + the eight state warps have read gamma
      mbar_init(&bars.dv[s], 1); mbar_init(&bars.dofull[s], 1); mbar_init(&bars.dofree[s], 4);
    }
    mbar_init(&bars.sb, 8); mbar_init(&bars.vb, 4); mbar_init(&bars.ds, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) tmem_alloc<C::TM_COLS>(&bars.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp == 0) {
    // ------------------------------- copy warp (TMA engine) ---------------------------
    const uint32_t* ready = ws.ready + ch0;
    uint32_t* progress = ws.progress + ((size_t)b * H + h) * GDN_NS + vh;
    int known = 0;  // chunks [0, known) are published
    for (int c = 0; c < NT; ++c) {
      if (c >= known) {
        long long spins = 0;
        do {
          const int idx = known + lane;
          const uint32_t f = (idx < NT) ? ld_acquire_gpu_t(ready + idx) : 0u;
          const uint32_t m = __ballot_sync(0xffffffffu, f != 0u);
          known += (m == 0xffffffffu) ? 32 : (__ffs(~m) - 1);
          if (c >= known) {
            __nanosleep(200);
            if (++spins > (1ll << 24)) asm volatile("trap;");  // the pre-pass never ran: fail loudly, do not hang
          }
        } while (c >= known);
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      const int sa = c % C::NA, sk = c % C::NK;
      const size_t cs = (size_t)((cb + c) % ring);  // image slot of chunk c
      const int tok0 = varlen ? __ldg(vl.chunk_tok0 + cb + c) : c * GDN_C;
      if (c >= C::NA) mbar_wait(&bars.emptyA[sa], (c / C::NA - 1) & 1);
      uint8_t* as = smem + C::OFF_A + sa * C::ASLOT;
      mbar_arrive_expect_tx_ws(&bars.fullA[sa], C::A_TX);
      bulk_g2s_ws(as + C::A_BW, blob + cs * BLOB_BYTES + BLOB_OFF_A1, A1_BYTES, &bars.fullA[sa]);
      bulk_g2s_ws(as + C::A_AU, aublob + cs * (GDN_NS * UBLOB_BYTES), AU_BYTES, &bars.fullA[sa]);
      tma_load_4d_ws(as + C::A_V, &tmV, col0, h, tok0, b, &bars.fullA[sa]);
      tma_load_4d_ws(as + C::A_V + C::V_PANEL, &tmV, col0 + 64, h, tok0, b, &bars.fullA[sa]);
      if (c >= C::NK) {
        mbar_wait(&bars.emptyK[sk], (c / C::NK - 1) & 1);
        // every product that read chunk c - NK has retired: its image slot may be overwritten (ring hand-off)
        if (lane == 0 && ring < NTROW)
          asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(progress), "r"((uint32_t)(c - C::NK + 1)) : "memory");
      }
      uint8_t* ks = smem + C::OFF_K + sk * C::KSLOT;
      mbar_arrive_expect_tx_ws(&bars.fullK[sk], C::K_TX);
      bulk_g2s_ws(ks + C::K_P, blob + cs * BLOB_BYTES + BLOB_OFF_P, C::K_TX, &bars.fullK[sk]);
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------------
    // All 32 lanes run converged; the elected lane issues (umma_*_ws), operands stay in uniform registers.
    constexpr uint32_t idescU = umma_idesc_bf16(128, 64, /*a_mn=*/1, /*b_mn=*/0);
    constexpr uint32_t idesc64 = umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idescB = umma_idesc_bf16(128, 128, 0, /*b_mn=*/1);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    auto issue_u = [&](int c) {   // DV[c & 1] = V^T Au^T
      const uint32_t as = sbase + C::OFF_A + (c % C::NA) * C::ASLOT;
      const uint64_t dV = umma_desc(as + C::A_V, C::V_PANEL, 1024, SWZ_128B);
      const uint64_t dAu = umma_desc(as + C::A_AU, 128, 1024, SWZ_NONE);
#pragma unroll
      for (int j = 0; j < 4; ++j) umma_bf16_ws(tm + C::TM_DV + (c & 1) * 64, dV + j * 128, dAu + j * 16, idescU, j > 0);
    };
    mbar_wait(&bars.fullA[0], 0);
    tc_fence_after();
    issue_u(0);
    for (int c = 0; c < NT; ++c) {
      const int sa = c % C::NA, sk = c % C::NK, buf = c & 1;
      const uint32_t as = sbase + C::OFF_A + sa * C::ASLOT, ks = sbase + C::OFF_K + sk * C::KSLOT;
      const uint64_t dW = umma_desc(as + C::A_BW, 128, 2048, SWZ_NONE);            // rows 0..63: -Wg
      const uint64_t dQ = umma_desc(as + C::A_BW + 8 * 2048, 128, 2048, SWZ_NONE); // rows 64..127: Qg
      const uint64_t dKt = umma_desc(ks + C::K_KT, 128, 1024, SWZ_NONE);
      const uint64_t dP = umma_desc(ks + C::K_P, 128, 1024, SWZ_NONE);
      const uint32_t dvb = tm + C::TM_DV + buf * 64, dob = tm + C::TM_DO + buf * 64;
      mbar_wait(&bars.sb, c & 1);                                    // bf16 S_c^T is in tensor memory
      if (c >= 2) mbar_wait(&bars.dofree[buf], ((c >> 1) - 1) & 1);  // O accumulator of chunk c - 2 has been read
      tc_fence_after();
      TTR(0);
#pragma unroll
      for (int j = 0; j < 8; ++j) umma_bf16_ts_ws(dvb, tm + C::TM_SB + j * 8, dW + j * 16, idesc64, 1);
      umma_commit_ws(&bars.dv[buf]);
#pragma unroll
      for (int j = 0; j < 8; ++j) umma_bf16_ts_ws(dob, tm + C::TM_SB + j * 8, dQ + j * 16, idesc64, j > 0);
      umma_commit_ws(&bars.emptyA[sa]);
      TTR(1);
      mbar_wait(&bars.fullK[sk], (c / C::NK) & 1);
      mbar_wait(&bars.vb, c & 1);                                    // bf16 v_new^T is in tensor memory
      tc_fence_after();
      TTR(2);
#pragma unroll
      for (int j = 0; j < 4; ++j) umma_bf16_ts_ws(tm + C::TM_DS, tm + C::TM_VB + j * 8, dKt + j * 16, idescB, j > 0);
      umma_commit_ws(&bars.ds);
#pragma unroll
      for (int j = 0; j < 4; ++j) umma_bf16_ts_ws(dob, tm + C::TM_VB + j * 8, dP + j * 16, idesc64, 1);
      umma_commit_ws(&bars.dofull[buf]);
      umma_commit_ws(&bars.emptyK[sk]);
      TTR(3);
      if (c + 1 < NT) {
        mbar_wait(&bars.fullA[(c + 1) % C::NA], ((c + 1) / C::NA) & 1);
        tc_fence_after();
        issue_u(c + 1);
      }
      TTR(4);
    }
  } else if (warp < 6) {
    // ------------------------------- v_new / output epilogue --------------------------
    const int quad = warp & 3;
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const int col = col0 + quad * 32 + lane;     // this thread's value column
    uint32_t r[32], r2[32], w[32];
    auto output = [&](int c) {
      const int buf = c & 1;
      mbar_wait(&bars.dofull[buf], (c >> 1) & 1);
      tc_fence_after();
      tmem_ld32(tlane + C::TM_DO + buf * 64, r);
      tmem_ld32(tlane + C::TM_DO + buf * 64 + 32, r2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.dofree[buf]);
      const int tok0 = varlen ? __ldg(vl.chunk_tok0 + cb + c) : c * GDN_C;
      const int valid = varlen ? __ldg(vl.chunk_valid + cb + c) : min(GDN_C, T - tok0);
      // one 64-byte row segment per warp and token (pointers advance by one token row: keeps them out of the
      // register file -- 64 precomputed addresses would spill)
      const size_t tstride = (size_t)H * GDN_V;
      __nv_bfloat16* p0 = o + (((size_t)b * T + tok0) * H + h) * GDN_V + col;
      __nv_bfloat16* p1 = p0 + 32 * tstride;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i < valid) *p0 = __float2bfloat16(__uint_as_float(r[i]));
        if (i + 32 < valid) *p1 = __float2bfloat16(__uint_as_float(r2[i]));
        asm volatile("" : "+l"(p0), "+l"(p1));   // keep the increments serial
        p0 += tstride;
        p1 += tstride;
      }
    };
    for (int c = 0; c < NT; ++c) {
      const int buf = c & 1;
      mbar_wait(&bars.dv[buf], (c >> 1) & 1);
      tc_fence_after();
      if (quad == 0) TTR(5);
      tmem_ld32(tlane + C::TM_DV + buf * 64, r);
      tmem_ld32(tlane + C::TM_DV + buf * 64 + 32, r2);
      tmem_ld_wait();
      if (quad == 0) TTR(6);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        w[i] = pack_bf16(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
        w[16 + i] = pack_bf16(__uint_as_float(r2[2 * i]), __uint_as_float(r2[2 * i + 1]));
      }
      tmem_st32(tlane + C::TM_VB, w);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.vb);
      if (quad == 0) TTR(7);
      if (c > 0) output(c - 1);
      if (quad == 0) TTR(8);
    }
    output(NT - 1);
  } else {
    // ------------------------------- state warps ---------------------------------------
    // two warps per TMEM lane quadrant: warp (quad, half) owns key dims 64 * half .. + 63 of its 32 value columns
    const int quad = warp & 3, half = (warp - 6) >> 2;
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const int col = col0 + quad * 32 + lane;
    float S[64];                                 // half a row of S^T (this value column, 64 key dims), fp32
    const size_t sbase_off = (((size_t)seq * H + h) * GDN_K + half * 64) * GDN_V + col;
    if (h0 == nullptr) {
#pragma unroll
      for (int i = 0; i < 64; ++i) S[i] = 0.f;
    } else if (h0_dtype == 0) {
      const float* p = static_cast<const float*>(h0) + sbase_off;
#pragma unroll
      for (int i = 0; i < 64; ++i) S[i] = __ldg(p + (size_t)i * GDN_V);
    } else {
      const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(h0) + sbase_off;
#pragma unroll
      for (int i = 0; i < 64; ++i) S[i] = __bfloat162float(p[(size_t)i * GDN_V]);
    }
    uint32_t r[32];
    auto publish = [&]() {   // bf16 S^T -> TMEM A operand (word i = key dims 2i, 2i+1)
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = pack_bf16(S[2 * i], S[2 * i + 1]);
      tmem_st32(tlane + C::TM_SB + half * 32, r);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.sb);
    };
    publish();
    for (int c = 0; c < NT; ++c) {
      const int sk = c % C::NK;
      mbar_wait(&bars.fullK[sk], (c / C::NK) & 1);
      const float gamma = *reinterpret_cast<const float*>(smem + C::OFF_K + sk * C::KSLOT + C::K_TAIL);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.emptyK[sk]);
      mbar_wait(&bars.ds, c & 1);
      tc_fence_after();
      if (warp == 8) TTR(9);
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        tmem_ld32(tlane + C::TM_DS + half * 64 + p * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) S[p * 32 + i] = fmaf(gamma, S[p * 32 + i], __uint_as_float(r[i]));
      }
      if (warp == 8) TTR(10);
      if (c + 1 < NT) {
        publish();
      } else {
        tc_fence_before();
      }
      if (warp == 8) TTR(11);
    }
    if (ht != nullptr) {
      if (ht_dtype == 0) {
        float* p = static_cast<float*>(ht) + sbase_off;
#pragma unroll
        for (int i = 0; i < 64; ++i) p[(size_t)i * GDN_V] = S[i];
      } else {
        __nv_bfloat16* p = static_cast<__nv_bfloat16*>(ht) + sbase_off;
#pragma unroll
        for (int i = 0; i < 64; ++i) p[(size_t)i * GDN_V] = __float2bfloat16(S[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TM_COLS>(tmem);
}

typedef CUresult (*EncodeTiledFnT)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFnT encode_fn_t() {
  static std::atomic<EncodeTiledFnT> fn{nullptr};
  EncodeTiledFnT f = fn.load(std::memory_order_acquire);
  if (!f) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      f = reinterpret_cast<EncodeTiledFnT>(p);
      fn.store(f, std::memory_order_release);
    }
  }
  return f;
}

}  // namespace

// v: the caller's value tensor [B, T, H, 256] bf16 (dense).  The scan reads its tiles directly (box = 64 value
// columns x 64 tokens, 128-byte swizzle); rows past T are zero-filled by the TMA engine.
cudaError_t launch_gdn_scan_t(const void* v, const GdnWorkspace& ws, const GdnVarlen& vl, int ntrow, int nseq, int B,
                              const void* h0, int h0_dtype, void* o, void* ht, int ht_dtype, int T, int H,
                              cudaStream_t stream) {
  using C = TCfg;
  static std::atomic<bool> configured[64];
  int dev = 0;
  if (cudaError_t e = cudaGetDevice(&dev)) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!configured[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(gdn_scan_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    configured[dev].store(true, std::memory_order_release);
  }
  EncodeTiledFnT enc = encode_fn_t();
  if (!enc) return cudaErrorNotSupported;
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)GDN_V, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)GDN_V * 2, (cuuint64_t)H * GDN_V * 2, (cuuint64_t)T * H * GDN_V * 2};
  cuuint32_t box[4] = {64, 1, 64, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v), dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorInvalidValue;
  dim3 grid(2, H, nseq);
  gdn_scan_t_kernel<<<grid, C::THREADS, C::SMEM, stream>>>(tm, ws, vl, h0, h0_dtype, static_cast<__nv_bfloat16*>(o), ht,
                                                           ht_dtype, T, H, ntrow);
  return cudaGetLastError();
}

#ifdef IVL_TRACE
extern "C" __attribute__((visibility("default"))) int ivl_debug_read_ttrace(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, ivl_ttrace_buf, sizeof(long long) * n);
}
#endif

}  // namespace ivl
