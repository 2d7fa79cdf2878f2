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
#include <mutex>
#include <cuda.h>

#include "gdn_layout.cuh"
#include "sm100.cuh"

namespace ivl {

namespace {

struct TCfg {
  static constexpr int THREADS = 448;   // 14 warps: copy, MMA, 4 v_new/output, 8 state
#ifndef IVL_TSCAN_NA
#define IVL_TSCAN_NA 2
#endif
  static constexpr int NA = IVL_TSCAN_NA, NK = 2;   // NA: 2 or 3 A slots (developer knob, tools/dev_tscan.py)
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
  uint64_t fullA[3], emptyA[3], fullK[2], emptyK[2];
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
  // resident: with an image ring the pre-pass is only launched once every scan CTA has checked in (ivl_abi.cu)
  if (tid == 0) atomicAdd(ws.checkin, 1u);
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
    for (int s = 0; s < C::NA; ++s) { mbar_init(&bars.fullA[s], 1); mbar_init(&bars.emptyA[s], 1); }
    for (int s = 0; s < 2; ++s) {
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



// =====================================================================================================
// Lag form of the transposed scan (operands from gdn_prep_kernel<2>): the serial chain is shortened from
//   W part -> epi V -> B part -> epi S          (state -> v_new -> state, two tensor / epilogue round trips per chunk)
// to
//   RC part -> epi V                            (v_new_c -> v_new_{c+1}: one product with K = 64 per chunk)
// by expanding the state inside v_new_{c+1} = U_{c+1} - Wg_{c+1} S_{c+1} with S_{c+1} = gamma_c S_c + Kt_c^T v_new_c:
//   v_new_{c+1}^T = U_{c+1}^T - bf16(S_c^T) (gamma_c Wg_{c+1})^T - bf16(v_new_c^T) (Wg_{c+1} Kt_c^T)^T .
// The state update (B part, epi S) and the outputs leave the chain; they only have to keep up on average.
//
// A tcgen05.mma with a TMEM A operand costs ~64 cycles whatever its N <= 128 (measured on the first version of this
// kernel: N = 64 products ran at 56-64 cycles per MMA, i.e. at half the tensor rate), so products that share an A
// operand are merged into ONE N = 128 MMA chain over operands stacked in shared memory and accumulators adjacent
// in tensor memory.  STEP k (k = 0 .. NT-1) owns the accumulator pair PAIR[k & 1] = [ DV: v_new_{k+1}^T | DO: O_k^T ]:
//   U part    PAIR.DV  = V_{k+1}^T Au_{k+1}^T                              SS  N64  K64   (issued two steps ahead)
//   XO part   PAIR    += bf16(S_k^T) [ -gamma_k Wg_{k+1} ; Qg_k ]^T        TS  N128 K128  after S_k is published
//   B part    DS       = bf16(v_new_k^T) Kt_k                              TS  N128 K64   after v_new_k is published
//   RC part   PAIR    += bf16(v_new_k^T) [ -R_{k+1} ; P_k ]^T              TS  N128 K64   after v_new_k and the XO part
//   epi V     PAIR.DV -> bf16 -> TMEM (v_new_{k+1});  PAIR.DO -> bf16 -> global (O_k), then PAIR.DO is zeroed
//   epi S     S^T = gamma_k S^T + DS (fp32, registers) -> bf16 -> TMEM (S_{k+1})
// (chunk 0's v_new = U_0 - Wg_0 S_0 is an extra N = 64 product in the prologue; the last step has no chunk k+1 and
// uses N = 64 products over the Qg / P halves.)  Two MMA-issuing warps, one per trigger: warp 2 follows the state
// (XO, U), warp 1 follows v_new (B, RC).  tcgen05.commit only tracks the issuing thread's MMAs; the summation order
// inside an accumulator is fixed (U, XO, RC) by making warp 1 wait for the XO part's commit, so results do not
// depend on timing.
// Shared memory: U slots (value tile + Au of a chunk, 24 KiB, three deep) and step slots ([-gamma Wg_{k+1}; Qg_k] |
// [-R_{k+1}; P_k] | Kt_k | gamma_k, 64 KiB + 128 B, two deep), each ring with its own copy warp.
// =====================================================================================================
struct T2Cfg {
  static constexpr int THREADS = 512;   // 16 warps: U copy, 2 x MMA, 4 v_new/output, 8 state, step copy
  static constexpr int NU = 3, NS = 2;
  static constexpr uint32_t U_V = 0;                          // value tile, 2 swizzled panels, 16 KiB
  static constexpr uint32_t V_PANEL = 8192;
  static constexpr uint32_t U_AU = 16384;                     // Au, 8 KiB
  static constexpr uint32_t USLOT = 24576;
  // every operand is a 128-byte-swizzled tile (1 KiB aligned); K-major tiles wider than 64 are two 64-wide panels
  static constexpr uint32_t S_XO = 0;                         // 2 panels x [rows 0..63: -gamma Wg_{k+1}; rows 64..127: Qg_k] (32 KiB)
  static constexpr uint32_t XO_PANEL = 16384;
  static constexpr uint32_t S_RC = 32768;                     // rows 0..63: -R_{k+1}; rows 64..127: P_k (16 KiB)
  static constexpr uint32_t S_KT = 49152;                     // Kt_k: 2 panels of 64 key dims x [64 tok][128 B] | gamma_k (128 B)
  static constexpr uint32_t KT_PANEL = 8192;
  static constexpr uint32_t S_TAIL = S_KT + KT_BYTES;
  static constexpr uint32_t SSLOT = S_TAIL + 1024;            // 65 KiB (swizzled tiles: slots stay 1 KiB aligned)
  static constexpr uint32_t OFF_U = 0;
  static constexpr uint32_t OFF_S = NU * USLOT;
  static constexpr uint32_t OFF_BARS = OFF_S + NS * SSLOT;
  static constexpr uint32_t SMEM = OFF_BARS + 512 + 1024;
  static constexpr uint32_t TM_DS = 0;       // 128
  static constexpr uint32_t TM_SB = 128;     //  64
  static constexpr uint32_t TM_PAIR = 192;   // 2 x (64 DV + 64 DO)
  static constexpr uint32_t TM_VB = 448;     // 2 x 32
  static constexpr uint32_t TM_COLS = 512;
  static_assert(USLOT % 1024 == 0 && OFF_S % 1024 == 0 && SSLOT % 1024 == 0 && S_RC % 1024 == 0 && S_KT % 1024 == 0, "alignment");
  static_assert(SMEM <= 232448, "exceeds 227 KiB");
};

struct T2Bars {
  uint64_t fullU[3], emptyU[3], fullS[2], emptyS[2];
  uint64_t sb, vb, ds, dsfree, w0full, w0done, dv0, xo[2], rc[2], dvfree[2], dofree[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(T2Cfg::THREADS, 1)
gdn_scan_t2_kernel(const __grid_constant__ CUtensorMap tmV, GdnWorkspace ws, GdnVarlen vl, const void* __restrict__ h0,
                   int h0_dtype, __nv_bfloat16* __restrict__ o, void* __restrict__ ht, int ht_dtype, int T, int H,
                   int NTROW) {
  using C = T2Cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  T2Bars& bars = *reinterpret_cast<T2Bars*>(smem + C::OFF_BARS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int vh = blockIdx.x, h = blockIdx.y;
  const bool varlen = vl.chunk_tok0 != nullptr;
  const int b = varlen ? 0 : blockIdx.z;
  const int seq = blockIdx.z;
  const int cb = varlen ? __ldg(vl.seq_chunk_begin + seq) : 0;
  const int NT = varlen ? __ldg(vl.seq_chunk_begin + seq + 1) - cb : NTROW;
  const size_t ch0 = ((size_t)b * H + h) * NTROW + cb;
  const size_t slot0 = ((size_t)b * H + h) * ws.ring;
  const int ring = ws.ring;
  const int col0 = vh * 128;
  // resident: with an image ring the pre-pass is only launched once every scan CTA has checked in (ivl_abi.cu)
  if (tid == 0) atomicAdd(ws.checkin, 1u);
  if (NT <= 0) {
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
    for (int s = 0; s < 3; ++s) { mbar_init(&bars.fullU[s], 1); mbar_init(&bars.emptyU[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.fullS[s], 1);
      mbar_init(&bars.emptyS[s], 2 + 8);   // XO part (warp 2), B + RC parts (warp 1) retired, 8 state warps read gamma
      mbar_init(&bars.xo[s], 1); mbar_init(&bars.rc[s], 1);
      mbar_init(&bars.dvfree[s], 4); mbar_init(&bars.dofree[s], 4);
    }
    mbar_init(&bars.sb, 8); mbar_init(&bars.vb, 4); mbar_init(&bars.ds, 1); mbar_init(&bars.dsfree, 8);
    mbar_init(&bars.w0full, 1); mbar_init(&bars.w0done, 1); mbar_init(&bars.dv0, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) tmem_alloc<C::TM_COLS>(&bars.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;
  if (warp >= 3 && warp < 7) {
    // the O halves of both accumulator pairs start at zero (every product accumulates)
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0u;
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      tmem_st32(tl + C::TM_PAIR + p * 128 + 64, z);
      tmem_st32(tl + C::TM_PAIR + p * 128 + 96, z);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr uint32_t idescU = umma_idesc_bf16(128, 64, /*a_mn=*/1, /*b_mn=*/0);
  constexpr uint32_t idesc64 = umma_idesc_bf16(128, 64, 0, 0);
  constexpr uint32_t idesc128 = umma_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t idescB = umma_idesc_bf16(128, 128, 0, /*b_mn=*/1);

  if (warp == 0 || warp == 15) {
    // ------------------------------- copy warps (TMA engine) ---------------------------
    // warp 0 streams the U slots (chunk c needs chunk c published), warp 15 the step slots (step k needs chunk
    // k + 1 published): two rings recycled at different points of a step, each with its own prefetch distance
    const bool uwarp = warp == 0;
    const uint32_t* ready = ws.ready + ch0;
    uint32_t* progress = ws.progress + ((size_t)b * H + h) * GDN_NS + vh;
    int known = 0;
    auto need = [&](int c) {   // block until chunks [0, c] are published by the pre-pass
      if (c < known) return;
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
    };
    if (uwarp) {
      for (int c = 0; c < NT; ++c) {
        need(c);
        const int su = c % C::NU;
        const size_t cs = (size_t)((cb + c) % ring);
        const int tok0 = varlen ? __ldg(vl.chunk_tok0 + cb + c) : c * GDN_C;
        if (c >= C::NU) mbar_wait(&bars.emptyU[su], (c / C::NU - 1) & 1);
        uint8_t* us = smem + C::OFF_U + su * C::USLOT;
        mbar_arrive_expect_tx_ws(&bars.fullU[su], C::USLOT);
        bulk_g2s_ws(us + C::U_AU, aublob + cs * (GDN_NS * UBLOB_BYTES), AU_BYTES, &bars.fullU[su]);
        tma_load_4d_ws(us + C::U_V, &tmV, col0, h, tok0, b, &bars.fullU[su]);
        tma_load_4d_ws(us + C::U_V + C::V_PANEL, &tmV, col0 + 64, h, tok0, b, &bars.fullU[su]);
      }
    } else {
      // chunk 0's own Wg rows (for v_new_0 = U_0 - Wg_0 S_0) borrow the XO area of step slot 1 until that product
      // has retired
      need(0);
      mbar_arrive_expect_tx_ws(&bars.w0full, 16384);
      for (int pn = 0; pn < 2; ++pn)   // rows 0..63 of both 64-wide panels
        bulk_g2s_ws(smem + C::OFF_S + C::SSLOT + C::S_XO + pn * C::XO_PANEL,
                    blob + (size_t)(cb % ring) * BLOB_BYTES + BLOB_OFF_A1 + pn * C::XO_PANEL, 8192, &bars.w0full);
      for (int k = 0; k < NT; ++k) {
        const bool last = k + 1 == NT;
        need(last ? k : k + 1);
        const int ss = k % C::NS;
        const size_t cs = (size_t)((cb + k) % ring), cn = (size_t)((cb + k + 1) % ring);
        if (k == 1) mbar_wait(&bars.w0done, 0);
        if (k >= C::NS) {
          mbar_wait(&bars.emptyS[ss], (k / C::NS - 1) & 1);
          // every product that read step k - NS has retired, so every copy out of chunk k - NS's images completed
          // long ago: that image slot may be overwritten (ring hand-off with the pre-pass)
          if (lane == 0 && ring < NTROW)
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(progress), "r"((uint32_t)(k - C::NS + 1)) : "memory");
        }
        uint8_t* sl = smem + C::OFF_S + ss * C::SSLOT;
        mbar_arrive_expect_tx_ws(&bars.fullS[ss], (last ? 0u : 16384u + AU_BYTES) + 16384u + P_BYTES + KT_BYTES + TAIL_BYTES);
        if (!last) {
          for (int pn = 0; pn < 2; ++pn)
            bulk_g2s_ws(sl + C::S_XO + pn * C::XO_PANEL, blob + cn * BLOB_BYTES + BLOB_OFF_A1 + pn * C::XO_PANEL, 8192,
                        &bars.fullS[ss]);
          bulk_g2s_ws(sl + C::S_RC, aublob + cn * (GDN_NS * UBLOB_BYTES) + AU_BYTES, AU_BYTES, &bars.fullS[ss]);
        }
        for (int pn = 0; pn < 2; ++pn)
          bulk_g2s_ws(sl + C::S_XO + pn * C::XO_PANEL + 8192, blob + cs * BLOB_BYTES + BLOB_OFF_A1 + pn * C::XO_PANEL + 8192,
                      8192, &bars.fullS[ss]);
        bulk_g2s_ws(sl + C::S_RC + AU_BYTES, blob + cs * BLOB_BYTES + BLOB_OFF_P, P_BYTES, &bars.fullS[ss]);
        bulk_g2s_ws(sl + C::S_KT, blob + cs * BLOB_BYTES + BLOB_OFF_KT, KT_BYTES + TAIL_BYTES, &bars.fullS[ss]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer that follows v_new ---------------------
    const uint32_t sbase = smem_u32(smem);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    for (int k = 0; k < NT; ++k) {
      const int ss = k % C::NS, p = k & 1;
      [[maybe_unused]] const int c = k;   // (timeline probe)
      const bool last = k + 1 == NT;
      const uint32_t sl = sbase + C::OFF_S + ss * C::SSLOT;
      const uint32_t vbk = tm + C::TM_VB + p * 32;
      const uint32_t pair = tm + C::TM_PAIR + p * 128;
      mbar_wait(&bars.vb, k & 1);                                    // bf16 v_new_k^T is in tensor memory
      mbar_wait(&bars.fullS[ss], (k / C::NS) & 1);
      TTR(0);
      const uint64_t dKt = umma_desc(sl + C::S_KT, C::KT_PANEL, 1024, SWZ_128B);   // MN-major: 16 tokens = 2 KiB per MMA
      const uint64_t dRC = umma_desc(sl + C::S_RC, 16, 1024, SWZ_128B);            // K-major: 32 B per MMA
      auto issue_b = [&]() {
        if (k >= 1) mbar_wait(&bars.dsfree, (k - 1) & 1);            // the state warps have read DS of step k - 1
        tc_fence_after();
        TTR(14);
#pragma unroll
        for (int j = 0; j < 4; ++j) umma_bf16_ts_ws(tm + C::TM_DS, vbk + j * 8, dKt + j * 128, idescB, j > 0);
        umma_commit_ws(&bars.ds);
        TTR(2);
      };
      auto issue_rc = [&]() {
        tc_fence_after();
        if (!last) {
#pragma unroll
          for (int j = 0; j < 4; ++j) umma_bf16_ts_ws(pair, vbk + j * 8, dRC + j * 2, idesc128, 1);
        } else {   // no chunk k + 1: only the P half, into the O half of the pair
#pragma unroll
          for (int j = 0; j < 4; ++j) umma_bf16_ts_ws(pair + 64, vbk + j * 8, dRC + (AU_BYTES >> 4) + j * 2, idesc64, 1);
        }
        umma_commit_ws(&bars.rc[p]);
      };
      // the RC part is the serial chain, the B part feeds the state: whichever is not blocked goes first (the two
      // write different accumulators, so the order does not change a bit of the result)
      if (__any_sync(0xffffffffu, mbar_test_wait(&bars.xo[p], (k >> 1) & 1))) {
        issue_rc();
        issue_b();
      } else {
        issue_b();
        mbar_wait(&bars.xo[p], (k >> 1) & 1);
        issue_rc();
      }
      umma_commit_ws(&bars.emptyS[ss]);
      TTR(1);
    }
  } else if (warp == 2) {
    // ------------------------------- MMA issuer that follows the state -----------------
    const uint32_t sbase = smem_u32(smem);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    auto issue_u = [&](int c) {   // DV half of PAIR[(c + 1) & 1] = V_c^T Au_c^T
      const int su = c % C::NU;
      const uint32_t us = sbase + C::OFF_U + su * C::USLOT;
      mbar_wait(&bars.fullU[su], (c / C::NU) & 1);
      if (c >= 2) mbar_wait(&bars.dvfree[(c + 1) & 1], ((c - 2) >> 1) & 1);   // v_new accumulator of chunk c - 2 has been read
      tc_fence_after();
      const uint64_t dV = umma_desc(us + C::U_V, C::V_PANEL, 1024, SWZ_128B);
      const uint64_t dAu = umma_desc(us + C::U_AU, 16, 1024, SWZ_128B);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        umma_bf16_ws(tm + C::TM_PAIR + ((c + 1) & 1) * 128, dV + j * 128, dAu + j * 2, idescU, j > 0);
      umma_commit_ws(&bars.emptyU[su]);
    };
    issue_u(0);
    if (NT > 1) issue_u(1);
    mbar_wait(&bars.sb, 0);
    mbar_wait(&bars.w0full, 0);
    tc_fence_after();
    {   // v_new_0^T = U_0^T - bf16(S_0^T) Wg_0^T
      const uint64_t dW0 = umma_desc(sbase + C::OFF_S + C::SSLOT + C::S_XO, 16, 1024, SWZ_128B);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        umma_bf16_ts_ws(tm + C::TM_PAIR + 128, tm + C::TM_SB + j * 8, dW0 + (j >> 2) * (C::XO_PANEL >> 4) + (j & 3) * 2, idesc64, 1);
      umma_commit_ws(&bars.dv0);
      umma_commit_ws(&bars.w0done);
    }
    for (int k = 0; k < NT; ++k) {
      const int ss = k % C::NS, p = k & 1;
      [[maybe_unused]] const int c = k;   // (timeline probe)
      const bool last = k + 1 == NT;
      if (k >= 1) mbar_wait(&bars.sb, k & 1);      // bf16 S_k^T is in tensor memory
      mbar_wait(&bars.fullS[ss], (k / C::NS) & 1);
      if (k >= 2) mbar_wait(&bars.dofree[p], ((k >> 1) - 1) & 1);   // O half of the pair has been read and zeroed
      tc_fence_after();
      TTR(3);
      const uint64_t dXO = umma_desc(sbase + C::OFF_S + ss * C::SSLOT + C::S_XO, 16, 1024, SWZ_128B);
      if (!last) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_bf16_ts_ws(tm + C::TM_PAIR + p * 128, tm + C::TM_SB + j * 8, dXO + (j >> 2) * (C::XO_PANEL >> 4) + (j & 3) * 2,
                          idesc128, 1);
      } else {   // no chunk k + 1: only the Qg rows (64..127 of each panel), into the O half of the pair
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_bf16_ts_ws(tm + C::TM_PAIR + p * 128 + 64, tm + C::TM_SB + j * 8,
                          dXO + (8192 >> 4) + (j >> 2) * (C::XO_PANEL >> 4) + (j & 3) * 2, idesc64, 1);
      }
      umma_commit_ws(&bars.xo[p]);
      umma_commit_ws(&bars.emptyS[ss]);
      TTR(4);
      if (k + 2 < NT) issue_u(k + 2);
      TTR(5);
    }
  } else if (warp < 7) {
    // ------------------------------- v_new / output epilogue (warps 3..6) ----------------
    const int quad = warp & 3;
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const int col = col0 + quad * 32 + lane;
    uint32_t r[32], r2[32], w[32];
    auto output = [&](int c) {   // O_c^T from the O half of PAIR[c & 1] (complete with the RC part of step c)
      const uint32_t dox = tlane + C::TM_PAIR + (c & 1) * 128 + 64;
      tmem_ld32(dox, r);
      tmem_ld32(dox + 32, r2);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) w[i] = 0u;
      tmem_st32(dox, w);
      tmem_st32(dox + 32, w);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.dofree[c & 1]);
      const int tok0 = varlen ? __ldg(vl.chunk_tok0 + cb + c) : c * GDN_C;
      const int valid = varlen ? __ldg(vl.chunk_valid + cb + c) : min(GDN_C, T - tok0);
      const size_t tstride = (size_t)H * GDN_V;
      __nv_bfloat16* p0 = o + (((size_t)b * T + tok0) * H + h) * GDN_V + col;
      __nv_bfloat16* p1 = p0 + 32 * tstride;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i < valid) *p0 = __float2bfloat16(__uint_as_float(r[i]));
        if (i + 32 < valid) *p1 = __float2bfloat16(__uint_as_float(r2[i]));
        asm volatile("" : "+l"(p0), "+l"(p1));
        p0 += tstride;
        p1 += tstride;
      }
    };
    for (int c = 0; c < NT; ++c) {
      // v_new_c^T sits in the DV half of PAIR[(c + 1) & 1]: complete with the RC part of step c - 1 (chunk 0: prologue)
      if (c == 0) mbar_wait(&bars.dv0, 0);
      else mbar_wait(&bars.rc[(c - 1) & 1], ((c - 1) >> 1) & 1);
      tc_fence_after();
      if (quad == 0) TTR(6);
      const uint32_t dvx = tlane + C::TM_PAIR + ((c + 1) & 1) * 128;
      tmem_ld32(dvx, r);
      tmem_ld32(dvx + 32, r2);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        w[i] = pack_bf16(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
        w[16 + i] = pack_bf16(__uint_as_float(r2[2 * i]), __uint_as_float(r2[2 * i + 1]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.dvfree[(c + 1) & 1]);
      tmem_st32(tlane + C::TM_VB + (c & 1) * 32, w);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.vb);
      if (quad == 0) TTR(7);
      if (c > 0) output(c - 1);
    }
    mbar_wait(&bars.rc[(NT - 1) & 1], ((NT - 1) >> 1) & 1);
    tc_fence_after();
    output(NT - 1);
  } else {
    // ------------------------------- state warps (7..14) --------------------------------
    const int quad = warp & 3, half = (warp - 7) >> 2;
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const int col = col0 + quad * 32 + lane;
    float S[64];
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
    auto publish = [&]() {
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
      const int ss = c % C::NS;
      if (warp == 8) TTR(12);
      mbar_wait(&bars.fullS[ss], (c / C::NS) & 1);
      if (warp == 8) TTR(13);
      const float gamma = *reinterpret_cast<const float*>(smem + C::OFF_S + ss * C::SSLOT + C::S_TAIL);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.emptyS[ss]);
      mbar_wait(&bars.ds, c & 1);
      tc_fence_after();
      if (warp == 8) TTR(8);
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        tmem_ld32(tlane + C::TM_DS + half * 64 + p * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) S[p * 32 + i] = fmaf(gamma, S[p * 32 + i], __uint_as_float(r[i]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.dsfree);   // DS may be overwritten by the next B part
      if (warp == 8) TTR(9);
      if (c + 1 < NT) {
        // the XO part of step c (the only reader of bf16 S_c^T; chunk 0's extra product was committed before it by
        // the same warp) has retired before the operand is overwritten
        mbar_wait(&bars.xo[c & 1], (c >> 1) & 1);
        if (warp == 8) TTR(10);
        publish();
        if (warp == 8) TTR(11);
      }
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

// =====================================================================================================
// Pipelined form of the transposed scan (the default): same operands, same arithmetic per product as
// gdn_scan_t_kernel, but the two tensor <-> epilogue hand-offs of the serial chain are PIPELINED over the contraction
// dimension, and the output products leave the chain.
//
// Timeline of gdn_scan_t_kernel (clock64 probe, cycles per chunk ~1900): W part done +466, epi V +310, B part +456
// (it queues behind the O part on the in-order tensor pipe), epi S +736 (TMEM -> registers moves 64 B/cycle per
// quadrant: 256 cycles for the 128 fp32 columns of DS alone, then pack / store / fence / arrive).  Here
//   * sixteen state warps (four per TMEM lane quadrant) own 32 key dims each and publish their quarter of bf16 S^T on
//     their own barrier; the MMA warp issues the two K = 16 slabs of the W part that read that quarter as soon as
//     it arrives, so the W part overlaps the state epilogue instead of following it;
//   * the v_new warps publish bf16 v_new^T in two halves of 32 tokens, the B part follows half by half;
//   * bf16 S^T is double-buffered in tensor memory (the v_new accumulator is single-buffered instead: its U part is
//     only issued after v_new has been read anyway), so the O part may read S_c while S_{c+1} is being published: the
//     O and C parts are issued AFTER the B part and run on the tensor pipe while the state epilogue works.
// Serial chain per chunk: state quarter 3 -> last W slabs -> epi V half 1 -> last B slabs.  Three A slots: the slot
// of chunk c is released after the O part, late in the step.
// =====================================================================================================
struct T3Cfg {
  static constexpr int THREADS = 704;   // 22 warps: copy, MMA, 4 v_new/output, 16 state
  static constexpr int NA = 3, NK = 2;
  static constexpr uint32_t A_BW = 0;                         // [-Wg ; Qg]  32 KiB, K-major, no swizzle
  static constexpr uint32_t A_V = A1_BYTES;                   // value tile: 2 panels [64 tok][64 val], 128B swizzle
  static constexpr uint32_t V_PANEL = 64 * 64 * 2;
  static constexpr uint32_t A_AU = A1_BYTES + 2 * V_PANEL;    // Au 8 KiB, K-major, no swizzle
  static constexpr uint32_t ASLOT = A_AU + AU_BYTES;          // 56 KiB
  static constexpr uint32_t A_TX = ASLOT;
  static constexpr uint32_t K_P = 0;
  static constexpr uint32_t K_KT = P_BYTES;
  static constexpr uint32_t K_TAIL = P_BYTES + KT_BYTES;
  static constexpr uint32_t K_TX = P_BYTES + KT_BYTES + TAIL_BYTES;
  static constexpr uint32_t KSLOT = P_BYTES + KT_BYTES + 1024;   // 25 KiB
  static constexpr uint32_t OFF_A = 0;
  static constexpr uint32_t OFF_K = NA * ASLOT;
  static constexpr uint32_t OFF_BARS = OFF_K + NK * KSLOT;
  static constexpr uint32_t SMEM = OFF_BARS + 512 + 1024;
  static constexpr uint32_t TM_DS = 0;       // 128: v_new^T Kt (state increment)
  static constexpr uint32_t TM_SB = 128;     // 2 x 64: bf16 S^T of even / odd chunks (A operand, K = 128)
  static constexpr uint32_t TM_DV = 256;     //  64: v_new^T accumulator
  static constexpr uint32_t TM_VB = 320;     //  32: bf16 v_new^T (A operand, K = 64)
  static constexpr uint32_t TM_DO = 384;     // 2 x 64: O^T accumulators
  static constexpr uint32_t TM_COLS = 512;
  static_assert(ASLOT % 1024 == 0 && KSLOT % 1024 == 0 && A_V % 1024 == 0, "swizzled tiles need 1 KiB alignment");
  static_assert(SMEM <= 232448, "exceeds 227 KiB");
};

struct T3Bars {
  uint64_t fullA[3], emptyA[3], fullK[2], emptyK[2];
  uint64_t sb[4], vb[2], ds, dv, dofull[2], dofree[2];
  uint32_t tmem_base;
};

// 704 threads: 65536 / 704 -> 88 registers per thread; a state thread keeps 32 fp32 state entries for the whole
// sequence plus one 32-column staging buffer, the v_new / output threads work in halves of 32 columns
__global__ void __launch_bounds__(T3Cfg::THREADS, 1)
gdn_scan_t3_kernel(const __grid_constant__ CUtensorMap tmV, GdnWorkspace ws, GdnVarlen vl, const void* __restrict__ h0,
                   int h0_dtype, __nv_bfloat16* __restrict__ o, void* __restrict__ ht, int ht_dtype, int T, int H,
                   int NTROW) {
  using C = T3Cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  T3Bars& bars = *reinterpret_cast<T3Bars*>(smem + C::OFF_BARS);
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
  // resident: with an image ring the pre-pass is only launched once every scan CTA has checked in (ivl_abi.cu)
  if (tid == 0) atomicAdd(ws.checkin, 1u);
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
    for (int s = 0; s < C::NA; ++s) { mbar_init(&bars.fullA[s], 1); mbar_init(&bars.emptyA[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.fullK[s], 1);
      mbar_init(&bars.emptyK[s], 1 + 16);   // the C part has retired + the sixteen state warps have read gamma
      mbar_init(&bars.dofull[s], 1); mbar_init(&bars.dofree[s], 4);
      mbar_init(&bars.vb[s], 4);
    }
    for (int s = 0; s < 4; ++s) mbar_init(&bars.sb[s], 4);
    mbar_init(&bars.ds, 1); mbar_init(&bars.dv, 1);
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
      if (c >= C::NA) mbar_wait_relaxed(&bars.emptyA[sa], (c / C::NA - 1) & 1);
      uint8_t* as = smem + C::OFF_A + sa * C::ASLOT;
      mbar_arrive_expect_tx_ws(&bars.fullA[sa], C::A_TX);
      bulk_g2s_ws(as + C::A_BW, blob + cs * BLOB_BYTES + BLOB_OFF_A1, A1_BYTES, &bars.fullA[sa]);
      bulk_g2s_ws(as + C::A_AU, aublob + cs * (GDN_NS * UBLOB_BYTES), AU_BYTES, &bars.fullA[sa]);
      tma_load_4d_ws(as + C::A_V, &tmV, col0, h, tok0, b, &bars.fullA[sa]);
      tma_load_4d_ws(as + C::A_V + C::V_PANEL, &tmV, col0 + 64, h, tok0, b, &bars.fullA[sa]);
      if (c >= C::NK) {
        mbar_wait_relaxed(&bars.emptyK[sk], (c / C::NK - 1) & 1);
        // every product that read the K slot of chunk c - NK has retired, and -- the O part of a chunk is issued
        // before its C part -- so has every product that read its A slot: the image slot may be overwritten
        if (lane == 0 && ring < NTROW)
          asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(progress), "r"((uint32_t)(c - C::NK + 1)) : "memory");
      }
      uint8_t* ks = smem + C::OFF_K + sk * C::KSLOT;
      mbar_arrive_expect_tx_ws(&bars.fullK[sk], C::K_TX);
      bulk_g2s_ws(ks + C::K_P, blob + cs * BLOB_BYTES + BLOB_OFF_P, C::K_TX, &bars.fullK[sk]);
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------------
    constexpr uint32_t idescU = umma_idesc_bf16(128, 64, /*a_mn=*/1, /*b_mn=*/0);
    constexpr uint32_t idesc64 = umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idescB = umma_idesc_bf16(128, 128, 0, /*b_mn=*/1);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    auto issue_u = [&](int c) {   // DV = V^T Au^T
      const uint32_t as = sbase + C::OFF_A + (c % C::NA) * C::ASLOT;
      const uint64_t dV = umma_desc(as + C::A_V, C::V_PANEL, 1024, SWZ_128B);
      const uint64_t dAu = umma_desc(as + C::A_AU, 128, 1024, SWZ_NONE);
#pragma unroll
      for (int j = 0; j < 4; ++j) umma_bf16_ws(tm + C::TM_DV, dV + j * 128, dAu + j * 16, idescU, j > 0);
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
      const uint32_t sbt = tm + C::TM_SB + buf * 64, dob = tm + C::TM_DO + buf * 64;
      // W part, key quarter by key quarter as the state warps publish bf16 S_c^T
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        mbar_wait(&bars.sb[q], c & 1);
        tc_fence_after();
        if (q == 0) TTR(0);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int j = 2 * q + jj;
          umma_bf16_ts_ws(tm + C::TM_DV, sbt + j * 8, dW + j * 16, idesc64, 1);
        }
      }
      umma_commit_ws(&bars.dv);
      TTR(1);
      // B part, token half by token half as the v_new warps publish bf16 v_new^T
      mbar_wait(&bars.fullK[sk], (c / C::NK) & 1);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        mbar_wait(&bars.vb[hh], c & 1);
        tc_fence_after();
        if (hh == 0) TTR(2);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int j = 2 * hh + jj;
          umma_bf16_ts_ws(tm + C::TM_DS, tm + C::TM_VB + j * 8, dKt + j * 16, idescB, j > 0);
        }
      }
      umma_commit_ws(&bars.ds);
      TTR(3);
      // O and C parts: off the serial chain, on the tensor pipe while the state epilogue works
      if (c >= 2) {
        mbar_wait(&bars.dofree[buf], ((c >> 1) - 1) & 1);   // O accumulator of chunk c - 2 has been read
        tc_fence_after();
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) umma_bf16_ts_ws(dob, sbt + j * 8, dQ + j * 16, idesc64, j > 0);
      umma_commit_ws(&bars.emptyA[sa]);
#pragma unroll
      for (int j = 0; j < 4; ++j) umma_bf16_ts_ws(dob, tm + C::TM_VB + j * 8, dP + j * 16, idesc64, 1);
      umma_commit_ws(&bars.dofull[buf]);
      umma_commit_ws(&bars.emptyK[sk]);
      TTR(4);
      if (c + 1 < NT) {
        // the v_new accumulator is free: both halves of v_new_c were read before they were published
        mbar_wait(&bars.fullA[(c + 1) % C::NA], ((c + 1) / C::NA) & 1);
        tc_fence_after();
        issue_u(c + 1);
      }
      TTR(5);
    }
  } else if (warp < 6) {
    // ------------------------------- v_new / output epilogue --------------------------
    const int quad = warp & 3;
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const int col = col0 + quad * 32 + lane;     // this thread's value column
    uint32_t r[32], w[16];
    auto output = [&](int c) {
      const int buf = c & 1;
      mbar_wait(&bars.dofull[buf], (c >> 1) & 1);
      tc_fence_after();
      const int tok0 = varlen ? __ldg(vl.chunk_tok0 + cb + c) : c * GDN_C;
      const int valid = varlen ? __ldg(vl.chunk_valid + cb + c) : min(GDN_C, T - tok0);
      const size_t tstride = (size_t)H * GDN_V;
      __nv_bfloat16* p0 = o + (((size_t)b * T + tok0) * H + h) * GDN_V + col;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        tmem_ld32(tlane + C::TM_DO + buf * 64 + hh * 32, r);
        tmem_ld_wait();
        if (hh == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.dofree[buf]);
        }
        // one 64-byte row segment per warp and token (the pointer advances by one token row)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (hh * 32 + i < valid) *p0 = __float2bfloat16(__uint_as_float(r[i]));
          asm volatile("" : "+l"(p0));   // keep the increments serial
          p0 += tstride;
        }
      }
    };
    for (int c = 0; c < NT; ++c) {
      mbar_wait(&bars.dv, c & 1);
      tc_fence_after();
      if (quad == 0) TTR(6);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        tmem_ld32(tlane + C::TM_DV + hh * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = pack_bf16(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
        tmem_st16(tlane + C::TM_VB + hh * 16, w);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.vb[hh]);
        if (quad == 0) TTR(7 + hh);
      }
      if (c > 0) output(c - 1);
    }
    output(NT - 1);
  } else {
    // ------------------------------- state warps ---------------------------------------
    // four warps per TMEM lane quadrant: warp (quad, kq) owns key dims 32 * kq .. + 31 of its 32 value columns
    const int quad = warp & 3, kq = (warp - 6) >> 2;
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const int col = col0 + quad * 32 + lane;
    float S[32];                                 // a quarter row of S^T (this value column, 32 key dims), fp32
    const size_t sbase_off = (((size_t)seq * H + h) * GDN_K + kq * 32) * GDN_V + col;
    if (h0 == nullptr) {
#pragma unroll
      for (int i = 0; i < 32; ++i) S[i] = 0.f;
    } else if (h0_dtype == 0) {
      const float* p = static_cast<const float*>(h0) + sbase_off;
#pragma unroll
      for (int i = 0; i < 32; ++i) S[i] = __ldg(p + (size_t)i * GDN_V);
    } else {
      const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(h0) + sbase_off;
#pragma unroll
      for (int i = 0; i < 32; ++i) S[i] = __bfloat162float(p[(size_t)i * GDN_V]);
    }
    uint32_t r[32];
    auto publish = [&](int n) {   // bf16 S_n^T -> TMEM A operand of chunk n (word i = key dims 2i, 2i+1)
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = pack_bf16(S[2 * i], S[2 * i + 1]);
      tmem_st16(tlane + C::TM_SB + (n & 1) * 64 + kq * 16, r);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.sb[kq]);
    };
    publish(0);
    for (int c = 0; c < NT; ++c) {
      const int sk = c % C::NK;
      mbar_wait(&bars.fullK[sk], (c / C::NK) & 1);
      const float gamma = *reinterpret_cast<const float*>(smem + C::OFF_K + sk * C::KSLOT + C::K_TAIL);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.emptyK[sk]);
      mbar_wait(&bars.ds, c & 1);
      tc_fence_after();
      if (warp == 8) TTR(9);
      tmem_ld32(tlane + C::TM_DS + kq * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) S[i] = fmaf(gamma, S[i], __uint_as_float(r[i]));
      if (warp == 8) TTR(10);
      if (c + 1 < NT) {
        // bf16 S_{c+1}^T goes into the other operand buffer: its last readers (W and O parts of chunk c - 1) were
        // issued before the B part of chunk c, whose commit this warp has just seen
        publish(c + 1);
      } else {
        tc_fence_before();
      }
      if (warp == 8) TTR(11);
    }
    if (ht != nullptr) {
      if (ht_dtype == 0) {
        float* p = static_cast<float*>(ht) + sbase_off;
#pragma unroll
        for (int i = 0; i < 32; ++i) p[(size_t)i * GDN_V] = S[i];
      } else {
        __nv_bfloat16* p = static_cast<__nv_bfloat16*>(ht) + sbase_off;
#pragma unroll
        for (int i = 0; i < 32; ++i) p[(size_t)i * GDN_V] = __float2bfloat16(S[i]);
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
                              const void* h0, int h0_dtype, void* o, void* ht, int ht_dtype, int T, int H, int form,
                              cudaStream_t stream) {
  using C = TCfg;
  static std::atomic<bool> configured[64];
  int dev = 0;
  if (cudaError_t e = cudaGetDevice(&dev)) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!configured[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(gdn_scan_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gdn_scan_t2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T2Cfg::SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gdn_scan_t3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T3Cfg::SMEM);
    if (e != cudaSuccess) return e;
    configured[dev].store(true, std::memory_order_release);
  }
  // The value tensor's map is a pure function of (pointer, B, T, H): the last few are kept, so a layer that is called
  // again with the same buffer (every step of a served model) does not re-encode -- and `ncu --replay-mode range`,
  // which refuses cuTensorMapEncodeTiled inside a range, can profile a warmed-up call
  struct MapKey { const void* v; int B, T, H; CUtensorMap tm; };
  static std::mutex map_mu;
  static MapKey map_cache[8];
  static unsigned map_next = 0;
  CUtensorMap tm;
  bool hit = false;
  {
    std::lock_guard<std::mutex> lock(map_mu);
    for (const MapKey& e : map_cache)
      if (e.v == v && e.B == B && e.T == T && e.H == H) { tm = e.tm; hit = true; break; }
  }
  if (!hit) {
    EncodeTiledFnT enc = encode_fn_t();
    if (!enc) return cudaErrorNotSupported;
    cuuint64_t dims[4] = {(cuuint64_t)GDN_V, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)GDN_V * 2, (cuuint64_t)H * GDN_V * 2, (cuuint64_t)T * H * GDN_V * 2};
    cuuint32_t box[4] = {64, 1, 64, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
    std::lock_guard<std::mutex> lock(map_mu);
    MapKey& e = map_cache[map_next++ % 8];
    e.v = v; e.B = B; e.T = T; e.H = H; e.tm = tm;
  }
  dim3 grid(2, H, nseq);
  if (form == 3)
    gdn_scan_t3_kernel<<<grid, T3Cfg::THREADS, T3Cfg::SMEM, stream>>>(tm, ws, vl, h0, h0_dtype,
                                                                      static_cast<__nv_bfloat16*>(o), ht, ht_dtype, T, H,
                                                                      ntrow);
  else if (form == 2)
    gdn_scan_t2_kernel<<<grid, T2Cfg::THREADS, T2Cfg::SMEM, stream>>>(tm, ws, vl, h0, h0_dtype,
                                                                      static_cast<__nv_bfloat16*>(o), ht, ht_dtype, T, H,
                                                                      ntrow);
  else
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
