// Gated DeltaNet inter-chunk scan: the serial part of the chunked delta rule, one
// persistent CTA per (batch, head, BV-column slice of V), the recurrent state slice
// S[128 x BV] kept on-chip (fp32 master in TMEM, bf16 shadow in shared memory) for the
// whole sequence.
//
// Replaces chunk_gated_delta_rule_fwd_h + chunk_fwd_o of the reference
// (src/llamafactory/model/fla/ops/common/chunk_delta_h.py:32-124,247-318 and
//  ops/common/chunk_o.py:32-114,456-497) without ever materialising the per-chunk
// states h[B,NT,H,K,V] (2.15 GB per layer at 128K tokens in the reference).
//
// Per chunk c and 32-column chain of the CTA (operands are the images written by gdn_prep.cu,
// loaded with 1-D bulk TMA copies through two mbarrier rings shared by the chains):
//   MMA-A  D1 = [-Wg ; Qg] . bf16(S)            M128 N32 K128   (tcgen05, accum in TMEM)
//   epi    Vn = U + D1[0:64]      -> bf16 -> shared (MN-major B operand)
//   MMA-B  S  = gamma S + Kt^T . Vn             M128 N32 K64    (gamma pre-applied in TMEM)
//   MMA-C  D1[64:128] += P . Vn                 M128 N32 K64    (rows 0..63 of the A operand are zero)
//   epi    O = D1[64:128] -> bf16 -> global;  S -> bf16 shadow, gamma_{c+1} S -> TMEM
// Warp roles: warp 0 copies (and follows the prep kernel's per-chunk ready flags), warps 1..BV/32 issue the
// MMAs of one chain each, then four epilogue warps per chain (TMEM lane quadrant = warp % 4).
// Packed variable-length batches (GdnVarlen tables) run one CTA group per sequence over that sequence's chunks.
#include <atomic>

#include "gdn_layout.cuh"
#include "sm100.cuh"

namespace ivl {

namespace {

// Per-variant geometry.  BV = value columns owned by one CTA (32, 64 or 128): the CTA count is
// B * H * 256 / BV.  The recurrence is independent per value column, and its step time is pure latency
// (measured: the same 790 ns per chunk for 8 or 128 CTAs), so a CTA runs BV / 32 independent 32-column
// CHAINS side by side -- each with its own MMA warp, four epilogue warps, barriers, TMEM columns and
// shadow / v_new buffers -- that share one copy of the operand rings.  While one chain waits on its
// MMA -> epilogue -> MMA round trip the others use the tensor pipe and the shared-memory ports.  (One wide
// N = BV chain was measured at 1079 / 1742 ns per chunk for BV = 64 / 128 against 767 ns for BV = 32:
// every dependent N = 128 MMA costs ~210 cycles and the shadow store alone 256.)  BV = 64 leaves
// 84 SMs free at B = 1, H = 16, which is what lets gdn_prep_kernel run concurrently (ivl_gdn_chunk_fwd).
//
// Shared memory: two operand rings that are recycled at different points of a step --
//   A ring  (NA slots of 32 KiB): [-Wg ; Qg], free as soon as MMA-A has retired (early in the step)
//   K ring  (NK slots): [8 KiB of zeros | P | Kt | gamma | U slice], free after MMA-C (end of the step)
// -- so a two-slot K ring still gives the A operand more than a full step of prefetch distance.
template <int BV>
struct ScanCfg {
  static_assert(BV == 32 || BV == 64 || BV == 128, "slice width");
  static constexpr int NCG = BV / 32;                 // independent 32-column chains sharing the operand rings
  static constexpr int EPI_WARPS = 4 * NCG;           // per chain: one epilogue warp per TMEM lane quadrant
  static constexpr int THREADS = 32 + 32 * NCG + 32 * EPI_WARPS;  // warp 0 copies, warps 1..NCG issue MMAs
  static constexpr int NA = (BV == 32) ? 3 : 2;
  static constexpr int NK = (BV == 128) ? 2 : 3;
  static constexpr uint32_t U_BYTES = 64 * BV * 2;
  static constexpr uint32_t KS_OFF_Z = 0;                      // rows 0..63 of the [0 ; P] operand, zeroed once
  static constexpr uint32_t KS_OFF_P = P_BYTES;                // P | Kt land here with one copy
  static constexpr uint32_t KS_OFF_KT = 2 * P_BYTES;
  static constexpr uint32_t KS_OFF_TAIL = 2 * P_BYTES + KT_BYTES;   // gamma of the chunk (first 4 bytes)
  static constexpr uint32_t KS_OFF_U = KS_OFF_TAIL + TAIL_BYTES;
  static constexpr uint32_t KSLOT = KS_OFF_U + U_BYTES + (1024 - TAIL_BYTES);
  static constexpr uint32_t OFF_A = 0;
  static constexpr uint32_t OFF_K = NA * A1_BYTES;
  static constexpr uint32_t SB_BYTES = 128 * BV * 2;           // bf16 shadow of S, MN-major B operand
  static constexpr uint32_t VN_BYTES = 64 * BV * 2;            // v_new, MN-major B operand
  static constexpr uint32_t OFF_SB = OFF_K + NK * KSLOT;
  static constexpr uint32_t OFF_VN = OFF_SB + SB_BYTES;
  static constexpr uint32_t OFF_BARS = OFF_VN + VN_BYTES;
  static constexpr uint32_t SMEM = OFF_BARS + 512 + 1024;      // + alignment slack
  static constexpr uint32_t TM_CHAIN = 96;                     // TMEM columns per chain: D1[0] | D1[1] | S
  static constexpr uint32_t TM_D1 = 0;
  static constexpr uint32_t TM_S = 64;
  static constexpr uint32_t TM_COLS = 4 * BV;                  // power of two >= 96 NCG
  static_assert(KSLOT % 1024 == 0, "slot alignment");
  static_assert(SMEM <= 232448, "exceeds 227 KiB");
};

// Developer-only timeline probe (compiled in with -DIVL_TRACE by tools/trace_scan.py; never in the product build)
#ifdef IVL_TRACE
__device__ long long ivl_trace_buf[64 * 16];
__device__ unsigned long long ivl_scan_wait[4];  // copy-warp cycles spent waiting for ready flags, polls, spins
__device__ unsigned long long ivl_scan_tl[16 * 2048 * 4];  // head 0: globaltimer when the copy warp of slice s issued chunk c
#define TR(slot)                                                                                         \
  do {                                                                                                   \
    if (tr_on && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && c >= 1000 && c < 1064)               \
      ivl_trace_buf[(c - 1000) * 16 + (slot)] = clock64();                                               \
  } while (0)
#else
#define TR(slot) do { } while (0)
#endif

struct Bars {
  uint64_t fullA[3], emptyA[3], fullK[3], emptyK[3];
  uint64_t a[4][2], o[4][2];   // per chain
  uint64_t vnst[4], s[4], sb[4];
  uint32_t tmem_base;
};

__device__ __forceinline__ void store_row_bf16(uint8_t* base, uint32_t piece_stride, const float* x) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    uint4 w;
    w.x = pack_bf16(x[p * 8 + 0], x[p * 8 + 1]);
    w.y = pack_bf16(x[p * 8 + 2], x[p * 8 + 3]);
    w.z = pack_bf16(x[p * 8 + 4], x[p * 8 + 5]);
    w.w = pack_bf16(x[p * 8 + 6], x[p * 8 + 7]);
    *reinterpret_cast<uint4*>(base + p * piece_stride) = w;
  }
}

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int BV>
__global__ void __launch_bounds__(ScanCfg<BV>::THREADS, 1)
gdn_scan_kernel(GdnWorkspace ws, GdnVarlen vl, const void* __restrict__ h0, int h0_dtype,
                __nv_bfloat16* __restrict__ o, void* __restrict__ ht, int ht_dtype, int T, int H, int NTROW) {
  using C = ScanCfg<BV>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars& bars = *reinterpret_cast<Bars*>(smem + C::OFF_BARS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.x, h = blockIdx.y;
  // Dense: blockIdx.z = batch row b, which owns chunks 0 .. NTROW-1 of workspace row (b, h).
  // Packed (vl): blockIdx.z = sequence n of the single flattened row; its chunks are cb .. cb+NT-1 of row (0, h).
  const bool varlen = vl.chunk_tok0 != nullptr;
  const int b = varlen ? 0 : blockIdx.z;
  const int seq = blockIdx.z;                           // index of the initial / final state
  const int cb = varlen ? __ldg(vl.seq_chunk_begin + seq) : 0;
  const int NT = varlen ? __ldg(vl.seq_chunk_begin + seq + 1) - cb : NTROW;   // chunks this CTA scans
  const size_t ch0 = ((size_t)b * H + h) * NTROW + cb;  // ready flag of this CTA's first chunk
  const size_t slot0 = ((size_t)b * H + h) * ws.ring;   // first image slot of the row; chunk c lives in c % ring
  const int ring = ws.ring;
  if (NT <= 0) {
    // empty sequence: the final state is the initial state
    if (ht != nullptr) {
      const size_t base = ((size_t)seq * H + h) * GDN_K * GDN_V;
      for (int i = tid; i < GDN_K * BV; i += C::THREADS) {
        const size_t off = base + (size_t)(i / BV) * GDN_V + slice * BV + (i % BV);
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
  const uint8_t* ublob = ws.ublob + (slot0 * GDN_NS + slice * C::NCG) * UBLOB_BYTES;

  if (tid == 0) {
    for (int s = 0; s < 3; ++s) {
      // a slot is recycled when every chain's MMA warp has released it
      mbar_init(&bars.fullA[s], 1); mbar_init(&bars.emptyA[s], C::NCG);
      mbar_init(&bars.fullK[s], 1); mbar_init(&bars.emptyK[s], C::NCG);
    }
    // epilogue warps arrive once per warp (lane 0 after __syncwarp)
    for (int ch = 0; ch < C::NCG; ++ch) {
      for (int i = 0; i < 2; ++i) { mbar_init(&bars.a[ch][i], 1); mbar_init(&bars.o[ch][i], 1); }
      mbar_init(&bars.s[ch], 1);
      // sb   (4 warps): bf16 shadow of S_c in shared memory + operands of chunk c landed -> MMA-A(c) may be
      //                 issued.  The O warps arrive here only after they finished reading D1 of chunk c-2,
      //                 so the accumulator buffer MMA-A(c) overwrites is free without a separate barrier.
      // vnst (2 + 4):   v_new(c) in shared memory (the warps of TMEM quadrants 0 and 1) and gamma_c S_c back
      //                 in TMEM (4 warps) -> MMA-B(c) / MMA-C(c) may be issued.
      mbar_init(&bars.sb[ch], 4);
      mbar_init(&bars.vnst[ch], 6);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<C::TM_COLS>(&bars.tmem_base);
  // rows 0..63 of the [0 ; P] operand: written once, never touched by the copies
  for (int s = 0; s < C::NK; ++s)
    for (int i = tid; i < (int)P_BYTES / 16; i += C::THREADS)
      reinterpret_cast<uint4*>(smem + C::OFF_K + s * C::KSLOT + C::KS_OFF_Z)[i] = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp == 0) {
    // ------------------------------- copy warp (TMA engine) ---------------------------
    // Chunk c may be fetched once gdn_prep_kernel has published it (ws.ready, see gdn_layout.cuh).  The
    // 32 lanes look at 32 flags at a time, so a scan that runs behind prep polls once per 32 chunks.
    const uint32_t* ready = ws.ready + ch0;
    uint32_t* progress = ws.progress + ((size_t)b * H + h) * GDN_NS + slice;
    int known = 0;  // chunks [0, known) are published
    for (int c = 0; c < NT; ++c) {
      if (c >= known) {
        long long spins = 0;
#ifdef IVL_TRACE
        const long long tw0 = clock64();
#endif
        do {
          const int idx = known + lane;
          const uint32_t f = (idx < NT) ? ld_acquire_gpu(ready + idx) : 0u;
          const uint32_t m = __ballot_sync(0xffffffffu, f != 0u);
          known += (m == 0xffffffffu) ? 32 : (__ffs(~m) - 1);
          if (c >= known) {
            __nanosleep(200);
            if (++spins > (1ll << 24)) asm volatile("trap;");  // prep never ran: fail loudly instead of hanging
          }
        } while (c >= known);
        // every lane acquired its own flag (pairs with prep's release; the ballot orders the lanes), then the
        // generic -> async proxy fence.  (A gpu-scope FENCE here instead would also wait for this warp's
        // outstanding bulk copies at every poll.)
        asm volatile("fence.proxy.async;" ::: "memory");
#ifdef IVL_TRACE
        if (lane == 0) {
          atomicAdd(&ivl_scan_wait[0], (unsigned long long)(clock64() - tw0));
          atomicAdd(&ivl_scan_wait[1], 1ull);
          atomicAdd(&ivl_scan_wait[2], (unsigned long long)spins);
        }
#endif
      }
      const int sa = c % C::NA, sk = c % C::NK;
      const size_t cs = (size_t)((cb + c) % ring);  // image slot of chunk c
#ifdef IVL_TRACE
      if (lane == 0 && h < 16 && c < 2048 && slice < 4) {
        unsigned long long tg;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tg));
        ivl_scan_tl[(h * 2048 + c) * 4 + slice] = tg;
      }
#endif
      if (c >= C::NA) mbar_wait(&bars.emptyA[sa], (c / C::NA - 1) & 1);
      mbar_arrive_expect_tx_ws(&bars.fullA[sa], A1_BYTES);
      bulk_g2s_ws(smem + C::OFF_A + sa * A1_BYTES, blob + cs * BLOB_BYTES + BLOB_OFF_A1, A1_BYTES, &bars.fullA[sa]);
      if (c >= C::NK) {
        mbar_wait(&bars.emptyK[sk], (c / C::NK - 1) & 1);
        // every MMA that read chunk c - NK has retired (the A slot of that chunk was released even earlier), so
        // its image slot may be overwritten: tell prep (the ring hand-off of gdn_layout.cuh)
        // (relaxed is enough: the copies out of that slot have COMPLETED -- this warp saw their mbarrier -- and
        //  a release here would make the copy warp wait for its own outstanding bulk copies every chunk)
        if (lane == 0 && ring < NTROW)
          asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(progress), "r"((uint32_t)(c - C::NK + 1)) : "memory");
      }
      uint8_t* ks = smem + C::OFF_K + sk * C::KSLOT;
      mbar_arrive_expect_tx_ws(&bars.fullK[sk], P_BYTES + KT_BYTES + TAIL_BYTES + C::U_BYTES);
      bulk_g2s_ws(ks + C::KS_OFF_P, blob + cs * BLOB_BYTES + BLOB_OFF_P, P_BYTES + KT_BYTES + TAIL_BYTES, &bars.fullK[sk]);
      bulk_g2s_ws(ks + C::KS_OFF_U, ublob + cs * (GDN_NS * UBLOB_BYTES), C::U_BYTES, &bars.fullK[sk]);
    }
  } else if (warp <= C::NCG) {
    // ------------------------------- MMA issuers (one warp per chain) -----------------
    // All 32 lanes run this code converged; the elected lane issues (umma_*_ws), so the operands stay in
    // uniform registers.  Two barrier waits per chunk.
    {
      const int ch = warp - 1;
      [[maybe_unused]] const bool tr_on = ch == 0;
      constexpr uint32_t idescA = umma_idesc_bf16(128, 32, /*a_mn=*/0, /*b_mn=*/1);
      constexpr uint32_t idescB = umma_idesc_bf16(128, 32, /*a_mn=*/1, /*b_mn=*/1);
      const uint32_t sbase = smem_u32(smem);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0) + ch * C::TM_CHAIN;
      const uint64_t dSb = umma_desc(sbase + C::OFF_SB + ch * 8192, 128, 2048, SWZ_NONE);
      const uint64_t dVn = umma_desc(sbase + C::OFF_VN + ch * 4096, 128, 1024, SWZ_NONE);
      for (int c = 0; c < NT; ++c) {
        const int sa = c % C::NA, sk = c % C::NK, buf = c & 1;
        const uint32_t ks = sbase + C::OFF_K + sk * C::KSLOT;
        const uint64_t dA1 = umma_desc(sbase + C::OFF_A + sa * A1_BYTES, 128, 2048, SWZ_NONE);
        const uint64_t dKt = umma_desc(ks + C::KS_OFF_KT, 128, 1024, SWZ_NONE);
        const uint64_t dP = umma_desc(ks + C::KS_OFF_Z, 128, 1024, SWZ_NONE);
        const uint32_t d1 = tm + C::TM_D1 + buf * 32;
        mbar_wait(&bars.sb[ch], c & 1);
        tc_fence_after();
        TR(0);
#pragma unroll
        for (int j = 0; j < 8; ++j) umma_bf16_ws(d1, dA1 + j * 16, dSb + j * 16, idescA, j > 0);
        umma_commit_ws(&bars.a[ch][buf]);
        umma_commit_ws(&bars.emptyA[sa]);
        TR(1);
        mbar_wait(&bars.vnst[ch], c & 1);
        tc_fence_after();
        TR(3);
#pragma unroll
        for (int j = 0; j < 4; ++j) umma_bf16_ws(tm + C::TM_S, dKt + j * 16, dVn + j * 16, idescB, 1);
        umma_commit_ws(&bars.s[ch]);
#pragma unroll
        for (int j = 0; j < 4; ++j) umma_bf16_ws(d1, dP + j * 16, dVn + j * 16, idescA, 1);
        umma_commit_ws(&bars.o[ch][buf]);
        umma_commit_ws(&bars.emptyK[sk]);
        TR(4);
      }
    }
  } else {
    // ------------------------------- epilogue warps -----------------------------------
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int cg = (warp - 1 - C::NCG) >> 2;  // chain = 32-column group of the slice (4 consecutive warps)
    const int row = quad * 32 + lane;      // accumulator row == TMEM lane
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16) + cg * C::TM_CHAIN;
    [[maybe_unused]] const bool tr_on = cg == 0;
    uint64_t* const bar_sb = &bars.sb[cg];
    uint64_t* const bar_vnst = &bars.vnst[cg];
    uint64_t* const bar_s = &bars.s[cg];
    const bool is_vn = quad < 2;           // rows 0..63  : v_new
    const int tok = row & 63;              // token inside the chunk for both halves
    const int col0 = slice * BV + cg * 32; // first value column of this thread
    uint8_t* sb_dst = smem + C::OFF_SB + cg * (4 * 2048) + (row >> 3) * 128 + (row & 7) * 16;
    uint8_t* vn_dst = smem + C::OFF_VN + cg * (4 * 1024) + (tok >> 3) * 128 + (tok & 7) * 16;
    float x[32];
    uint32_t r[32];

    // initial state: S_0 -> bf16 shadow, gamma_0 S_0 -> TMEM
    {
      const size_t soff = (((size_t)seq * H + h) * GDN_K + row) * GDN_V + col0;
      if (h0 == nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = 0.f;
      } else if (h0_dtype == 0) {
        const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(h0) + soff);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 f = __ldg(p + i);
          x[4 * i] = f.x; x[4 * i + 1] = f.y; x[4 * i + 2] = f.z; x[4 * i + 3] = f.w;
        }
      } else {
        const uint4* p = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(h0) + soff);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u = __ldg(p + i);
          const uint32_t* w = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) { x[8 * i + 2 * e] = bf16_lo(w[e]); x[8 * i + 2 * e + 1] = bf16_hi(w[e]); }
        }
      }
      store_row_bf16(sb_dst, 2048, x);
      fence_async_smem();
      // the epilogue warps vouch for the operands on behalf of the MMA warp
      mbar_wait(&bars.fullA[0], 0);
      mbar_wait(&bars.fullK[0], 0);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sb);
      const float g0 = *reinterpret_cast<const float*>(smem + C::OFF_K + C::KS_OFF_TAIL);  // slot 0 has landed
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(x[i] * g0);
      tmem_st32(tlane + C::TM_S, r);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_vnst);
    }

    for (int c = 0; c < NT; ++c) {
      const int sk = c % C::NK, buf = c & 1;
      if (is_vn) {
        // v_new = U - Wg S   (A1 holds -Wg, so the accumulator is added)
        // (the K slot of chunk c was awaited in the S hand-over of chunk c-1)
        const uint8_t* usrc = smem + C::OFF_K + sk * C::KSLOT + C::KS_OFF_U + cg * (4 * 1024) + tok * 16;
        uint4 u[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) u[p] = *reinterpret_cast<const uint4*>(usrc + p * 1024);
        mbar_wait(&bars.a[cg][buf], (c >> 1) & 1);
        tc_fence_after();
        if (quad == 0) TR(5);
        tmem_ld32(tlane + C::TM_D1 + buf * 32, r);
        tmem_ld_wait();
        if (quad == 0) TR(6);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const uint32_t* w = reinterpret_cast<const uint32_t*>(&u[p]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            x[8 * p + 2 * e] = __uint_as_float(r[8 * p + 2 * e]) + bf16_lo(w[e]);
            x[8 * p + 2 * e + 1] = __uint_as_float(r[8 * p + 2 * e + 1]) + bf16_hi(w[e]);
          }
        }
        store_row_bf16(vn_dst, 1024, x);
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_vnst);
        if (quad == 0) TR(7);
      }
      // state hand-over: S_{c+1} is complete once MMA-B has retired.  While waiting for it, make sure the
      // next chunk's operands have landed (off the critical path here, and it spares the MMA warp a wait).
      float gnext = 1.f;
      if (c + 1 < NT) {
        mbar_wait(&bars.fullA[(c + 1) % C::NA], ((c + 1) / C::NA) & 1);
        mbar_wait(&bars.fullK[(c + 1) % C::NK], ((c + 1) / C::NK) & 1);
        gnext = *reinterpret_cast<const float*>(smem + C::OFF_K + ((c + 1) % C::NK) * C::KSLOT + C::KS_OFF_TAIL);
      }
      mbar_wait(bar_s, c & 1);
      tc_fence_after();
      if (quad == 0) TR(8);
      tmem_ld32(tlane + C::TM_S, r);
      tmem_ld_wait();
      if (quad == 0) TR(9);
      if (c + 1 < NT) {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(r[i]);
        store_row_bf16(sb_dst, 2048, x);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sb);     // MMA-A of the next chunk may start
        if (quad == 0) TR(10);
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(x[i] * gnext);
        tmem_st32(tlane + C::TM_S, r);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_vnst);   // MMA-B of the next chunk may accumulate
        if (quad == 0) TR(11);
      } else if (ht != nullptr) {
        const size_t soff = (((size_t)seq * H + h) * GDN_K + row) * GDN_V + col0;
        if (ht_dtype == 0) {
          float4* p = reinterpret_cast<float4*>(static_cast<float*>(ht) + soff);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            p[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                               __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(r[i]);
          store_row_bf16(reinterpret_cast<uint8_t*>(static_cast<__nv_bfloat16*>(ht) + soff), 16, x);
        }
      }
      if (!is_vn) {
        // output rows: O = Qg S + P Vn (scale and exp(G) are folded into Qg and P)
        mbar_wait(&bars.o[cg][buf], (c >> 1) & 1);
        tc_fence_after();
        if (quad == 2) TR(12);
        tmem_ld32(tlane + C::TM_D1 + buf * 32, r);
        tmem_ld_wait();  // D1[buf] is free again once this warp's next sb arrival is observed by the MMA warp
        const int t = varlen ? (tok < __ldg(vl.chunk_valid + cb + c) ? __ldg(vl.chunk_tok0 + cb + c) + tok : T)
                             : c * GDN_C + tok;
        if (t < T) {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(r[i]);
          store_row_bf16(reinterpret_cast<uint8_t*>(o + (((size_t)b * T + t) * H + h) * GDN_V + col0), 16, x);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TM_COLS>(tmem);
}

template <int BV>
cudaError_t launch_scan_variant(const GdnWorkspace& ws, const GdnVarlen& vl, int ntrow, int nseq, const void* h0,
                                int h0_dtype, void* o, void* ht, int ht_dtype, int T, int H, cudaStream_t stream) {
  using C = ScanCfg<BV>;
  static std::atomic<bool> configured[64];
  int dev = 0;
  if (cudaError_t e = cudaGetDevice(&dev)) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!configured[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(gdn_scan_kernel<BV>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    configured[dev].store(true, std::memory_order_release);
  }
  dim3 grid(GDN_V / BV, H, nseq);
  gdn_scan_kernel<BV><<<grid, C::THREADS, C::SMEM, stream>>>(ws, vl, h0, h0_dtype, static_cast<__nv_bfloat16*>(o), ht,
                                                             ht_dtype, T, H, ntrow);
  return cudaGetLastError();
}

}  // namespace

// bv: value columns per CTA (32, 64 or 128).  ntrow: chunks per workspace row; nseq: batch rows (dense) or
// sequences of the packed batch (vl tables set).
cudaError_t launch_gdn_scan(const GdnWorkspace& ws, const GdnVarlen& vl, int ntrow, int nseq, const void* h0,
                            int h0_dtype, void* o, void* ht, int ht_dtype, int T, int H, int bv, cudaStream_t stream) {
  switch (bv) {
    case 32: return launch_scan_variant<32>(ws, vl, ntrow, nseq, h0, h0_dtype, o, ht, ht_dtype, T, H, stream);
    case 64: return launch_scan_variant<64>(ws, vl, ntrow, nseq, h0, h0_dtype, o, ht, ht_dtype, T, H, stream);
    case 128: return launch_scan_variant<128>(ws, vl, ntrow, nseq, h0, h0_dtype, o, ht, ht_dtype, T, H, stream);
    default: return cudaErrorInvalidValue;
  }
}

#ifdef IVL_TRACE
extern "C" __attribute__((visibility("default"))) int ivl_debug_read_trace(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, ivl_trace_buf, sizeof(long long) * n);
}
extern "C" __attribute__((visibility("default"))) int ivl_debug_read_scan_tl(unsigned long long* host) {
  return (int)cudaMemcpyFromSymbol(host, ivl_scan_tl, sizeof(unsigned long long) * 16 * 2048 * 4);
}
extern "C" __attribute__((visibility("default"))) int ivl_debug_read_scan_wait(unsigned long long* host, int reset) {
  int e = (int)cudaMemcpyFromSymbol(host, ivl_scan_wait, sizeof(unsigned long long) * 4);
  if (reset) { unsigned long long z[4] = {0, 0, 0, 0}; cudaMemcpyToSymbol(ivl_scan_wait, z, sizeof(z)); }
  return e;
}
#endif

}  // namespace ivl
