// Gated DeltaNet inter-chunk scan: the serial part of the chunked delta rule, one
// persistent CTA per (batch, head, 32-column slice of V), the recurrent state slice
// S[128 x 32] kept on-chip (fp32 master in TMEM, bf16 shadow in shared memory) for the
// whole sequence.
//
// Replaces chunk_gated_delta_rule_fwd_h + chunk_fwd_o of the reference
// (src/llamafactory/model/fla/ops/common/chunk_delta_h.py:32-124,247-318 and
//  ops/common/chunk_o.py:32-114,456-497) without ever materialising the per-chunk
// states h[B,NT,H,K,V] (2.15 GB per layer at 128K tokens in the reference).
//
// Per chunk c (operands are the images written by gdn_prep.cu, loaded with 1-D bulk
// TMA copies through a 3-stage mbarrier ring):
//   MMA-A  D1 = [-Wg ; Qg] . bf16(S)            M128 N32 K128   (tcgen05, accum in TMEM)
//   epi    Vn = U + D1[0:64]      -> bf16 -> shared (MN-major B operand)
//   MMA-B  S  = gamma S + Kt^T . Vn             M128 N32 K64    (gamma pre-applied in TMEM)
//   MMA-C  D1[64:128] += P . Vn                 M128 N32 K64    (rows 0..63 of the A operand are zero)
//   epi    O = D1[64:128] -> bf16 -> global;  S -> bf16 shadow, gamma_{c+1} S -> TMEM
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (one thread), warps 2-5 epilogue
// (TMEM lane quadrant = warp % 4).
#include "gdn_layout.cuh"
#include "sm100.cuh"

namespace ivl {

namespace {

constexpr int SCAN_THREADS = 192;
constexpr int STAGES = 3;
constexpr uint32_t ST_OFF_P = 0;                        // [0 ; P] operand, first 8 KiB stay zero
constexpr uint32_t ST_OFF_BLOB = 8192;                  // blob lands here: P | A1 | Kt
constexpr uint32_t ST_OFF_A1 = ST_OFF_BLOB + BLOB_OFF_A1;
constexpr uint32_t ST_OFF_KT = ST_OFF_BLOB + BLOB_OFF_KT;
constexpr uint32_t ST_OFF_U = ST_OFF_BLOB + BLOB_BYTES;
constexpr uint32_t STAGE_BYTES = ST_OFF_U + UBLOB_BYTES;  // 69632
constexpr uint32_t SB_BYTES = 128 * GDN_BV * 2;           // bf16 shadow of S, MN-major B operand
constexpr uint32_t VN_BYTES = 64 * GDN_BV * 2;            // v_new, MN-major B operand
constexpr uint32_t OFF_SB = STAGES * STAGE_BYTES;
constexpr uint32_t OFF_VN = OFF_SB + SB_BYTES;
constexpr uint32_t OFF_BARS = OFF_VN + VN_BYTES;
constexpr uint32_t SCAN_SMEM = OFF_BARS + 256 + 1024;     // + alignment slack
static_assert(STAGE_BYTES % 1024 == 0, "stage alignment");
static_assert(SCAN_SMEM <= 232448, "exceeds 227 KiB");

constexpr uint32_t TM_D1 = 0;    // two accumulators of 32 columns
constexpr uint32_t TM_S = 64;    // state, 32 columns
constexpr uint32_t TM_COLS = 128;

// Developer-only timeline probe (compiled in with -DIVL_TRACE by tools/trace_scan.py; never in the product build)
#ifdef IVL_TRACE
__device__ long long ivl_trace_buf[64 * 16];
#define TR(slot)                                                                                         \
  do {                                                                                                   \
    if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && c >= 1000 && c < 1064)                        \
      ivl_trace_buf[(c - 1000) * 16 + (slot)] = clock64();                                               \
  } while (0)
#else
#define TR(slot) do { } while (0)
#endif

struct Bars {
  uint64_t full[STAGES], empty[STAGES];
  uint64_t a[2], o[2];
  uint64_t vnst, s, sb;
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

__global__ void __launch_bounds__(SCAN_THREADS, 1)
gdn_scan_kernel(GdnWorkspace ws, const void* __restrict__ h0, int h0_dtype, __nv_bfloat16* __restrict__ o,
                void* __restrict__ ht, int ht_dtype, int T, int H, int NT) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars& bars = *reinterpret_cast<Bars*>(smem + OFF_BARS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const size_t ch0 = ((size_t)b * H + h) * NT;
  const uint8_t* blob = ws.blob + ch0 * BLOB_BYTES;
  const uint8_t* ublob = ws.ublob + (ch0 * GDN_NS + slice) * UBLOB_BYTES;
  const float* gamma = ws.gamma + ch0;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
    // epilogue warps arrive once per warp (lane 0 after __syncwarp)
    for (int i = 0; i < 2; ++i) { mbar_init(&bars.a[i], 1); mbar_init(&bars.o[i], 1); }
    mbar_init(&bars.s, 1);
    // sb   (4 warps): bf16 shadow of S_c in shared memory + stage c landed  -> MMA-A(c) may be issued.
    //                 The O warps arrive here only after they finished reading D1 of chunk c-2, so the
    //                 accumulator buffer MMA-A(c) overwrites is free without a separate barrier.
    // vnst (2 + 4):   v_new(c) in shared memory (2 VN warps) and gamma_c S_c back in TMEM (4 warps)
    //                 -> MMA-B(c) / MMA-C(c) may be issued.
    mbar_init(&bars.sb, 4);
    mbar_init(&bars.vnst, 6);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TM_COLS>(&bars.tmem_base);
  // rows 0..63 of the [0 ; P] operand: written once, never touched by the copies
  for (int s = 0; s < STAGES; ++s)
    for (int i = tid; i < 8192 / 16; i += SCAN_THREADS)
      reinterpret_cast<uint4*>(smem + s * STAGE_BYTES + ST_OFF_P)[i] = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------------
    for (int c = 0; c < NT; ++c) {
      const int s = c % STAGES, it = c / STAGES;
      if (c >= STAGES) mbar_wait(&bars.empty[s], (it - 1) & 1);
      uint8_t* st = smem + s * STAGE_BYTES;
      mbar_arrive_expect_tx_ws(&bars.full[s], BLOB_BYTES + UBLOB_BYTES);
      bulk_g2s_ws(st + ST_OFF_BLOB, blob + (size_t)c * BLOB_BYTES, BLOB_BYTES, &bars.full[s]);
      bulk_g2s_ws(st + ST_OFF_U, ublob + (size_t)c * (GDN_NS * UBLOB_BYTES), UBLOB_BYTES, &bars.full[s]);
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------------
    // All 32 lanes run this code converged; the elected lane issues (umma_*_ws), so the operands stay in
    // uniform registers.  Two barrier waits per chunk.
    {
      constexpr uint32_t idescA = umma_idesc_bf16(128, GDN_BV, /*a_mn=*/0, /*b_mn=*/1);
      constexpr uint32_t idescB = umma_idesc_bf16(128, GDN_BV, /*a_mn=*/1, /*b_mn=*/1);
      const uint32_t sbase = smem_u32(smem);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint64_t dSb = umma_desc(sbase + OFF_SB, 128, 2048, SWZ_NONE);
      const uint64_t dVn = umma_desc(sbase + OFF_VN, 128, 1024, SWZ_NONE);
      for (int c = 0; c < NT; ++c) {
        const int s = c % STAGES, buf = c & 1;
        const uint32_t st = sbase + s * STAGE_BYTES;
        const uint64_t dA1 = umma_desc(st + ST_OFF_A1, 128, 2048, SWZ_NONE);
        const uint64_t dKt = umma_desc(st + ST_OFF_KT, 128, 1024, SWZ_NONE);
        const uint64_t dP = umma_desc(st + ST_OFF_P, 128, 1024, SWZ_NONE);
        const uint32_t d1 = tm + TM_D1 + buf * GDN_BV;
        mbar_wait(&bars.sb, c & 1);
        tc_fence_after();
        TR(0);
#pragma unroll
        for (int j = 0; j < 8; ++j) umma_bf16_ws(d1, dA1 + j * 16, dSb + j * 16, idescA, j > 0);
        umma_commit_ws(&bars.a[buf]);
        TR(1);
        mbar_wait(&bars.vnst, c & 1);
        tc_fence_after();
        TR(3);
#pragma unroll
        for (int j = 0; j < 4; ++j) umma_bf16_ws(tm + TM_S, dKt + j * 16, dVn + j * 16, idescB, 1);
        umma_commit_ws(&bars.s);
#pragma unroll
        for (int j = 0; j < 4; ++j) umma_bf16_ws(d1, dP + j * 16, dVn + j * 16, idescA, 1);
        umma_commit_ws(&bars.o[buf]);
        umma_commit_ws(&bars.empty[s]);
        TR(4);
      }
    }
  } else {
    // ------------------------------- epilogue warps -----------------------------------
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;      // accumulator row == TMEM lane
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const bool is_vn = quad < 2;           // rows 0..63  : v_new
    const int tok = row & 63;              // token inside the chunk for both halves
    uint8_t* sb_dst = smem + OFF_SB + (row >> 3) * 128 + (row & 7) * 16;
    uint8_t* vn_dst = smem + OFF_VN + (tok >> 3) * 128 + (tok & 7) * 16;
    float x[32];
    uint32_t r[32];

    // initial state: S_0 -> bf16 shadow, gamma_0 S_0 -> TMEM
    {
      const size_t soff = (((size_t)b * H + h) * GDN_K + row) * GDN_V + slice * GDN_BV;
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
      mbar_wait(&bars.full[0], 0);  // the epilogue warps vouch for the stage on behalf of the MMA warp
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.sb);
      const float g0 = __ldg(gamma);
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(x[i] * g0);
      tmem_st32(tlane + TM_S, r);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.vnst);
    }

    for (int c = 0; c < NT; ++c) {
      const int s = c % STAGES, buf = c & 1;
      const float gnext = (c + 1 < NT) ? __ldg(gamma + c + 1) : 1.f;
      if (is_vn) {
        // v_new = U - Wg S   (A1 holds -Wg, so the accumulator is added)
        const uint8_t* usrc = smem + s * STAGE_BYTES + ST_OFF_U + tok * 16;  // stage c was awaited in the S hand-over of c-1
        uint4 u[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) u[p] = *reinterpret_cast<const uint4*>(usrc + p * 1024);
        mbar_wait(&bars.a[buf], (c >> 1) & 1);
        tc_fence_after();
        if (quad == 0) TR(5);
        tmem_ld32(tlane + TM_D1 + buf * GDN_BV, r);
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
        if (lane == 0) mbar_arrive(&bars.vnst);
        if (quad == 0) TR(7);
      }
      // state hand-over: S_{c+1} is complete once MMA-B has retired.  While waiting for it, make sure the
      // next chunk's operands have landed (off the critical path here, and it spares the MMA warp a wait).
      if (c + 1 < NT) mbar_wait(&bars.full[(c + 1) % STAGES], ((c + 1) / STAGES) & 1);
      mbar_wait(&bars.s, c & 1);
      tc_fence_after();
      if (quad == 0) TR(8);
      tmem_ld32(tlane + TM_S, r);
      tmem_ld_wait();
      if (quad == 0) TR(9);
      if (c + 1 < NT) {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(r[i]);
        store_row_bf16(sb_dst, 2048, x);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.sb);   // MMA-A of the next chunk may start
        if (quad == 0) TR(10);
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(x[i] * gnext);
        tmem_st32(tlane + TM_S, r);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.vnst); // MMA-B of the next chunk may accumulate
        if (quad == 0) TR(11);
      } else if (ht != nullptr) {
        const size_t soff = (((size_t)b * H + h) * GDN_K + row) * GDN_V + slice * GDN_BV;
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
        mbar_wait(&bars.o[buf], (c >> 1) & 1);
        tc_fence_after();
        if (quad == 2) TR(12);
        tmem_ld32(tlane + TM_D1 + buf * GDN_BV, r);
        tmem_ld_wait();  // D1[buf] is free again once this warp's next sb arrival is observed by the MMA warp
        const int t = c * GDN_C + tok;
        if (t < T) {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(r[i]);
          store_row_bf16(reinterpret_cast<uint8_t*>(o + (((size_t)b * T + t) * H + h) * GDN_V + slice * GDN_BV), 16, x);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TM_COLS>(tmem);
}

}  // namespace

cudaError_t launch_gdn_scan(const GdnWorkspace& ws, const void* h0, int h0_dtype, void* o, void* ht, int ht_dtype,
                            int B, int T, int H, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gdn_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SCAN_SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid(GDN_NS, H, B);
  gdn_scan_kernel<<<grid, SCAN_THREADS, SCAN_SMEM, stream>>>(ws, h0, h0_dtype, static_cast<__nv_bfloat16*>(o), ht,
                                                             ht_dtype, T, H, gdn_num_chunks(T));
  return cudaGetLastError();
}

#ifdef IVL_TRACE
extern "C" __attribute__((visibility("default"))) int ivl_debug_read_trace(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, ivl_trace_buf, sizeof(long long) * n);
}
#endif

}  // namespace ivl
