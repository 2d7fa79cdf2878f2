// Gated DeltaNet chunk pre-pass: everything that is parallel over (batch, head, chunk).
//
// Replaces, in one launch, the reference's l2norm x2, chunk_local_cumsum,
// fwd_prepare_wy_repr and fwd_recompute_w_u kernels
// (src/llamafactory/model/fla/modules/l2norm.py:21-42, ops/utils/cumsum.py:26-59,
//  ops/gated_delta_rule/wy_fast.py:114-320) plus the intra-chunk attention matrix of
// chunk_fwd_kernel_o (ops/common/chunk_o.py:77-113), and emits the operands of the
// serial scan as ready-to-load tensor-core images (gdn_layout.cuh).
//
// Per (b, h, chunk c), with Qn, Kn the L2-normalised rows (rounded to bf16 as the
// reference does, l2norm.py:42), G the in-chunk cumsum of g, Gamma_ij = exp(G_i - G_j):
//   L  = tril_-1(diag(beta) Kn Kn^T * Gamma)                 fp32, tensor cores (mma.sync)
//   T  = (I + L)^-1                                           fp32 forward substitution
//   Wg = (T diag(beta exp G)) Kn,   U = (T diag(beta)) V      bf16 operands, fp32 accumulate
//   P  = tril(Qn Kn^T * Gamma) * scale,  Qg = Qn exp(G) scale,  Kt = Kn exp(G_C - G)
//
// This kernel is throughput work (32K independent CTAs at 128K tokens).  For the transposed scan's images (the product
// path) the three large products run on tcgen05 with TMEM accumulators (template flag TC below); the 16 x 16 tiles of
// the blocked inverse use split-bf16 warp-level MMAs; the row-major / lag image modes keep the warp-level path.
#include <atomic>
#include <stddef.h>
#include <stdlib.h>

#include "gdn_layout.cuh"
#include "sm100.cuh"

namespace ivl {

namespace {

constexpr int PREP_TC_DEFAULT = 1;   // see launch_gdn_prep
constexpr int PREP_THREADS = 256;  // 8 warps: warp w owns rows 16*(w>>1).. of the chunk and column half (w&1)
constexpr int KH_LD = 136;  // bf16 elements per row: 272 B, rows shift by 16 B mod 128 -> conflict-free ldmatrix
constexpr int V_LD = 136;   // V is staged in two halves of 128 value columns
constexpr int A_LD = 72;
constexpr int L_LD = 68;   // fp32 row stride of L and of T = (I + L)^-1

// 69.5 KiB, so that three CTAs are resident per SM (the kernel is a chain of short latency-bound
// phases; occupancy is what hides them).  Buffers are reused as their contents die:
//   qh : Qn rows                  -> T = (I + L)^-1 (fp32)        -> second half of V
//   LA : L (fp32)                 -> Aw | Au (bf16 A operands)
struct __align__(16) PrepSmem {
  __nv_bfloat16 kh[64 * KH_LD];
  __nv_bfloat16 qh[64 * KH_LD];
  __nv_bfloat16 vb[64 * V_LD];   // value columns 0..127
  float LA[2 * 64 * A_LD / 2];   // 18432 B >= 64 * L_LD floats
  float G[64];
  float beta[64];
  float Gp[64];                  // lag form: cumsum of the previous chunk's g
};
static_assert(sizeof(float) * 64 * L_LD <= sizeof(PrepSmem::LA), "L must fit");
static_assert(sizeof(__nv_bfloat16) * 64 * V_LD <= sizeof(PrepSmem::qh), "V half must fit in the q tile");

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                          uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem));
}

// Load 32 contiguous bf16 (one quarter of a 128-wide row), L2-normalise over the full row (the four
// threads of a row combine their partial sums), round to bf16 (the reference's rounding point), keep
// the rounded row in shared memory for the tensor-core products, and write the exponentially
// weighted copy straight into its operand image in global memory.
//   img_piece(p) returns the byte offset of 16-byte piece p (8 elements) of this thread's quarter row.
__device__ __forceinline__ void load_row_quarter(const __nv_bfloat16* src, bool valid, uint4* raw) {
#pragma unroll
  for (int p = 0; p < 4; ++p)
    raw[p] = valid ? __ldg(reinterpret_cast<const uint4*>(src) + p) : make_uint4(0, 0, 0, 0);
}

// Fused prologue: depthwise causal conv (kernel 4) + SiLU of 32 channels of one token row, rounded to bf16 exactly
// as ivl_short_conv_fwd does (gdn_fused.cu: same fp32 expression, same rounding point), from the RAW projection
// rows t-3 .. t.  x: the row of token t (this thread's 32 channels); row_stride: elements between token rows; t: index
// of the token in its sequence (rows before 0 come from the carried tail `cache` [D][4], newest in column 3, or are
// zero); w: conv weights of this thread's first channel ([ch][4]).
__device__ __forceinline__ void conv_row_quarter(const __nv_bfloat16* x, long long row_stride, int t, bool valid,
                                                 const __nv_bfloat16* w, const __nv_bfloat16* cache, uint4* raw) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (!valid) { raw[p] = make_uint4(0, 0, 0, 0); continue; }
    float acc[8];
    float wt[8][4];
    {
      const uint4* wp = reinterpret_cast<const uint4*>(w + p * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 u = __ldg(wp + i);
        const uint32_t* ww = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          wt[(i * 8 + 2 * e) / 4][(i * 8 + 2 * e) % 4] = bf16_lo(ww[e]);
          wt[(i * 8 + 2 * e + 1) / 4][(i * 8 + 2 * e + 1) % 4] = bf16_hi(ww[e]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {          // tap j multiplies the input of token t - 3 + j
      const int tt = t - 3 + j;
      float xin[8];
      if (tt >= 0) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (long long)(j - 3) * row_stride) + p);
        const uint32_t* xw = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) { xin[2 * e] = bf16_lo(xw[e]); xin[2 * e + 1] = bf16_hi(xw[e]); }
      } else if (cache != nullptr) {
#pragma unroll
        for (int e = 0; e < 8; ++e) xin[e] = __bfloat162float(cache[(p * 8 + e) * 4 + 4 + tt]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) xin[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = (j == 0) ? wt[e][0] * xin[e] : fmaf(wt[e][j], xin[e], acc[e]);
    }
    uint32_t* out = reinterpret_cast<uint32_t*>(&raw[p]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a0 = acc[2 * e], a1 = acc[2 * e + 1];
      out[e] = pack_bf16(a0 / (1.0f + __expf(-a0)), a1 / (1.0f + __expf(-a1)));
    }
  }
}

template <class SmemPiece, class PieceOffset>
__device__ __forceinline__ void norm_row_quarter(const uint4* raw, bool l2norm, float weight,
                                                 SmemPiece smem_piece, uint8_t* img, PieceOffset img_piece) {
  float ss = 0.f;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&raw[p]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float lo = bf16_lo(w[e]), hi = bf16_hi(w[e]);
      ss += lo * lo + hi * hi;
    }
  }
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
  const float rstd = l2norm ? 1.0f / sqrtf(ss + 1e-6f) : 1.0f;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&raw[p]);
    uint4 nrm, wgt;
    uint32_t* n = reinterpret_cast<uint32_t*>(&nrm);
    uint32_t* g = reinterpret_cast<uint32_t*>(&wgt);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      n[e] = pack_bf16(bf16_lo(w[e]) * rstd, bf16_hi(w[e]) * rstd);
      g[e] = pack_bf16(bf16_lo(n[e]) * weight, bf16_hi(n[e]) * weight);
    }
    *smem_piece(p) = nrm;
    *reinterpret_cast<uint4*>(img + img_piece(p)) = wgt;
  }
}

// Byte offsets of operand-image elements.  MODE 0 / 1 (gdn_scan.cu, first transposed scan): no-swizzle core-matrix
// tiling (gdn_layout.cuh).  MODE 2 (lag scan): the canonical 128-byte-swizzled layouts -- rows of 128 bytes (64
// elements of the contiguous dimension), 16-byte chunk index XORed with (row % 8); wider tiles are 64-element panels
// back to back.  The tensor core reads a no-swizzle B operand at roughly one 16-byte row piece per cycle (an N = 128
// K = 16 MMA was measured at ~200 cycles on such an image), a swizzled one at full rate.
template <int MODE>
__device__ __forceinline__ uint32_t a1_off(int row, int kd) {   // [-Wg ; Qg]: 128 rows x 128 key dims, K-major
  return MODE == 2 ? (uint32_t)((kd >> 6) * 16384) + swz128((uint32_t)(row * 128 + (kd & 63) * 2))
                   : (uint32_t)((row >> 3) * 2048 + (kd >> 3) * 128 + (row & 7) * 16 + (kd & 7) * 2);
}
template <int MODE>
__device__ __forceinline__ uint32_t kt_off(int tok, int kd) {   // Kt: 64 tokens x 128 key dims, MN-major (kd contiguous)
  return MODE == 2 ? (uint32_t)((kd >> 6) * 8192) + swz128((uint32_t)(tok * 128 + (kd & 63) * 2))
                   : (uint32_t)((kd >> 3) * 1024 + (tok >> 3) * 128 + (tok & 7) * 16 + (kd & 7) * 2);
}
template <int MODE>
__device__ __forceinline__ uint32_t sq_off(int i, int j) {      // P, Au, R: 64 rows x 64, K-major (j contiguous)
  return MODE == 2 ? swz128((uint32_t)(i * 128 + j * 2))
                   : (uint32_t)((i >> 3) * 1024 + (j >> 3) * 128 + (i & 7) * 16 + (j & 7) * 2);
}

#ifdef IVL_TRACE
__device__ long long ivl_prep_trace[16 * 8];
__device__ unsigned long long ivl_prep_wait[4];  // cycles waited on the ring, CTAs that waited, poll iterations
__device__ unsigned long long ivl_prep_tl[16 * 2048 * 4];  // head 0, chunk c: globaltimer at CTA start, after the ring wait, at publish
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PTR(slot)                                                                     \
  do {                                                                                \
    if (blockIdx.x >= 16000 && blockIdx.x < 16016 && threadIdx.x == 0)                \
      ivl_prep_trace[(blockIdx.x - 16000) * 8 + (slot)] = clock64();                  \
  } while (0)
#else
#define PTR(slot) do { } while (0)
#endif

// TR = true: operands for the transposed scan (gdn_scan_t.cu).  That kernel forms U = Au V itself from the raw
// value rows (one more small MMA per chunk on a tensor pipe it does not saturate), so this kernel neither reads V
// nor writes the eight U slices: it emits the 8 KiB Au image instead (a third less tensor work, half the loads
// and 36 % fewer image bytes per chunk).
//
// MODE 0: row-major scan (U slices).  MODE 1: transposed scan (Au image).  MODE 2: transposed scan with the
// shortened serial chain (gdn_scan_t.cu, lag form): v_new of chunk c is formed from the state BEFORE chunk c-1,
//   v_new_c = U_c - gamma_{c-1} Wg_c S_{c-1} - (Wg_c Kt_{c-1}^T) v_new_{c-1},
// so this kernel additionally scales the Wg rows by the previous chunk's decay gamma_{c-1} and emits the 64 x 64
// coupling matrix R_c = Wg_c Kt_{c-1}^T (negated, 8 KiB image behind Au); the first chunk of a sequence has
// gamma_{-1} = 1 and R = 0.  It recomputes Kt_{c-1} from the previous chunk's k rows and g (bit-identical to the
// image the previous chunk's CTA writes, which the state update uses).
// TC = true (MODE 1 only): the three large products of the chunk -- Kn Kn^T, Qn Kn^T (M64 N64 K128) and Wg = Aw Kn
// (M64 N128 K64) -- run on the tcgen05 tensor cores: the normalised rows are written to shared memory as 128-byte
// swizzled UMMA tiles (the SAME Kn tile is the K-major A / B operand of the first two products and the MN-major B operand
// of the third), accumulators live in 64 columns of tensor memory -- an M = 64 accumulator fills lanes 0..15 of every
// lane quadrant, a second one issued with lane offset 16 fills lanes 16..31 (tools/umma_probe.cu), so two products
// share the columns and all 32 lanes of the epilogue warps have a row to read with tcgen05.ld.
// The 16 x 16 tiles of the blocked inverse keep their split-bf16 warp-level MMAs (they need fp32 accuracy).
template <int MODE, bool FUSED = false, bool TC = false>
__global__ void __launch_bounds__(PREP_THREADS, 3)
gdn_prep_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                const __nv_bfloat16* __restrict__ v, const float* __restrict__ g,
                const __nv_bfloat16* __restrict__ beta, GdnWorkspace ws, GdnVarlen vl, int T, int H, float scale,
                int l2norm, int prefetch_ahead, int scan_ctas_per_head, GdnPrepFused fz) {
  constexpr bool TR = MODE != 0, LAG = MODE == 2;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // (TC: the swizzled operand tiles need 1 KiB alignment; kh, qh and vb sit at multiples of 17 KiB inside the struct)
  PrepSmem& s = *reinterpret_cast<PrepSmem*>(TC ? smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) : smem_raw);
  static_assert(offsetof(PrepSmem, qh) % 1024 == 0 && offsetof(PrepSmem, vb) % 1024 == 0, "tile alignment");
  __shared__ uint64_t tc_bar[2];
  __shared__ uint32_t tc_tmem;
  uint8_t* const tKh = reinterpret_cast<uint8_t*>(s.kh);   // TC: Kn tile, 2 panels [64 tok][128 B], 128B swizzle
  uint8_t* const tQh = reinterpret_cast<uint8_t*>(s.qh);   // TC: Qn tile (then T, as in the other path)
  uint8_t* const tAw = reinterpret_cast<uint8_t*>(s.vb);   // TC: Aw tile [64][128 B], 128B swizzle
  float* const sL = s.LA;                                                  // strictly lower triangular, fp32
  __nv_bfloat16* const sAw = reinterpret_cast<__nv_bfloat16*>(s.LA);       // after the solve
  __nv_bfloat16* const sAu = sAw + 64 * A_LD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (TC) {
    if (tid == 0) {
      mbar_init(&tc_bar[0], 1);
      mbar_init(&tc_bar[1], 1);
      fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<64>(&tc_tmem);
    tc_fence_before();
    // (the __syncthreads() at the end of stage 0 publishes the barriers and the TMEM address)
  }
  // CTAs are dispatched in linear block order: head fastest, then chunk, so the grid walks the sequence
  // front to back over all heads and a concurrently running scan (gdn_scan.cu) can follow it.
  const int c = blockIdx.x / H, h = blockIdx.x % H, b = blockIdx.z, NT = gridDim.x / H;
  const bool varlen = vl.chunk_tok0 != nullptr;  // packed sequences: chunk geometry comes from the tables
  const int t0 = varlen ? __ldg(vl.chunk_tok0 + c) : c * GDN_C;
  const int valid = varlen ? __ldg(vl.chunk_valid + c) : min(GDN_C, T - t0);
  const size_t tok0 = (size_t)b * T + t0;
  const size_t ch = ((size_t)b * H + h) * NT + c;  // chunk-head index (ready flag)
  const size_t slot = ((size_t)b * H + h) * ws.ring + (c % ws.ring);  // where its images live
  uint8_t* blob = ws.blob + slot * BLOB_BYTES;
  uint8_t* ublob = ws.ublob + slot * (GDN_NS * UBLOB_BYTES);
  // lag form: does this chunk have a predecessor in its own sequence?  (packed batches: c is the first chunk of a
  // sequence iff it appears in seq_chunk_begin -- binary search, the table is sorted)
  bool has_prev = false;
  if (LAG) {
    if (!varlen) {
      has_prev = c > 0;
    } else {
      int lo = 0, hi = vl.num_seqs;   // find the last n with seq_chunk_begin[n] <= c
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(vl.seq_chunk_begin + mid) <= c) lo = mid; else hi = mid - 1;
      }
      has_prev = __ldg(vl.seq_chunk_begin + lo) != c;
    }
  }
  const size_t ptok0 = (size_t)b * T + (has_prev ? (varlen ? __ldg(vl.chunk_tok0 + c - 1) : t0 - GDN_C) : 0);

#ifdef IVL_TRACE
  if (tid == 0 && h < 16 && c < 2048) ivl_prep_tl[(h * 2048 + c) * 4 + 0] = gtime();
#endif
  PTR(0);
  // ---- stage 0: issue every global load of the chunk up front (q, k rows into registers, V by cp.async,
  //      g / beta), then the chunk-local cumsum of g ------------------------------------------------
  uint4 rawq[4], rawk[4];
  {
    const int row = tid >> 2, qt = tid & 3;  // 4 threads per row, 32 elements each
    const size_t off = ((tok0 + row) * H + h) * GDN_K + qt * 32;
    if (FUSED) {
      // q, k are the RAW projection outputs: conv + SiLU here (dense batches only: t0 + row is the position in the sequence)
      const size_t ch = (size_t)h * GDN_K + qt * 32;
      const size_t D = (size_t)H * GDN_K;
      const __nv_bfloat16* cq = fz.cq_in ? static_cast<const __nv_bfloat16*>(fz.cq_in) + ((size_t)b * D + ch) * 4 : nullptr;
      const __nv_bfloat16* ck = fz.ck_in ? static_cast<const __nv_bfloat16*>(fz.ck_in) + ((size_t)b * D + ch) * 4 : nullptr;
      conv_row_quarter(q + off, (long long)H * GDN_K, t0 + row, row < valid,
                       static_cast<const __nv_bfloat16*>(fz.wq) + ch * 4, cq, rawq);
      conv_row_quarter(k + off, (long long)H * GDN_K, t0 + row, row < valid,
                       static_cast<const __nv_bfloat16*>(fz.wk) + ch * 4, ck, rawk);
    } else {
      load_row_quarter(q + off, row < valid, rawq);
      load_row_quarter(k + off, row < valid, rawk);
    }
  }
  if (FUSED && c == NT - 1 && fz.cq_out != nullptr) {
    // carried conv tails after this call: the last four raw inputs of [tail_in | x] per channel (gdn_fused.cu)
    const int chl = tid & 127;
    const size_t D = (size_t)H * GDN_K, ch = (size_t)h * GDN_K + chl;
    const __nv_bfloat16* x = tid < 128 ? q : k;
    const __nv_bfloat16* cin = static_cast<const __nv_bfloat16*>(tid < 128 ? fz.cq_in : fz.ck_in);
    __nv_bfloat16* cout = static_cast<__nv_bfloat16*>(tid < 128 ? fz.cq_out : fz.ck_out) + ((size_t)b * D + ch) * 4;
    float tl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int tt = T - 4 + j;
      tl[j] = tt >= 0 ? __bfloat162float(x[(((size_t)b * T + tt) * H + h) * GDN_K + chl])
                      : (cin ? __bfloat162float(cin[((size_t)b * D + ch) * 4 + 4 + tt]) : 0.f);
    }
    uint2 o2;
    o2.x = pack_bf16(tl[0], tl[1]);
    o2.y = pack_bf16(tl[2], tl[3]);
    *reinterpret_cast<uint2*>(cout) = o2;
  }
  {
    // Warm L2 for the CTA that will run on this SM slot one wave later: its q/k/v/g/beta lines then hit
    // L2 instead of paying HBM latency at the head of a short CTA.
    const long long ahead = (long long)blockIdx.x + prefetch_ahead;
    const int pc = (int)(ahead / H), ph = (int)(ahead % H);
    if (pc < NT) {
      const int pt0 = varlen ? __ldg(vl.chunk_tok0 + pc) : pc * GDN_C;
      const size_t ptok = (size_t)b * T + (size_t)pt0;
      const int row = tid >> 2, qt = tid & 3;
      if (pt0 + row < T) {
        const size_t poff = ((ptok + row) * H + ph) * GDN_K + qt * 32;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q + poff));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(k + poff));
        if (!TR) asm volatile("prefetch.global.L2 [%0];" ::"l"(v + ((ptok + row) * H + ph) * GDN_V + qt * 64));
        if (qt == 0 && !FUSED) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(g + (ptok + row) * H + ph));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(beta + (ptok + row) * H + ph));
        }
      }
    }
  }
  if (!TR) {
    for (int i = tid; i < 64 * 16; i += PREP_THREADS) {
      const int row = i >> 4, piece = i & 15;
      __nv_bfloat16* dst = &s.vb[row * V_LD + piece * 8];
      if (row < valid)
        cp_async16(dst, v + ((tok0 + row) * H + h) * GDN_V + piece * 8);
      else
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int i = tid; i < 64 * L_LD; i += PREP_THREADS) sL[i] = 0.f;
  if (warp == 0) {
    float g0, g1;
    if (FUSED) {
      // g = -exp(A_log) softplus(a + dt_bias), beta = sigmoid(b) rounded to bf16: the expressions of gdn_gate_kernel
      const __nv_bfloat16* ar = static_cast<const __nv_bfloat16*>(fz.a);
      const float dtb = fz.dt_bias[h], na = -expf(fz.A_log[h]);
      auto gate = [&](int i) {
        const float z = __bfloat162float(ar[(tok0 + i) * H + h]) + dtb;
        return na * (z > 20.f ? z : log1pf(expf(z)));
      };
      g0 = (lane < valid) ? gate(lane) : 0.f;
      g1 = (lane + 32 < valid) ? gate(lane + 32) : 0.f;
    } else {
      g0 = (lane < valid) ? g[(tok0 + lane) * H + h] : 0.f;
      g1 = (lane + 32 < valid) ? g[(tok0 + lane + 32) * H + h] : 0.f;
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      float a0 = __shfl_up_sync(0xffffffffu, g0, d), a1 = __shfl_up_sync(0xffffffffu, g1, d);
      if (lane >= d) { g0 += a0; g1 += a1; }
    }
    g1 += __shfl_sync(0xffffffffu, g0, 31);
    s.G[lane] = g0;
    s.G[lane + 32] = g1;
    if (FUSED) {
      const __nv_bfloat16* br = static_cast<const __nv_bfloat16*>(fz.b);
      auto sig = [&](int i) {
        const float bb = __bfloat162float(br[(tok0 + i) * H + h]);
        return __bfloat162float(__float2bfloat16(1.0f / (1.0f + expf(-bb))));
      };
      s.beta[lane] = (lane < valid) ? sig(lane) : 0.f;
      s.beta[lane + 32] = (lane + 32 < valid) ? sig(lane + 32) : 0.f;
    } else {
      s.beta[lane] = (lane < valid) ? __bfloat162float(beta[(tok0 + lane) * H + h]) : 0.f;
      s.beta[lane + 32] = (lane + 32 < valid) ? __bfloat162float(beta[(tok0 + lane + 32) * H + h]) : 0.f;
    }
  }
  if (LAG && warp == 2) {
    // in-chunk cumsum of the previous chunk's g (a full chunk: only the last chunk of a sequence can be short)
    float g0 = has_prev ? g[(ptok0 + lane) * H + h] : 0.f;
    float g1 = has_prev ? g[(ptok0 + lane + 32) * H + h] : 0.f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      float a0 = __shfl_up_sync(0xffffffffu, g0, d), a1 = __shfl_up_sync(0xffffffffu, g1, d);
      if (lane >= d) { g0 += a0; g1 += a1; }
    }
    g1 += __shfl_sync(0xffffffffu, g0, 31);
    s.Gp[lane] = g0;
    s.Gp[lane + 32] = g1;
  }
  if (warp == 1 && c >= ws.ring) {
    // Ring hand-off (overlapped form only): the slot still holds chunk c - ring until every scan CTA of this
    // head has consumed it.  CTAs are dispatched in chunk order, so everything the scan is waiting for is
    // resident or done and this wait cannot deadlock; it is normally already satisfied.
    const uint32_t need = (uint32_t)(c - ws.ring + 1);
#ifdef IVL_TRACE
    const long long tw0 = clock64();
#endif
    const uint32_t* prog = ws.progress + ((size_t)b * H + h) * GDN_NS;
    long long spins = 0;
    for (;;) {
      uint32_t pv = 0xffffffffu;
      // relaxed polls (an acquire load costs an L1 invalidate per iteration), one acquire fence at the end
      if (lane < scan_ctas_per_head)
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(pv) : "l"(prog + lane) : "memory");
      if (__all_sync(0xffffffffu, pv >= need)) break;
      // back off in proportion to the distance (a scan step is ~1.2 us): frequent polls of the one line the
      // scan's copy warps keep writing were measured to slow the SCAN down by 2x
      const uint32_t behind = need - __reduce_min_sync(0xffffffffu, pv);
      __nanosleep(behind > 16 ? 16000 : behind * 1000);
      if (++spins > (1ll << 23)) asm volatile("trap;");  // the scan is not running: fail loudly, do not hang
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
#ifdef IVL_TRACE
    if (lane == 0) {
      atomicAdd(&ivl_prep_wait[0], (unsigned long long)(clock64() - tw0));
      atomicAdd(&ivl_prep_wait[1], 1ull);
      atomicAdd(&ivl_prep_wait[2], (unsigned long long)spins);
    }
#endif
  }
  __syncthreads();
#ifdef IVL_TRACE
  if (tid == 0 && h < 16 && c < 2048) ivl_prep_tl[(h * 2048 + c) * 4 + 1] = gtime();
#endif
  PTR(1);

  // ---- stage 1: normalise q, k rows; emit Qg and Kt images --------------------------------
  {
    const int row = tid >> 2, qt = tid & 3;
    const float Gr = s.G[row], Gc = s.G[63];
    const int R = 64 + row;  // Qg occupies rows 64..127 of the stacked [-Wg ; Qg] operand
    // where piece p (8 elements, 16 B) of this thread's quarter row goes in shared memory: padded rows for the
    // ldmatrix path, the 128-byte-swizzled UMMA tile (panel = 64 key dims) for the tcgen05 path
    auto piece = [&](uint8_t* tile, __nv_bfloat16* padded, int p) {
      const int kd = qt * 32 + p * 8;
      return TC ? reinterpret_cast<uint4*>(tile + (kd >> 6) * 8192 + swz128((uint32_t)(row * 128 + (kd & 63) * 2)))
                : reinterpret_cast<uint4*>(padded + row * KH_LD + kd);
    };
    norm_row_quarter(rawq, l2norm != 0, __expf(Gr) * scale, [&](int p) { return piece(tQh, s.qh, p); },
                     blob + BLOB_OFF_A1, [&](int p) { return a1_off<MODE>(R, (qt * 4 + p) * 8); });
    norm_row_quarter(rawk, l2norm != 0, __expf(Gc - Gr), [&](int p) { return piece(tKh, s.kh, p); },
                     blob + BLOB_OFF_KT, [&](int p) { return kt_off<MODE>(row, (qt * 4 + p) * 8); });
    if (tid == 0) *reinterpret_cast<float*>(blob + BLOB_OFF_TAIL) = __expf(Gc);
  }
  if (LAG && has_prev) {
    // Kt_{c-1} rows (the previous chunk's normalised keys with their decay to the end of that chunk), bf16, into
    // the value staging tile (unused in the transposed modes): the B operand of R = Wg Kt_{c-1}^T below
    const int row = tid >> 2, qt = tid & 3;
    uint4 rawp[4];
    load_row_quarter(k + ((ptok0 + row) * H + h) * GDN_K + qt * 32, true, rawp);
    float ss = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&rawp[p]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float lo = bf16_lo(w[e]), hi = bf16_hi(w[e]);
        ss += lo * lo + hi * hi;
      }
    }
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    const float rstd = l2norm ? 1.0f / sqrtf(ss + 1e-6f) : 1.0f;
    const float weight = __expf(s.Gp[63] - s.Gp[row]);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&rawp[p]);
      uint4 wgt;
      uint32_t* gg = reinterpret_cast<uint32_t*>(&wgt);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t n = pack_bf16(bf16_lo(w[e]) * rstd, bf16_hi(w[e]) * rstd);
        gg[e] = pack_bf16(bf16_lo(n) * weight, bf16_hi(n) * weight);
      }
      *reinterpret_cast<uint4*>(&s.vb[row * V_LD + qt * 32 + p * 8]) = wgt;
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  PTR(2);

  const int gq = lane >> 2, tq = lane & 3;  // mma fragment coordinates
  const int strip = warp >> 1, half = warp & 1;
  const int r0 = strip * 16;                // this warp's 16-row strip

  // ---- stage 2: Kn Kn^T and Qn Kn^T (lower triangle only) -> L (fp32, smem), P image -------
  if (TC) {
    // the tiles were written through the generic proxy: make them visible to the tensor core (async proxy)
    fence_async_smem();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tc_tmem;
    if (warp == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(64, 64, 0, 0);
      const uint64_t dK = umma_desc(smem_u32(tKh), 16, 1024, SWZ_128B);
      const uint64_t dQ = umma_desc(smem_u32(tQh), 16, 1024, SWZ_128B);
#pragma unroll
      for (int j = 0; j < 8; ++j) {   // K = 128 key dims: 2 panels x 4 slabs of 16
        const uint32_t off = (uint32_t)((j >> 2) * 512 + (j & 3) * 2);
        // an M = 64 accumulator fills lanes 0..15 of every 32-lane quadrant; with lane offset 16 it fills lanes 16..31:
        // the two products share the 64 columns, and every lane of the epilogue warps has a row to work on
        umma_bf16_ws(tm, dK + off, dK + off, idesc, j > 0);                  // lanes 32q + 0..15 : Kn Kn^T
        umma_bf16_ws(tm + (16u << 16), dQ + off, dK + off, idesc, j > 0);    // lanes 32q + 16..31: Qn Kn^T
      }
      umma_commit_ws(&tc_bar[0]);
    }
    mbar_wait(&tc_bar[0], 0);
    tc_fence_after();
    // warp w: lane quadrant w & 3, columns 32 (w >> 2) .. + 31; lane l: product l >> 4 (0: Kn Kn^T -> L, 1: Qn Kn^T -> P),
    // row 16 q + (l & 15)
    const int q4 = warp & 3, ch = warp >> 2, sub = lane >> 4;
    const int i = q4 * 16 + (lane & 15);
    const float Gi = s.G[i], bi = s.beta[i];
    uint32_t r[32];
    tmem_ld32(tm + ((uint32_t)(q4 * 32) << 16) + ch * 32, r);
    tmem_ld_wait();
    float val[32];
#pragma unroll
    for (int jj = 0; jj < 32; ++jj)
      val[jj] = __uint_as_float(r[jj]) * __expf(fminf(Gi - s.G[ch * 32 + jj], 0.f));
    if (sub == 0) {
#pragma unroll
      for (int jj = 0; jj < 32; ++jj)
        if (ch * 32 + jj < i) sL[i * L_LD + ch * 32 + jj] = bi * val[jj];
    } else {
      uint8_t* pimg = blob + BLOB_OFF_P;
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        uint32_t w4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j0 = ch * 32 + g8 * 8 + 2 * e;
          w4[e] = pack_bf16(j0 <= i ? val[g8 * 8 + 2 * e] * scale : 0.f, j0 + 1 <= i ? val[g8 * 8 + 2 * e + 1] * scale : 0.f);
        }
        *reinterpret_cast<uint4*>(pimg + sq_off<MODE>(i, ch * 32 + g8 * 8)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      }
    }
    tc_fence_before();
  } else {
  // warp (strip, half) computes column tiles 4*half .. 4*half+3 (32 key columns)
    float ckk[4][4], cqk[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) ckk[i][e] = cqk[i][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t ak[4], aq[4];
      const int arow = r0 + (lane & 15), acol = ks * 16 + (lane >> 4) * 8;
      ldsm_x4(smem_u32(&s.kh[arow * KH_LD + acol]), ak[0], ak[1], ak[2], ak[3]);
      ldsm_x4(smem_u32(&s.qh[arow * KH_LD + acol]), aq[0], aq[1], aq[2], aq[3]);
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
        const int ntp = half * 2 + pp;  // pair of 8-column tiles
        if (ntp <= strip) {             // column tiles right of the diagonal are never needed
          uint32_t b0, b1, b2, b3;
          const int n = ntp * 16 + (lane & 7) + (lane >> 4) * 8, kk = ks * 16 + ((lane >> 3) & 1) * 8;
          ldsm_x4(smem_u32(&s.kh[n * KH_LD + kk]), b0, b1, b2, b3);
          mma16816(ckk[2 * pp], ak, b0, b1);
          mma16816(ckk[2 * pp + 1], ak, b2, b3);
          mma16816(cqk[2 * pp], aq, b0, b1);
          mma16816(cqk[2 * pp + 1], aq, b2, b3);
        }
      }
    }
    const int i0 = r0 + gq, i1 = i0 + 8;
    const float G0 = s.G[i0], G1 = s.G[i1], be0 = s.beta[i0], be1 = s.beta[i1];
#pragma unroll
    for (int lt = 0; lt < 4; ++lt) {
      const int nt = half * 4 + lt;
      const int j0 = nt * 8 + 2 * tq, j1 = j0 + 1;
      float p00 = 0.f, p01 = 0.f, p10 = 0.f, p11 = 0.f;
      if (nt <= 2 * strip + 1) {
        const float Gj0 = s.G[j0], Gj1 = s.G[j1];
        const float e00 = __expf(fminf(G0 - Gj0, 0.f)), e01 = __expf(fminf(G0 - Gj1, 0.f));
        const float e10 = __expf(fminf(G1 - Gj0, 0.f)), e11 = __expf(fminf(G1 - Gj1, 0.f));
        // strictly-lower entries of L; entries on/above the diagonal stay zero
        if (i0 > j0) sL[i0 * L_LD + j0] = be0 * ckk[lt][0] * e00;
        if (i0 > j1) sL[i0 * L_LD + j1] = be0 * ckk[lt][1] * e01;
        if (i1 > j0) sL[i1 * L_LD + j0] = be1 * ckk[lt][2] * e10;
        if (i1 > j1) sL[i1 * L_LD + j1] = be1 * ckk[lt][3] * e11;
        p00 = (i0 >= j0) ? cqk[lt][0] * e00 * scale : 0.f;
        p01 = (i0 >= j1) ? cqk[lt][1] * e01 * scale : 0.f;
        p10 = (i1 >= j0) ? cqk[lt][2] * e10 * scale : 0.f;
        p11 = (i1 >= j1) ? cqk[lt][3] * e11 * scale : 0.f;
      }
      uint8_t* pimg = blob + BLOB_OFF_P;
      *reinterpret_cast<uint32_t*>(pimg + sq_off<MODE>(i0, nt * 8 + 2 * tq)) = pack_bf16(p00, p01);
      *reinterpret_cast<uint32_t*>(pimg + sq_off<MODE>(i1, nt * 8 + 2 * tq)) = pack_bf16(p10, p11);
    }
  }
  __syncthreads();
  PTR(3);

  // ---- stage 3: T = (I + L)^-1, blocked 4 x 4 in 16 x 16 tiles -------------------------------------
  //   diagonal tiles  T_ii = (I + L_ii)^-1                      fp32 forward substitution (16 steps)
  //   below diagonal  T_ij = -T_ii sum_{k=j}^{i-1} L_ik T_kj    tensor cores, operands split hi + lo in
  //                                                             bf16 (3 MMAs per product ~ fp32 accuracy)
  // T lives in the (now dead) q tile; the reference likewise solves in fp32 and rounds the result to
  // bf16 afterwards (wy_fast.py:188-210,342-343; 16 x 16 tiles as in pip-fla's solve_tril).
  float* Tm = reinterpret_cast<float*>(s.qh);
  static_assert(sizeof(float) * 64 * L_LD <= sizeof(__nv_bfloat16) * 64 * KH_LD, "T must fit in the q tile");
  // tiles above the diagonal are zero (6 tiles x 256 entries); everything else is written below
  for (int e = tid; e < 6 * 256; e += PREP_THREADS) {
    const int t6 = e >> 8, rr = (e >> 4) & 15, cc = e & 15;
    const int ti = (t6 < 3) ? 0 : (t6 < 5 ? 1 : 2);
    const int tj = (t6 < 3) ? t6 + 1 : (t6 < 5 ? t6 - 1 : 3);
    Tm[(ti * 16 + rr) * L_LD + tj * 16 + cc] = 0.f;
  }
  if (tid < 64) {
    // one column of one diagonal tile per thread
    const int blk = tid >> 4, col = tid & 15, base = blk * 16;
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < i; ++j) acc = fmaf(sL[(base + i) * L_LD + base + j], x[j], acc);
      x[i] = ((i == col) ? 1.f : 0.f) - acc;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) Tm[(base + i) * L_LD + base + col] = x[i];
  }
  __syncthreads();
  PTR(4);
  if (warp < 6) {
    // warps 2j, 2j+1 own tile column j (8 of its 16 columns each -- columns of T are independent) and walk
    // down it; each 16x16x8 product is 3 split MMAs
    const int j = warp >> 1, nh = warp & 1;
    auto split = [](float a, float b, uint32_t& hi, uint32_t& lo) {
      const __nv_bfloat16 ha = __float2bfloat16(a), hb = __float2bfloat16(b);
      hi = (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
      lo = pack_bf16(a - __bfloat162float(ha), b - __bfloat162float(hb));
    };
    // acc[4] += X(rows xr..+15, cols xc..+15) * Y(rows yr..+15, cols yc..+7), X and Y fp32 in shared memory
    auto tile_mma = [&](float* acc, const float* X, int xr, int xc, const float* Y, int yr, int yc) {
      uint32_t ah[4], al[4];
      const float* xa = X + (xr + gq) * L_LD + xc + 2 * tq;
      split(xa[0], xa[1], ah[0], al[0]);
      split(xa[8 * L_LD], xa[8 * L_LD + 1], ah[1], al[1]);
      split(xa[8], xa[9], ah[2], al[2]);
      split(xa[8 * L_LD + 8], xa[8 * L_LD + 9], ah[3], al[3]);
      const float* yb = Y + (yr + 2 * tq) * L_LD + yc + gq;
      uint32_t bh0, bl0, bh1, bl1;
      split(yb[0], yb[L_LD], bh0, bl0);
      split(yb[8 * L_LD], yb[9 * L_LD], bh1, bl1);
      mma16816(acc, ah, bh0, bh1);
      mma16816(acc, ah, bl0, bl1);
      mma16816(acc, al, bh0, bh1);
    };
    const int nc = j * 16 + nh * 8;  // first of this warp's 8 columns
    for (int i = j + 1; i < 4; ++i) {
      float m[4] = {0.f, 0.f, 0.f, 0.f};
      for (int kb = j; kb < i; ++kb) tile_mma(m, sL, i * 16, kb * 16, Tm, kb * 16, nc);
      // stage M_ij in tile (i, j) of T (not read by anyone else), then T_ij = -T_ii M_ij
      float* dst = Tm + (i * 16) * L_LD + nc;
      dst[gq * L_LD + 2 * tq] = m[0];
      dst[gq * L_LD + 2 * tq + 1] = m[1];
      dst[(gq + 8) * L_LD + 2 * tq] = m[2];
      dst[(gq + 8) * L_LD + 2 * tq + 1] = m[3];
      __syncwarp();
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      tile_mma(t, Tm, i * 16, i * 16, Tm, i * 16, nc);
      __syncwarp();
      dst[gq * L_LD + 2 * tq] = -t[0];
      dst[gq * L_LD + 2 * tq + 1] = -t[1];
      dst[(gq + 8) * L_LD + 2 * tq] = -t[2];
      dst[(gq + 8) * L_LD + 2 * tq + 1] = -t[3];
      __syncwarp();
    }
  }
  __syncthreads();
  PTR(5);
  // T -> the two bf16 A operands: Aw = T diag(beta exp G), Au = T diag(beta); a thread keeps one column
  {
    const int cc = tid & 63, rbase = tid >> 6;
    const float bu = s.beta[cc], bw = bu * __expf(s.G[cc]);
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const int i = rbase + 4 * it;
      const float tv = Tm[i * L_LD + cc];
      if (TC) *reinterpret_cast<__nv_bfloat16*>(tAw + swz128((uint32_t)(i * 128 + cc * 2))) = __float2bfloat16(tv * bw);
      else sAw[i * A_LD + cc] = __float2bfloat16(tv * bw);
      sAu[i * A_LD + cc] = __float2bfloat16(tv * bu);
    }
  }
  __syncthreads();
  PTR(6);
  // T is dead: fetch value columns 128..255 into its place while Wg and the first half of U are computed
  __nv_bfloat16* const vb2 = s.qh;
  if (!TR) {
    for (int i = tid; i < 64 * 16; i += PREP_THREADS) {
      const int row = i >> 4, piece = i & 15;
      __nv_bfloat16* dst = &vb2[row * V_LD + piece * 8];
      if (row < valid)
        cp_async16(dst, v + ((tok0 + row) * H + h) * GDN_V + 128 + piece * 8);
      else
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  } else {
    // Au image for the transposed scan: K-major no-swizzle [64 rows i][64 k = j] (core (i/8, j/8) at
    // (i/8) * 1024 + (j/8) * 128), 16-byte pieces, eight consecutive threads fill one 128-byte core matrix
    for (int e = tid; e < 512; e += PREP_THREADS) {
      const int r = e & 7, kg = (e >> 3) & 7, rg = e >> 6, i = rg * 8 + r;
      *reinterpret_cast<uint4*>(ublob + sq_off<MODE>(i, kg * 8)) =
          *reinterpret_cast<const uint4*>(&sAu[i * A_LD + kg * 8]);
    }
  }
  // ---- stage 4: Wg = Aw Kn (negated, into rows 0..63 of the A1 image), U = Au V ------------
  if (TC) {
    fence_async_smem();      // the Aw tile was written through the generic proxy
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tc_tmem;
    if (warp == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(64, 64, 0, /*b_mn=*/1);
      const uint64_t dA = umma_desc(smem_u32(tAw), 16, 1024, SWZ_128B);            // K-major: 16 tokens = 32 B per MMA
      const uint64_t dB0 = umma_desc(smem_u32(tKh), 8192, 1024, SWZ_128B);         // MN-major panel: key dims 0..63
      const uint64_t dB1 = umma_desc(smem_u32(tKh) + 8192, 8192, 1024, SWZ_128B);  // key dims 64..127
#pragma unroll
      for (int j = 0; j < 4; ++j) {   // K = 64 tokens: 16 per MMA = 2 KiB of the MN-major tile
        umma_bf16_ws(tm, dA + j * 2, dB0 + j * 128, idesc, j > 0);                 // lanes 32q + 0..15 : key dims 0..63
        umma_bf16_ws(tm + (16u << 16), dA + j * 2, dB1 + j * 128, idesc, j > 0);   // lanes 32q + 16..31: key dims 64..127
      }
      umma_commit_ws(&tc_bar[1]);
    }
    mbar_wait(&tc_bar[1], 0);
    tc_fence_after();
    const int q4 = warp & 3, ch = warp >> 2;           // lane quadrant, 32-column group
    const int i = q4 * 16 + (lane & 15);
    const int kd0 = (lane >> 4) * 64 + ch * 32;        // first of this thread's 32 key dims
    uint32_t r[32];
    tmem_ld32(tm + ((uint32_t)(q4 * 32) << 16) + ch * 32, r);
    tmem_ld_wait();
    uint8_t* img = blob + BLOB_OFF_A1;
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      uint4 w4;
      w4.x = pack_bf16(-__uint_as_float(r[g8 * 8 + 0]), -__uint_as_float(r[g8 * 8 + 1]));
      w4.y = pack_bf16(-__uint_as_float(r[g8 * 8 + 2]), -__uint_as_float(r[g8 * 8 + 3]));
      w4.z = pack_bf16(-__uint_as_float(r[g8 * 8 + 4]), -__uint_as_float(r[g8 * 8 + 5]));
      w4.w = pack_bf16(-__uint_as_float(r[g8 * 8 + 6]), -__uint_as_float(r[g8 * 8 + 7]));
      *reinterpret_cast<uint4*>(img + a1_off<MODE>(i, kd0 + g8 * 8)) = w4;
    }
    tc_fence_before();
  } else {
  // warp (strip, half): key dims 64*half..+63 of Wg; value columns 128*hv + 64*half..+63 of U in pass hv
    const int i0 = r0 + gq, i1 = i0 + 8;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      if (ks <= strip) {  // T is lower triangular
        uint32_t a[4];
        ldsm_x4(smem_u32(&sAw[(r0 + (lane & 15)) * A_LD + ks * 16 + (lane >> 4) * 8]), a[0], a[1], a[2], a[3]);
        const int kk = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int ntp = 0; ntp < 4; ++ntp) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(smem_u32(&s.kh[kk * KH_LD + half * 64 + ntp * 16 + (lane >> 4) * 8]), b0, b1, b2, b3);
          mma16816(acc[2 * ntp], a, b0, b1);
          mma16816(acc[2 * ntp + 1], a, b2, b3);
        }
      }
    }
    const float wsc = (LAG && has_prev) ? -__expf(s.Gp[63]) : -1.f;   // lag form: - gamma_{c-1} Wg
#pragma unroll
    for (int lt = 0; lt < 8; ++lt) {
      const int nt = half * 8 + lt;
      uint8_t* img = blob + BLOB_OFF_A1;
      *reinterpret_cast<uint32_t*>(img + a1_off<MODE>(i0, nt * 8 + 2 * tq)) = pack_bf16(wsc * acc[lt][0], wsc * acc[lt][1]);
      *reinterpret_cast<uint32_t*>(img + a1_off<MODE>(i1, nt * 8 + 2 * tq)) = pack_bf16(wsc * acc[lt][2], wsc * acc[lt][3]);
      if (LAG) {   // bf16 Wg rows (unscaled) into the dead T tile: the A operand of R
        *reinterpret_cast<uint32_t*>(&s.qh[i0 * KH_LD + nt * 8 + 2 * tq]) = pack_bf16(acc[lt][0], acc[lt][1]);
        *reinterpret_cast<uint32_t*>(&s.qh[i1 * KH_LD + nt * 8 + 2 * tq]) = pack_bf16(acc[lt][2], acc[lt][3]);
      }
    }
    if (LAG) {
      // R = Wg Kt_{c-1}^T (64 x 64, contraction over the 128 key dims), negated, K-major no-swizzle image
      // [row i = token of this chunk][k = j = token of the previous chunk] behind the Au image
      __syncthreads();
      float cr[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) cr[i][e] = 0.f;
      if (has_prev) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          uint32_t aw[4];
          ldsm_x4(smem_u32(&s.qh[(r0 + (lane & 15)) * KH_LD + ks * 16 + (lane >> 4) * 8]), aw[0], aw[1], aw[2], aw[3]);
#pragma unroll
          for (int pp = 0; pp < 2; ++pp) {
            uint32_t b0, b1, b2, b3;
            const int n = (half * 2 + pp) * 16 + (lane & 7) + (lane >> 4) * 8, kk = ks * 16 + ((lane >> 3) & 1) * 8;
            ldsm_x4(smem_u32(&s.vb[n * V_LD + kk]), b0, b1, b2, b3);
            mma16816(cr[2 * pp], aw, b0, b1);
            mma16816(cr[2 * pp + 1], aw, b2, b3);
          }
        }
      }
#pragma unroll
      for (int lt = 0; lt < 4; ++lt) {
        const int nt = half * 4 + lt;
        uint8_t* rimg = ublob + AU_BYTES;
        *reinterpret_cast<uint32_t*>(rimg + sq_off<MODE>(i0, nt * 8 + 2 * tq)) = pack_bf16(-cr[lt][0], -cr[lt][1]);
        *reinterpret_cast<uint32_t*>(rimg + sq_off<MODE>(i1, nt * 8 + 2 * tq)) = pack_bf16(-cr[lt][2], -cr[lt][3]);
      }
    }
#pragma unroll
    for (int hv = 0; hv < (TR ? 0 : 2); ++hv) {  // two passes of 64 value columns, one per staged half of V
      const int vbase = hv * 128 + half * 64;
      const __nv_bfloat16* vsrc = hv == 0 ? s.vb : vb2;
      if (hv == 1) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (ks <= strip) {
          uint32_t a[4];
          ldsm_x4(smem_u32(&sAu[(r0 + (lane & 15)) * A_LD + ks * 16 + (lane >> 4) * 8]), a[0], a[1], a[2], a[3]);
          const int kk = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
          for (int ntp = 0; ntp < 4; ++ntp) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(smem_u32(&vsrc[kk * V_LD + half * 64 + ntp * 16 + (lane >> 4) * 8]), b0, b1, b2, b3);
            mma16816(acc[2 * ntp], a, b0, b1);
            mma16816(acc[2 * ntp + 1], a, b2, b3);
          }
        }
      }
#pragma unroll
      for (int lt = 0; lt < 8; ++lt) {
        const int colbase = vbase + lt * 8;  // first of the 8 value columns of this tile
        uint8_t* img = ublob + (colbase >> 5) * UBLOB_BYTES + ((colbase & 31) >> 3) * 1024 + tq * 4;
        *reinterpret_cast<uint32_t*>(img + i0 * 16) = pack_bf16(acc[lt][0], acc[lt][1]);
        *reinterpret_cast<uint32_t*>(img + i1 * 16) = pack_bf16(acc[lt][2], acc[lt][3]);
      }
    }
  }
  PTR(7);
  if (TC) {
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tc_tmem);
  }
  // publish the chunk: every image store of this CTA happens-before the flag (bar.sync, then a gpu-scope
  // release by the publishing thread -- the split-K semaphore pattern).  The consumer pairs it with an acquire
  // and a proxy fence before its bulk copies (async proxy) read the images.
  __syncthreads();
  if (tid == 0)
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ws.ready + ch), "r"(1u) : "memory");
#ifdef IVL_TRACE
  if (tid == 0 && h < 16 && c < 2048) ivl_prep_tl[(h * 2048 + c) * 4 + 2] = gtime();
#endif
}

}  // namespace

#ifdef IVL_TRACE
extern "C" __attribute__((visibility("default"))) int ivl_debug_read_prep_trace(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, ivl_prep_trace, sizeof(long long) * n);
}
extern "C" __attribute__((visibility("default"))) int ivl_debug_read_prep_tl(unsigned long long* host) {
  return (int)cudaMemcpyFromSymbol(host, ivl_prep_tl, sizeof(unsigned long long) * 16 * 2048 * 4);
}
extern "C" __attribute__((visibility("default"))) int ivl_debug_read_prep_wait(unsigned long long* host, int reset) {
  int e = (int)cudaMemcpyFromSymbol(host, ivl_prep_wait, sizeof(unsigned long long) * 4);
  if (reset) { unsigned long long z[4] = {0, 0, 0, 0}; cudaMemcpyToSymbol(ivl_prep_wait, z, sizeof(z)); }
  return e;
}
#endif

// Loads the kernel and raises its shared-memory limit (once per device).  The overlapped chunk operator calls
// this BEFORE it launches the scan: with lazy module loading the first launch of a kernel can block on the
// kernels already running, and a running scan is waiting for this very kernel.
cudaError_t configure_gdn_prep() {
  static std::atomic<bool> configured[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!configured[dev].load(std::memory_order_acquire)) {
    e = cudaFuncSetAttribute(gdn_prep_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PrepSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gdn_prep_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PrepSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gdn_prep_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PrepSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gdn_prep_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PrepSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gdn_prep_kernel<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(PrepSmem) + 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gdn_prep_kernel<1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(PrepSmem) + 1024);
    if (e != cudaSuccess) return e;
    configured[dev].store(true, std::memory_order_release);
  }
  return cudaSuccess;
}

// scan_ctas_per_head: number of scan CTAs per head whose progress counters gate the ring (ignored when the
// ring holds every chunk)
// vl.chunk_tok0 != nullptr: packed variable-length batch (B must be 1) of num_chunks chunks
cudaError_t launch_gdn_prep(const void* q, const void* k, const void* v, const float* g, const void* beta,
                            const GdnWorkspace& ws, const GdnVarlen& vl, int num_chunks, int B, int T, int H,
                            float scale, int l2norm, int scan_ctas_per_head, int transposed, cudaStream_t stream,
                            const GdnPrepFused* fused) {
  const int smem = (int)sizeof(PrepSmem);
  if (cudaError_t e = configure_gdn_prep()) return e;
  static int resident = 0;
  if (resident == 0) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    resident = 3 * sms;
  }
  dim3 grid((unsigned)num_chunks * (unsigned)H, 1, B);
  const bool fz = fused != nullptr && fused->wq != nullptr;
  if (fz && (transposed != 1 || vl.chunk_tok0 != nullptr)) return cudaErrorInvalidValue;   // transposed scan, dense only
  // IVL_GDN_PREP_TC (developer knob): 1 = the large products on tcgen05 (default for the transposed scan's images),
  // 0 = warp-level mma.sync everywhere
  const char* tc_env = getenv("IVL_GDN_PREP_TC");
  const bool tc = transposed == 1 && ((tc_env && *tc_env) ? atoi(tc_env) : PREP_TC_DEFAULT) != 0;
  auto kern = fz ? (tc ? gdn_prep_kernel<1, true, true> : gdn_prep_kernel<1, true>)
                 : (tc ? gdn_prep_kernel<1, false, true>
                       : (transposed == 2 ? gdn_prep_kernel<2> : (transposed == 1 ? gdn_prep_kernel<1> : gdn_prep_kernel<0>)));
  kern<<<grid, PREP_THREADS, tc ? smem + 1024 : smem, stream>>>(
      static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
      static_cast<const __nv_bfloat16*>(v), g, static_cast<const __nv_bfloat16*>(beta), ws, vl, T, H, scale, l2norm,
      resident, scan_ctas_per_head, fz ? *fused : GdnPrepFused{});
  return cudaGetLastError();
}

}  // namespace ivl
