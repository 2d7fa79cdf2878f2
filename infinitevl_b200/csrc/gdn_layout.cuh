// Workspace layout shared by gdn_prep.cu (producer) and gdn_scan.cu (consumer).
//
// The prep kernel writes, for every (batch b, head h, chunk c of 64 tokens), the
// tensor-core operand IMAGES the scan kernel needs, already in the exact byte
// order tcgen05.mma expects in shared memory (no-swizzle core-matrix tiling, see
// sm100.cuh).  The scan kernel then only issues 1-D bulk copies (TMA engine)
// into its pipeline stages -- no layout work on the serial path.
//
//   blob[b][h][slot] (BLOB_BYTES = 57 KiB, contiguous; slot = c mod ring):
//     A1  image 32 KiB   K-major operand [-Wg ; Qg]  (128 rows x 128 k)
//                        Wg = A (beta Kn exp(G)),  Qg = Qn exp(G) scale
//     P   image  8 KiB   rows 64..127 of the K-major operand [0 ; P],  P = tril(Qn Kn^T * Gamma) * scale
//     Kt  image 16 KiB   MN-major operand Kt^T (M = 128 key dims, K = 64 tokens), Kt = Kn exp(G_C - G)
//     tail        1 KiB   fp32 gamma = exp(G_C), the whole-chunk decay, in the first 4 bytes (128 B are copied)
//                        (P, Kt and the tail are adjacent: the scan fetches them with one copy, A1 with
//                         another, because their shared-memory slots are recycled at different times;
//                         gamma travels with the operands so that the scan never reads it before the
//                         chunk is published)
//   ublob[b][h][slot][s] (4 KiB each, s = V slice of 32 columns): U = A (beta V), laid out
//                        [piece p of 8 columns][token][8] so that a thread owning a token
//                        row reads four conflict-free 16-byte pieces.
//   Transposed scan (gdn_scan_t.cu): the first 8 KiB of the chunk's ublob region hold the Au image instead
//                        (K-major no-swizzle [64][64], Au = T diag(beta)); U = Au V is formed by the scan itself.
//                        Lag form: the next 8 KiB hold -R_c = -Wg_c Kt_{c-1}^T (same layout; k = token of chunk c-1)
//                        and rows 0..63 of the A1 image hold -gamma_{c-1} Wg_c.
//   ready[b][h][c]      uint32 flag, zeroed by the host entry point before the launch and set to 1
//                       (release, gpu scope) by the prep CTA once every image of the chunk is written.
//                       The scan's copy warp polls it (acquire), so the two kernels can run
//                       concurrently: prep walks the sequence front to back over all heads, the scan
//                       follows a few chunks behind and finds the images still in L2.
//   progress[b][h][8]   uint32, zeroed with the flags: number of chunks scan CTA s of the head has fully
//                       consumed.  In the overlapped form the images live in a RING of `ring` chunk slots
//                       per head: the prep CTA of chunk c waits until every scan CTA of its head has
//                       consumed chunk c - ring before it overwrites the slot.  A ring of a few dozen
//                       chunks stays resident in the 126 MB L2, so the images never travel to HBM.
//                       Back to back (ring = number of chunks) nothing waits.
//
// K-major no-swizzle image of an [R rows][Kd] bf16 matrix: 8x8 "core matrices" of
// 128 contiguous bytes (8 rows x 16 B); core (rg, kg) at rg * (Kd/8)*128 + kg * 128.
//   -> descriptor LBO = 128 (k-adjacent), SBO = (Kd/8)*128 (row-group-adjacent);
//      one MMA (16 k) advances the start address by 256 B.
// MN-major no-swizzle image of an operand with MN extent Mn and K extent Kd:
//   core (mg, kg) = 8 k-rows x 8 mn elements at mg * (Kd/8)*128 + kg * 128.
//   -> LBO = 128 (k-adjacent), SBO = (Kd/8)*128 (mn-adjacent); one MMA advances 256 B.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace ivl {

constexpr int GDN_C = 64;    // chunk length (fla/ops/gated_delta_rule/chunk.py:199 hard-codes 64)
constexpr int GDN_K = 128;   // key/query head dim
constexpr int GDN_V = 256;   // value head dim (expand_v = 2)
constexpr int GDN_BV = 32;   // value columns per U sub-slice (a scan CTA owns 1, 2 or 4 adjacent sub-slices)
constexpr int GDN_NS = GDN_V / GDN_BV;  // U sub-slices per head

constexpr uint32_t P_BYTES = 64 * 64 * 2;        // 8 KiB
constexpr uint32_t A1_BYTES = 128 * 128 * 2;     // 32 KiB
constexpr uint32_t KT_BYTES = 128 * 64 * 2;      // 16 KiB
constexpr uint32_t TAIL_BYTES = 128;             // gamma (fp32) + padding: what the scan copies
constexpr uint32_t TAIL_STRIDE = 1024;           // what the blob reserves, so that blobs stay 1 KiB aligned
constexpr uint32_t BLOB_BYTES = P_BYTES + A1_BYTES + KT_BYTES + TAIL_STRIDE;  // 57 KiB
constexpr uint32_t UBLOB_BYTES = 64 * GDN_BV * 2;               // 4 KiB
constexpr uint32_t AU_BYTES = 64 * 64 * 2;                      // 8 KiB: Au = T diag(beta) image (transposed scan)

constexpr uint32_t BLOB_OFF_A1 = 0;
constexpr uint32_t BLOB_OFF_P = A1_BYTES;
constexpr uint32_t BLOB_OFF_KT = A1_BYTES + P_BYTES;
constexpr uint32_t BLOB_OFF_TAIL = A1_BYTES + P_BYTES + KT_BYTES;

// Packed variable-length batch (cu_seqlens of the reference operator): the flattened token axis is cut into
// chunks per SEQUENCE (a chunk never straddles two sequences; the last chunk of a sequence may be short).
// All three tables live in device memory; null pointers mean the dense [B, T] case.
struct GdnVarlen {
  const int* chunk_tok0;       // [num_chunks]  first token of the chunk on the flattened axis
  const int* chunk_valid;      // [num_chunks]  tokens in the chunk (1..64)
  const int* seq_chunk_begin;  // [N + 1]       chunks of sequence n are seq_chunk_begin[n] .. seq_chunk_begin[n+1]-1
  int num_seqs;                // N
};

// Prefill-side fusion: when `wq` is set, gdn_prep_kernel reads the RAW q / k projection
// outputs and applies the depthwise causal conv (kernel 4) + SiLU itself, and derives g / beta from the raw a / b
// projections -- the three launches in front of the chunk operator (ivl_short_conv_fwd x2 on q and k, ivl_gdn_gate_fwd)
// and their 2.2 GB of HBM round trip per 128K-token layer disappear.  All pointers are device pointers.
struct GdnPrepFused {
  const void* wq = nullptr;        // bf16 [H*128][4] conv weights of q (null: not fused)
  const void* wk = nullptr;        // bf16 [H*128][4]
  const void* cq_in = nullptr;     // bf16 [B][H*128][4] carried conv tails (newest input in column 3) or null
  const void* ck_in = nullptr;
  void* cq_out = nullptr;          // bf16 [B][H*128][4] tails after this call, or null
  void* ck_out = nullptr;
  const void* a = nullptr;         // bf16 [B][T][H] raw a_proj / b_proj outputs (replace g / beta)
  const void* b = nullptr;
  const float* A_log = nullptr;    // fp32 [H]
  const float* dt_bias = nullptr;  // fp32 [H]
};

struct GdnWorkspace {
  uint8_t* blob;       // [B][H][ring][BLOB_BYTES]
  uint8_t* ublob;      // [B][H][ring][NS][UBLOB_BYTES]
  uint32_t* ready;     // [B][H][NT]
  uint32_t* progress;  // [B][H][GDN_NS]
  uint32_t* checkin;   // one word: scan CTAs that are resident (zeroed with the flags; gates prep's launch when a ring is used)
  int ring;            // chunk slots per head (<= NT)
};

__host__ __device__ inline int gdn_num_chunks(int T) { return (T + GDN_C - 1) / GDN_C; }

// flags + progress counters (one memset), rounded to 1 KiB; they sit at the FRONT of the workspace so that
// their address does not depend on the ring length
__host__ inline size_t gdn_sync_bytes(int B, int T, int H) {
  size_t n = ((size_t)B * H * (gdn_num_chunks(T) + GDN_NS) + 1) * sizeof(uint32_t);
  return (n + 1023) / 1024 * 1024;
}

// Sized for the back-to-back form (one slot per chunk); the overlapped form uses a prefix of it.
__host__ inline size_t gdn_workspace_bytes(int B, int T, int H) {
  size_t n = (size_t)B * H * gdn_num_chunks(T);
  return gdn_sync_bytes(B, T, H) + n * (BLOB_BYTES + GDN_NS * UBLOB_BYTES);
}

// ring: chunk slots per head, clamped to [1, NT]
__host__ inline GdnWorkspace gdn_carve(void* ws, int B, int T, int H, int ring) {
  const int NT = gdn_num_chunks(T);
  if (ring <= 0 || ring > NT) ring = NT;
  size_t n = (size_t)B * H * ring;
  GdnWorkspace w;
  w.ready = static_cast<uint32_t*>(ws);
  w.progress = w.ready + (size_t)B * H * NT;
  w.checkin = w.progress + (size_t)B * H * GDN_NS;
  w.blob = static_cast<uint8_t*>(ws) + gdn_sync_bytes(B, T, H);
  w.ublob = w.blob + n * BLOB_BYTES;
  w.ring = ring;
  return w;
}

}  // namespace ivl
