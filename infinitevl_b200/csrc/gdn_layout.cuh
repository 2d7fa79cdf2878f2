// Workspace layout shared by gdn_prep.cu (producer) and gdn_scan.cu (consumer).
//
// The prep kernel writes, for every (batch b, head h, chunk c of 64 tokens), the
// tensor-core operand IMAGES the scan kernel needs, already in the exact byte
// order tcgen05.mma expects in shared memory (no-swizzle core-matrix tiling, see
// sm100.cuh).  The scan kernel then only issues 1-D bulk copies (TMA engine)
// into its pipeline stages -- no layout work on the serial path.
//
//   blob[b][h][c] (BLOB_BYTES = 56 KiB, contiguous):
//     A1  image 32 KiB   K-major operand [-Wg ; Qg]  (128 rows x 128 k)
//                        Wg = A (beta Kn exp(G)),  Qg = Qn exp(G) scale
//     P   image  8 KiB   rows 64..127 of the K-major operand [0 ; P],  P = tril(Qn Kn^T * Gamma) * scale
//     Kt  image 16 KiB   MN-major operand Kt^T (M = 128 key dims, K = 64 tokens), Kt = Kn exp(G_C - G)
//                        (P and Kt are adjacent: the scan fetches them with one copy, A1 with another,
//                         because their shared-memory slots are recycled at different times)
//   ublob[b][h][c][s] (4 KiB each, s = V slice of 32 columns): U = A (beta V), laid out
//                        [piece p of 8 columns][token][8] so that a thread owning a token
//                        row reads four conflict-free 16-byte pieces.
//   gamma[b][h][c]      fp32 exp(G_C), the whole-chunk decay.
//   ready[b][h][c]      uint32 flag, zeroed by the host entry point before the launch and set to 1
//                       (release, gpu scope) by the prep CTA once every image of the chunk is written.
//                       The scan's copy warp polls it (acquire), so the two kernels can run
//                       concurrently: prep walks the sequence front to back over all heads, the scan
//                       follows a few chunks behind and finds the images still in L2.
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
constexpr uint32_t BLOB_BYTES = P_BYTES + A1_BYTES + KT_BYTES;  // 56 KiB
constexpr uint32_t UBLOB_BYTES = 64 * GDN_BV * 2;               // 4 KiB

constexpr uint32_t BLOB_OFF_A1 = 0;
constexpr uint32_t BLOB_OFF_P = A1_BYTES;
constexpr uint32_t BLOB_OFF_KT = A1_BYTES + P_BYTES;

struct GdnWorkspace {
  uint8_t* blob;    // [B][H][NT][BLOB_BYTES]
  uint8_t* ublob;   // [B][H][NT][NS][UBLOB_BYTES]
  float* gamma;     // [B][H][NT]
  uint32_t* ready;  // [B][H][NT]
};

__host__ __device__ inline int gdn_num_chunks(int T) { return (T + GDN_C - 1) / GDN_C; }

__host__ inline size_t gdn_workspace_bytes(int B, int T, int H) {
  size_t n = (size_t)B * H * gdn_num_chunks(T);
  size_t gamma = (n * sizeof(float) + 1023) / 1024 * 1024;
  return n * BLOB_BYTES + n * GDN_NS * UBLOB_BYTES + 2 * gamma;
}

__host__ inline size_t gdn_ready_bytes(int B, int T, int H) {
  return (size_t)B * H * gdn_num_chunks(T) * sizeof(uint32_t);
}

__host__ inline GdnWorkspace gdn_carve(void* ws, int B, int T, int H) {
  size_t n = (size_t)B * H * gdn_num_chunks(T);
  GdnWorkspace w;
  w.blob = static_cast<uint8_t*>(ws);
  w.ublob = w.blob + n * BLOB_BYTES;
  w.gamma = reinterpret_cast<float*>(w.ublob + n * GDN_NS * UBLOB_BYTES);
  w.ready = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(w.gamma) + (n * sizeof(float) + 1023) / 1024 * 1024);
  return w;
}

}  // namespace ivl
