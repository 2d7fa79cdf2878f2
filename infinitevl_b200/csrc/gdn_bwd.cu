// Gated DeltaNet backward: exact gradient of the token recurrence (the definition of the operator,
// src/llamafactory/model/fla/ops/gated_delta_rule/fused_recurrent.py:85-108), in fp32.
//
// Replaces chunk_gated_delta_rule_bwd of the reference (fla/ops/gated_delta_rule/chunk.py:74-177: recompute of w, u,
// h; chunk_bwd_dv_local, chunk_gated_delta_rule_bwd_dhu, chunk_bwd_dqkwg, the WY backward of wy_fast.py:432-620) for
// the training drop-in.  The reference differentiates its chunked bf16 algorithm; this
// kernel differentiates the recurrence those kernels approximate, with recomputation instead of stored states:
//
//   forward step   St~ = a_t S_{t-1};  r = St~^T k;  u = b_t (v - r);  S_t = St~ + k u^T;  o = scale S_t^T q
//   backward step  dS += scale q do^T;           dq  = scale S_t do
//                  du  = dS^T k;                  dk  = dS u
//                  db  = du . (v - r);            dv  = b_t du;      dr = -b_t du
//                  dSt~ = dS + k dr^T;            dk += St~ dr
//                  da  = <dSt~, S_{t-1}>;         dg  = da a_t;      dS <- a_t dSt~
// (q, k are the L2-normalised rows; the chain rule through the normalisation is element-wise and done by the
// caller.)  One CTA owns 16 value columns of one (batch, head) -- the recurrence is independent per value column --
// as 8 warps x 2 columns, a lane holding 4 of the 128 key rows of both columns: 8 state entries per thread.
// Pass 1 runs the recurrence forward and stores the state every 16 tokens (workspace: T/16 x 8 KiB per CTA); pass 2
// walks the 16-token blocks backwards: reload the block's start state, replay the block keeping the 16 intermediate
// states IN REGISTERS (128 per thread), then run the 16 backward steps.  Sums over key rows are warp shuffles; sums
// over value columns (dq, dk, db, dg: one contribution per warp) go through shared-memory accumulators per block
// and one global atomicAdd per element and block.
#include <atomic>

#include "gdn_layout.cuh"
#include "sm100.cuh"

namespace ivl {

namespace {

constexpr int BW_CH = 16;     // tokens per recompute block
constexpr int BW_COLS = 16;   // value columns per CTA
constexpr int BW_THREADS = 256;

struct BwSmem {
  float qn[BW_CH][GDN_K], kn[BW_CH][GDN_K];
  float dq[BW_CH][GDN_K], dk[BW_CH][GDN_K];
  float v[BW_CH][BW_COLS], dout[BW_CH][BW_COLS], u[BW_CH][BW_COLS], vr[BW_CH][BW_COLS];
  float alpha[BW_CH], beta[BW_CH], dg[BW_CH], db[BW_CH];
};

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
  return x;
}

__global__ void __launch_bounds__(BW_THREADS, 1)
gdn_bwd_kernel(const float* __restrict__ qn, const float* __restrict__ kn, const __nv_bfloat16* __restrict__ v,
               const float* __restrict__ g, const float* __restrict__ beta, const __nv_bfloat16* __restrict__ dout,
               const float* __restrict__ h0, const float* __restrict__ dht, float* __restrict__ dqn,
               float* __restrict__ dkn, float* __restrict__ dv, float* __restrict__ dg, float* __restrict__ dbeta,
               float* __restrict__ dh0, float* __restrict__ ckpt, int T, int H, float scale) {
  __shared__ BwSmem s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int NC = (T + BW_CH - 1) / BW_CH;
  const int col0 = slice * BW_COLS + 2 * warp;              // this warp's two value columns (global index)
  const int lc = 2 * warp;                                  // ... and their index inside the CTA's slice
  const int r0 = 4 * lane;                                  // this lane's four key rows
  const size_t cta = ((size_t)b * H + h) * (GDN_V / BW_COLS) + slice;
  float* ck = ckpt + cta * (size_t)NC * (BW_THREADS * 8);
  const size_t state_off = ((size_t)b * H + h) * GDN_K * GDN_V;
  float S[2][4];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int r = 0; r < 4; ++r) S[c][r] = h0 ? h0[state_off + (size_t)(r0 + r) * GDN_V + col0 + c] : 0.f;

  // operands of one 16-token block -> shared memory (token rows past T are never touched)
  auto stage = [&](int blk, bool with_q) {
    const int t0 = blk * BW_CH, n = min(BW_CH, T - t0);
    for (int i = tid; i < n * (GDN_K / 4); i += BW_THREADS) {
      const int tk = i / (GDN_K / 4), c4 = i % (GDN_K / 4);
      const size_t off = (((size_t)b * T + t0 + tk) * H + h) * GDN_K + c4 * 4;
      *reinterpret_cast<float4*>(&s.kn[tk][c4 * 4]) = __ldg(reinterpret_cast<const float4*>(kn + off));
      if (with_q) *reinterpret_cast<float4*>(&s.qn[tk][c4 * 4]) = __ldg(reinterpret_cast<const float4*>(qn + off));
    }
    for (int i = tid; i < n * BW_COLS; i += BW_THREADS) {
      const int tk = i / BW_COLS, c = i % BW_COLS;
      const size_t off = (((size_t)b * T + t0 + tk) * H + h) * GDN_V + slice * BW_COLS + c;
      s.v[tk][c] = __bfloat162float(v[off]);
      if (with_q) s.dout[tk][c] = __bfloat162float(dout[off]);
    }
    if (tid < n) {
      const size_t off = ((size_t)b * T + t0 + tid) * H + h;
      s.alpha[tid] = __expf(g[off]);
      s.beta[tid] = beta[off];
    }
    return n;
  };
  // one forward step on the registers; returns (u, v - r) of this warp's two columns
  auto fwd_step = [&](int i, float (&u)[2], float (&vr)[2]) {
    const float a = s.alpha[i], bt = s.beta[i];
    const float4 k4 = *reinterpret_cast<const float4*>(&s.kn[i][r0]);
    const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float r = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) { S[c][q] *= a; r = fmaf(S[c][q], kk[q], r); }
      r = warp_sum(r);
      vr[c] = s.v[i][lc + c] - r;
      u[c] = bt * vr[c];
#pragma unroll
      for (int q = 0; q < 4; ++q) S[c][q] = fmaf(kk[q], u[c], S[c][q]);
    }
  };

  // ---------------- pass 1: forward, state checkpoints every 16 tokens ----------------
  for (int blk = 0; blk < NC; ++blk) {
    float4* dst = reinterpret_cast<float4*>(ck + ((size_t)blk * BW_THREADS + tid) * 8);
    dst[0] = make_float4(S[0][0], S[0][1], S[0][2], S[0][3]);
    dst[1] = make_float4(S[1][0], S[1][1], S[1][2], S[1][3]);
    const int n = stage(blk, false);
    __syncthreads();
    for (int i = 0; i < n; ++i) {
      float u[2], vr[2];
      fwd_step(i, u, vr);
    }
    __syncthreads();
  }

  // ---------------- pass 2: backward over the blocks ----------------
  float dS[2][4];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int r = 0; r < 4; ++r) dS[c][r] = dht ? dht[state_off + (size_t)(r0 + r) * GDN_V + col0 + c] : 0.f;
  for (int blk = NC - 1; blk >= 0; --blk) {
    const float4* src = reinterpret_cast<const float4*>(ck + ((size_t)blk * BW_THREADS + tid) * 8);
    const float4 a0 = src[0], a1 = src[1];
    S[0][0] = a0.x; S[0][1] = a0.y; S[0][2] = a0.z; S[0][3] = a0.w;
    S[1][0] = a1.x; S[1][1] = a1.y; S[1][2] = a1.z; S[1][3] = a1.w;
    const int n = stage(blk, true);
    for (int i = tid; i < BW_CH * GDN_K; i += BW_THREADS) { (&s.dq[0][0])[i] = 0.f; (&s.dk[0][0])[i] = 0.f; }
    if (tid < BW_CH) { s.dg[tid] = 0.f; s.db[tid] = 0.f; }
    __syncthreads();
    // replay the block, keeping the state BEFORE every step in registers
    float hist[BW_CH][2][4];
#pragma unroll
    for (int i = 0; i < BW_CH; ++i) {
      if (i < n) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int q = 0; q < 4; ++q) hist[i][c][q] = S[c][q];
        float u[2], vr[2];
        fwd_step(i, u, vr);
        if (lane == 0) {
          s.u[i][lc] = u[0]; s.u[i][lc + 1] = u[1];
          s.vr[i][lc] = vr[0]; s.vr[i][lc + 1] = vr[1];
        }
      }
    }
    __syncwarp();
    const int t0 = blk * BW_CH;
#pragma unroll
    for (int i = BW_CH - 1; i >= 0; --i) {
      if (i < n) {
        const float a = s.alpha[i], bt = s.beta[i];
        const float4 k4 = *reinterpret_cast<const float4*>(&s.kn[i][r0]);
        const float4 q4 = *reinterpret_cast<const float4*>(&s.qn[i][r0]);
        const float kk[4] = {k4.x, k4.y, k4.z, k4.w}, qq[4] = {q4.x, q4.y, q4.z, q4.w};
        float dqa[4] = {0.f, 0.f, 0.f, 0.f}, dka[4] = {0.f, 0.f, 0.f, 0.f};
        float dba = 0.f, dga = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float uc = s.u[i][lc + c], vrc = s.vr[i][lc + c], doc = s.dout[i][lc + c] * scale;
          float du = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float st = a * hist[i][c][q];            // St~
            dqa[q] = fmaf(fmaf(kk[q], uc, st), doc, dqa[q]);   // S_t = St~ + k u
            dS[c][q] = fmaf(qq[q], doc, dS[c][q]);
            du = fmaf(dS[c][q], kk[q], du);
            dka[q] = fmaf(dS[c][q], uc, dka[q]);
          }
          du = warp_sum(du);
          dba = fmaf(du, vrc, dba);
          const float dr = -bt * du;
          if (lane == 0) dv[(((size_t)b * T + t0 + i) * H + h) * GDN_V + col0 + c] = bt * du;
          float da = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float st = a * hist[i][c][q];
            dka[q] = fmaf(st, dr, dka[q]);
            const float dst = fmaf(kk[q], dr, dS[c][q]);   // dSt~
            da = fmaf(dst, hist[i][c][q], da);
            dS[c][q] = a * dst;
          }
          dga += da;
        }
        dga = warp_sum(dga) * a;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          atomicAdd(&s.dq[i][r0 + q], dqa[q]);
          atomicAdd(&s.dk[i][r0 + q], dka[q]);
        }
        if (lane == 0) { atomicAdd(&s.dg[i], dga); atomicAdd(&s.db[i], dba); }
      }
    }
    __syncthreads();
    for (int i = tid; i < n * GDN_K; i += BW_THREADS) {
      const int tk = i / GDN_K, r = i % GDN_K;
      const size_t off = (((size_t)b * T + t0 + tk) * H + h) * GDN_K + r;
      atomicAdd(dqn + off, s.dq[tk][r]);
      atomicAdd(dkn + off, s.dk[tk][r]);
    }
    if (tid < n) {
      const size_t off = ((size_t)b * T + t0 + tid) * H + h;
      atomicAdd(dg + off, s.dg[tid]);
      atomicAdd(dbeta + off, s.db[tid]);
    }
    __syncthreads();
  }
  if (dh0) {
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) dh0[state_off + (size_t)(r0 + r) * GDN_V + col0 + c] = dS[c][r];
  }
}

}  // namespace

size_t gdn_bwd_workspace_bytes(int B, int T, int H) {
  const size_t NC = (T + BW_CH - 1) / BW_CH;
  return (size_t)B * H * (GDN_V / BW_COLS) * NC * BW_THREADS * 8 * sizeof(float);
}

// dqn, dkn, dg, dbeta are ACCUMULATED into (atomics): the caller zeroes them.  dv, dh0 are overwritten.
cudaError_t launch_gdn_bwd(const float* qn, const float* kn, const void* v, const float* g, const float* beta,
                           const void* dout, const float* h0, const float* dht, float* dqn, float* dkn, float* dv,
                           float* dg, float* dbeta, float* dh0, float* workspace, int B, int T, int H, float scale,
                           cudaStream_t stream) {
  dim3 grid(GDN_V / BW_COLS, H, B);
  gdn_bwd_kernel<<<grid, BW_THREADS, 0, stream>>>(qn, kn, static_cast<const __nv_bfloat16*>(v), g, beta,
                                                  static_cast<const __nv_bfloat16*>(dout), h0, dht, dqn, dkn, dv, dg,
                                                  dbeta, dh0, workspace, T, H, scale);
  return cudaGetLastError();
}

}  // namespace ivl
