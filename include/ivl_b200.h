/*
 * ivl_b200.h -- C ABI of the B200-native InfiniteVL hybrid-attention hot path.
 *
 * One shared library (libivl_b200.so), plain pointers and sizes, no C++ or
 * torch types.  Every entry point
 *   - runs asynchronously on the CUDA stream passed in (a cudaStream_t cast to void*),
 *   - never allocates device memory, never synchronises, never throws (ivl_gdn_chunk_fwd creates one helper
 *     stream and two events per caller stream the first time it overlaps its kernels -- or ivl_stream_init does,
 *     ahead of time -- and nothing afterwards; ivl_stream_release frees them; outside of stream capture the
 *     overlapped form also issues one stream memory operation, cuStreamWaitValue32, on that helper stream; the
 *     ivl_ipc_* functions of the multi-GPU hand-off map / unmap peer memory and are set-up calls, not stream work),
 *   - returns IVL_ERR_ARCH on a device that is not sm_100 (the kernels exist for sm_100a only),
 *   - returns IVL_OK or a negative IVL_ERR_* code (ivl_strerror() names it),
 * so it is safe inside CUDA-graph capture (the reference demo captures the whole
 * forward, inference_examples/demo_streaming_inference.py:473-486).
 *
 * Each function names the reference interface it replaces.  Paths are relative to
 * the reference checkout; "fla/" = src/llamafactory/model/fla/, "std" =
 * infinitevl/infinitevl_standard/modeling_infinitevl.py.
 *
 * Tensor layouts are the reference's time-first ones: q,k [B,T,H,K], v,o [B,T,H,V],
 * g,beta [B,T,H], state [B,H,K,V]; all dense/contiguous.  bf16 is passed as
 * uint16_t-sized storage (const void*).
 */
#ifndef IVL_B200_H_
#define IVL_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(IVL_BUILDING_DLL)
#define IVL_API __attribute__((visibility("default")))
#else
#define IVL_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define IVL_OK 0
#define IVL_ERR_BAD_SHAPE (-1)     /* unsupported head dims / sizes                     */
#define IVL_ERR_NULL (-2)          /* required pointer is NULL                           */
#define IVL_ERR_WORKSPACE (-3)     /* workspace too small                                */
#define IVL_ERR_DTYPE (-4)         /* unknown dtype code                                 */
#define IVL_ERR_LAUNCH (-5)        /* CUDA launch failed (cudaGetLastError was non-zero) */
#define IVL_ERR_ARCH (-6)          /* device is not sm_100                               */

#define IVL_DTYPE_F32 0
#define IVL_DTYPE_BF16 1

/* Library identification (3: ivl_swa_fwd_pos / _varlen, fused and packed entry points, peer-memory hand-off); also the
 * cheapest "does the .so load" check. */
IVL_API int ivl_abi_version(void);
IVL_API const char* ivl_strerror(int code);
/* After IVL_ERR_LAUNCH: the CUDA runtime call that failed and its error text (per calling thread). */
IVL_API const char* ivl_last_cuda_error(void);

/* Optional: create, ahead of time, the helper stream and the two events ivl_gdn_chunk_fwd uses when it overlaps its
 * two kernels for calls issued on `stream` of the current device (otherwise created lazily by the first such call;
 * both are legal during stream capture).  ivl_stream_release destroys them again: call it before destroying a
 * stream that has been used with this library (the set is keyed by the stream handle), with no call in flight.
 * Replaces nothing in the reference (its Triton launches own no streams); part of the drop-in's resource contract. */
IVL_API int ivl_stream_init(void* stream);
IVL_API int ivl_stream_release(void* stream);

/* Make `stream` wait until the 32-bit word at device address `addr` (4-byte aligned, device memory of the current
 * device) is >= `value` (cuStreamWaitValue32, wrap-around compare; no kernel, no SM).  The consumer half of the
 * peer-memory neighbour hand-off of the sequence-sharded prefill (infinitevl_b200/dist.py PeerLink): the previous
 * rank copies the DeltaNet state / SWA halo straight into this rank's buffers over NVLink and then writes the flag.
 * Replaces nothing in the reference (it never shards a sequence). */
IVL_API int ivl_stream_wait_value32(void* stream, const void* addr, uint32_t value);

/* The producer half: copy `bytes` (multiple of 16; both pointers 16-byte aligned) from `src` (this device) to `dst`
 * -- memory of ANOTHER GPU of the node, mapped into this process (CUDA IPC) -- with ordinary stores over NVLink, then
 * write `value` to `*flag` (also peer memory) with a system-scope release, in one launch on `stream`.  bytes == 0
 * writes the flag only.  `counter`: one zero-initialised 32-bit word in THIS device's memory per concurrently
 * running put (the kernel leaves it at zero). */
IVL_API int ivl_peer_put(void* dst, const void* src, size_t bytes, uint32_t* flag, uint32_t value, uint32_t* counter,
                         void* stream);
/* Map a device allocation exported by another process of the node (the 64-byte cudaIpcMemHandle_t of its BASE
 * allocation) for the CURRENT device, peer access enabled -- the `dst` / `flag` addresses of ivl_peer_put.  One open
 * per handle and process; ivl_ipc_close unmaps. */
/* Export side: the handle of the cudaMalloc allocation that contains `ptr` and ptr's byte offset inside it. */
IVL_API int ivl_ipc_export(const void* ptr, void* handle64_out, uint64_t* offset_out);
IVL_API int ivl_ipc_open(const void* handle64, void** base_ptr);
IVL_API int ivl_ipc_close(void* base_ptr);

/* ------------------------------------------------------------------------------------
 * Gated DeltaNet, chunked prefill (T > 64 in the model, any T >= 1 here).
 * Replaces chunk_gated_delta_rule(q,k,v,g,beta,scale,initial_state,output_final_state,
 * cu_seqlens,use_qk_l2norm_in_kernel) -- fla/ops/gated_delta_rule/chunk.py:273-392,
 * called from std:1298-1308 and fla/layers/gated_deltanet.py:275-286.
 *
 *   q,k   bf16 [B,T,H,128]      v bf16 [B,T,H,256]     g fp32 [B,T,H] (log decay)
 *   beta  bf16 [B,T,H]          o bf16 [B,T,H,256]
 *   h0    initial state [B,H,128,256] in h0_dtype, or NULL (zeros)
 *   ht    final state   [B,H,128,256] in ht_dtype, or NULL (not written)
 *   scale <= 0 selects K^-0.5;  l2norm_qk != 0 normalises q,k rows in-kernel (eps 1e-6).
 *   workspace: ivl_gdn_chunk_workspace_bytes(B,T,H) bytes, 1024-byte aligned.
 * Only H*K*V = (any H)*128*256 is supported (the InfiniteVL shape); else IVL_ERR_BAD_SHAPE.
 *
 * Execution: two kernels, gdn_prep_kernel (parallel over chunks) and gdn_scan_kernel (the
 * recurrence).  For T >= 2048 ivl_gdn_chunk_fwd runs them OVERLAPPED: the scan is launched on
 * `stream`, prep on a library-owned second stream forked from and joined back to `stream` with
 * events (legal inside a stream capture, so the call stays CUDA-graph safe); the scan follows
 * per-chunk ready flags in the workspace.  Work submitted to `stream` after the call is ordered
 * after both kernels.  Shorter inputs run prep then scan on `stream`.
 * ---------------------------------------------------------------------------------- */
IVL_API size_t ivl_gdn_chunk_workspace_bytes(int B, int T, int H);

IVL_API int ivl_gdn_chunk_fwd(const void* q, const void* k, const void* v, const float* g, const void* beta,
                      const void* h0, int h0_dtype, void* o, void* ht, int ht_dtype, int B, int T,
                      int H, int K, int V, float scale, int l2norm_qk, void* workspace,
                      size_t workspace_bytes, void* stream);

/* Prefill-side fusion (replaces the two ShortConvolution calls on q and k and the gate
 * math in front of the operator, std:1263-1294): the chunk operator on the RAW projection outputs.
 *   xq, xk  bf16 [B,T,H*128]  q_proj / k_proj outputs BEFORE the short conv (conv + SiLU + L2 norm happen in the pre-pass)
 *   v       bf16 [B,T,H,256]  values AFTER their short conv (the scan reads them with TMA)
 *   a, b    bf16 [B,T,H]      a_proj / b_proj outputs: g = -exp(A_log) softplus(a + dt_bias), beta = sigmoid(b)
 *   conv_wq/wk bf16 [H*128,4]; conv_q_in/k_in bf16 [B,H*128,4] carried tails or NULL; conv_q_out/k_out: tails after
 *   the call or NULL (both or neither; must not alias the inputs); A_log, dt_bias fp32 [H].
 * Rounding points are those of the unfused chain (ivl_short_conv_fwd, ivl_gdn_gate_fwd, ivl_gdn_chunk_fwd with
 * l2norm_qk = 1), so the results are bit-identical to it.  Dense batches only. */
IVL_API int ivl_gdn_chunk_fwd_fused(const void* xq, const void* xk, const void* v, const void* a, const void* b,
                                    const void* conv_wq, const void* conv_wk, const void* conv_q_in,
                                    const void* conv_k_in, void* conv_q_out, void* conv_k_out, const float* A_log,
                                    const float* dt_bias, const void* h0, int h0_dtype, void* o, void* ht, int ht_dtype,
                                    int B, int T, int H, int K, int V, float scale, void* workspace,
                                    size_t workspace_bytes, void* stream);

/* Packed variable-length batch: the reference's `cu_seqlens` form (fla/ops/gated_delta_rule/chunk.py:211-214,
 * 355-369: B = 1, sequences concatenated on the token axis, one initial / final state per sequence).  One launch
 * pair for the whole batch.  The caller cuts every sequence into chunks of 64 tokens (the last one may be
 * short; empty sequences have no chunk) and passes three DEVICE arrays:
 *   chunk_tok0[num_chunks], chunk_valid[num_chunks]   first token / token count of each chunk,
 *   seq_chunk_begin[num_seqs + 1]                      chunk range of each sequence.
 *   h0, ht: [num_seqs, H, 128, 256];  workspace: ivl_gdn_chunk_workspace_bytes(1, 64 * num_chunks, H). */
IVL_API int ivl_gdn_chunk_fwd_varlen(const void* q, const void* k, const void* v, const float* g, const void* beta,
                             const void* h0, int h0_dtype, void* o, void* ht, int ht_dtype, int T, int H,
                             int K, int V, float scale, int l2norm_qk, const int32_t* chunk_tok0,
                             const int32_t* chunk_valid, int num_chunks, const int32_t* seq_chunk_begin,
                             int num_seqs, void* workspace, size_t workspace_bytes, void* stream);

/* The two halves of the above, exposed so that the bench can time them separately and so that a
 * sequence-sharded caller can start prep before the previous rank's state has arrived.  prep must
 * have been enqueued on `stream` (or be otherwise ordered) before scan.  `v` is the same value tensor prep was given:
 * the scan reads its tiles itself (prep only writes the operands derived from q, k, g, beta). */
IVL_API int ivl_gdn_chunk_prep(const void* q, const void* k, const void* v, const float* g, const void* beta,
                       int B, int T, int H, float scale, int l2norm_qk, void* workspace,
                       size_t workspace_bytes, void* stream);
IVL_API int ivl_gdn_chunk_scan(const void* v, const void* h0, int h0_dtype, void* o, void* ht, int ht_dtype, int B,
                       int T, int H, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Gated DeltaNet backward (training).  Replaces chunk_gated_delta_rule_bwd --
 * fla/ops/gated_delta_rule/chunk.py:74-177,237-269 (kernels wy_fast.py:432-620, common/chunk_delta_h.py:143-244,
 * common/chunk_o.py:131-454) -- with the exact fp32 gradient of the token recurrence the operator is defined by
 * (fused_recurrent.py:85-108), recomputing states from checkpoints instead of storing them.
 *   qn, kn   fp32 [B,T,H,128]: the rows the forward used (L2-normalised and rounded to bf16 when l2norm_qk was set);
 *            the caller applies the chain rule through the normalisation to d_qn / d_kn (element-wise).
 *   v, d_o   bf16 [B,T,H,256];  g, beta fp32 [B,T,H];  h0, d_ht fp32 [B,H,128,256] or NULL (zeros).
 *   d_qn, d_kn fp32 [B,T,H,128];  d_v fp32 [B,T,H,256];  d_g, d_beta fp32 [B,T,H];  d_h0 fp32 [B,H,128,256] or NULL.
 *   workspace: ivl_gdn_bwd_workspace_bytes(B,T,H) bytes (state checkpoints every 16 tokens: 8 KiB per token per head
 *   ... / 16), 16-byte aligned.  d_qn, d_kn, d_g, d_beta are summed with atomics: results are reproducible to
 *   fp32 rounding of a sum of 16 terms, not bit for bit.
 * ---------------------------------------------------------------------------------- */
IVL_API size_t ivl_gdn_bwd_workspace_bytes(int B, int T, int H);
IVL_API int ivl_gdn_bwd(const float* qn, const float* kn, const void* v, const float* g, const float* beta,
                        const void* d_o, const float* h0, const float* d_ht, float* d_qn, float* d_kn, float* d_v,
                        float* d_g, float* d_beta, float* d_h0, int B, int T, int H, int K, int V, float scale,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Gated DeltaNet, ONE decode step of the whole mixer core in one launch: everything between the input
 * projections and o_proj of GatedDeltaNet.forward for q_len == 1 (std:1263-1342): the three
 * ShortConvolution steps (fla/modules/convolution.py:224-293), the gate math (std:1293-1294), the
 * recurrence (fused_recurrent_gated_delta_rule, std:1310-1320) and FusedRMSNormGated (std:1338).
 *   q_in,k_in [B,H*128], v_in [B,H*256]  bf16 outputs of q_proj/k_proj/v_proj for the new token
 *   a_in,b_in [B,H] bf16 (a_proj, b_proj);  gate_in [B,H*256] bf16 (g_proj)
 *   conv_weight_* [D,1,4] bf16;  A_log, dt_bias fp32 [H];  norm_weight bf16 [256]
 *   conv_state_* [B,D,4] bf16 and state [B,H,128,256] (state_dtype) are UPDATED IN PLACE
 *   out [B,H*256] bf16: the normalised, gated mixer output (input of o_proj)
 * Bit-identical to the chain ivl_short_conv_fwd x3 + ivl_gdn_gate_fwd + ivl_gdn_recurrent_fwd +
 * ivl_rmsnorm_gated_fwd on the same inputs.  Requires num_key_value_heads == num_heads.
 * ---------------------------------------------------------------------------------- */
IVL_API int ivl_gdn_decode_step(const void* q_in, const void* k_in, const void* v_in, const void* a_in,
                        const void* b_in, const void* gate_in, const void* conv_weight_q,
                        const void* conv_weight_k, const void* conv_weight_v, const float* A_log,
                        const float* dt_bias, const void* norm_weight, void* conv_state_q,
                        void* conv_state_k, void* conv_state_v, void* state, int state_dtype, void* out,
                        int B, int H, int K, int V, float scale, float norm_eps, void* stream);

/* ------------------------------------------------------------------------------------
 * Gated DeltaNet, token recurrence (decode and q_len <= 64, std:1230).
 * Replaces fused_recurrent_gated_delta_rule -- fla/ops/gated_delta_rule/fused_recurrent.py:218-335,
 * called from std:1310-1320.  Same tensors as above; exact fp32 recurrence.
 * h0 and ht may alias (in-place state update for CUDA graphs).
 * ---------------------------------------------------------------------------------- */
IVL_API int ivl_gdn_recurrent_fwd(const void* q, const void* k, const void* v, const float* g,
                          const void* beta, const void* h0, int h0_dtype, void* o, void* ht,
                          int ht_dtype, int B, int T, int H, int K, int V, float scale,
                          int l2norm_qk, void* stream);

/* ------------------------------------------------------------------------------------
 * Sliding-window causal GQA attention, prefill / chunked prefill (any Tq >= 1, Tk >= Tq).
 * Replaces the HF attention-interface callable ALL_ATTENTION_FUNCTIONS["flash_attention_2"]
 * (flash-attn wheel) used at std:1092-1108:  key j visible to query i iff
 * 0 <= (i + Tk - Tq) - j <= window - 1; the window is only enforced when Tk > window
 * (transformers/modeling_flash_attention_utils.py:627-632), window <= 0 = plain causal.
 *
 *   q [B,Tq,Hq,128], k,v [B,Tk,Hkv,128], o [B,Tq,Hq,128]  bf16, innermost dim contiguous;
 *   *_strides = {batch, time, head} strides in ELEMENTS (multiples of 8), so the HF layout
 *   [B,H,T,D] views produced by `.view(B,T,H,D).transpose(1,2)` are accepted without a copy.
 *   scale <= 0 selects D^-0.5.  D must be 128 and Hq a multiple of Hkv.
 * ---------------------------------------------------------------------------------- */
IVL_API int ivl_swa_fwd(const void* q, const int64_t* q_strides, const void* k, const int64_t* k_strides,
                        const void* v, const int64_t* v_strides, void* o, const int64_t* o_strides, int B,
                        int Tq, int Tk, int Hq, int Hkv, int D, int window, float scale, void* stream);
/* Same, with the position of key 0 in its sequence (key_pos0 >= 0; ivl_swa_fwd passes 0).  The kernel anchors its
 * 64-key tiles at absolute positions that are multiples of 64, so the result for a query does not depend on how its
 * keys were delivered: cache + new tokens (key_pos0 = tokens seen before - cached keys), halo + local shard of a
 * sequence-sharded prefill, or one long prefill all give BIT-IDENTICAL outputs (the sharded-vs-single parity gate of
 * BASELINE.md 3c).  The ring-cache entry point derives the position from its device-side counter. */
IVL_API int ivl_swa_fwd_pos(const void* q, const int64_t* q_strides, const void* k, const int64_t* k_strides,
                            const void* v, const int64_t* v_strides, void* o, const int64_t* o_strides, int B,
                            int Tq, int Tk, int Hq, int Hkv, int D, int window, float scale, int64_t key_pos0,
                            void* stream);

/* Packed variable-length batch (the reference reaches flash_attn_varlen_func through the HF
 * glue when the collator packs sequences, src/llamafactory/.../dt/workflow.py:83-92): q, k, v, o [1,T,H,128], the
 * token axis holds several sequences back to back, every sequence attends to itself only (causal, window as above
 * per sequence).  The host cuts every sequence into query tiles of <= 128 tokens that never straddle a boundary:
 * tile_tok0[i] first token of tile i, tile_seq_lo/hi[i] the [lo, hi) token range of its sequence (device int32).
 * Key tiles are anchored at the start of each sequence: bit-identical to running the sequences one by one. */
IVL_API int ivl_swa_fwd_varlen(const void* q, const int64_t* q_strides, const void* k, const int64_t* k_strides,
                               const void* v, const int64_t* v_strides, void* o, const int64_t* o_strides, int T,
                               int Hq, int Hkv, int D, int window, float scale, const int32_t* tile_tok0,
                               const int32_t* tile_seq_lo, const int32_t* tile_seq_hi, int num_tiles, void* stream);

/* Decode step of the same attention: ONE new query token per sequence (Tq == 1) against the
 * cached window, split over the key axis (HBM-bound).  q, o bf16 [B,1,Hq,128] contiguous; k, v as
 * above.  workspace: ivl_swa_decode_workspace_bytes(B, Tk, Hq) bytes of fp32 scratch. */
IVL_API size_t ivl_swa_decode_workspace_bytes(int B, int Tk, int Hq);
IVL_API int ivl_swa_decode_fwd(const void* q, const void* k, const int64_t* k_strides, const void* v,
                               const int64_t* v_strides, void* o, int B, int Tk, int Hq, int Hkv, int D,
                               int window, float scale, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Ring-buffer window cache for the sliding-window layers.  Replaces the O(W) roll of
 * StaticSlidingWindowLayerPrealloc.update (std:142-173: two concatenations + two copies of the whole 8191-token
 * window per call) and keeps the token counter in DEVICE memory, so a captured CUDA graph of a decode step or of
 * a streamed frame replays correctly (the reference's Python counters freeze under replay).
 *   ring_k, ring_v  bf16 [B, 2R, Hkv, 128] contiguous; token t lives in slots t % R and t % R + R, so the last
 *                   n <= R tokens are one contiguous run.  R >= window (decode) / >= window - 1 + Tq (ivl_swa_ring_fwd).
 *   state           int32[ivl_swa_ring_state_bytes / 4], zero-initialised by the owner: [0] = tokens appended so far.
 * ivl_swa_ring_decode  ONE launch per layer and decoded token: appends the token's K/V (k_new, v_new: [B,1,Hkv,128]
 *                   projection outputs, {batch, head} strides), attends over the last min(cum + 1, window) tokens,
 *                   combines the split-KV partials and advances the counter.  q, o: bf16 [B,1,Hq,128] contiguous.
 * ivl_swa_ring_append  appends Tq tokens ([B,Tq,Hkv,128], {batch,time,head} strides) and advances the counter.
 * ivl_swa_ring_fwd  prefill-style attention of the Tq tokens just appended against the ring: query i of the call sees
 *                   the keys of the last window tokens up to and including its own (same rule as ivl_swa_fwd).
 * ---------------------------------------------------------------------------------- */
IVL_API size_t ivl_swa_ring_workspace_bytes(int B, int Hq, int window);
IVL_API size_t ivl_swa_ring_state_bytes(int B, int Hkv);
IVL_API int ivl_swa_ring_append(const void* k, const int64_t* k_strides, const void* v, const int64_t* v_strides,
                                void* ring_k, void* ring_v, int32_t* state, int B, int Tq, int Hkv, int D, int R,
                                void* stream);
IVL_API int ivl_swa_ring_decode(const void* q, const void* k_new, const int64_t* k_new_strides, const void* v_new,
                                const int64_t* v_new_strides, void* ring_k, void* ring_v, int32_t* state, void* o, int B,
                                int Hq, int Hkv, int D, int window, int R, float scale, void* workspace,
                                size_t workspace_bytes, void* stream);
IVL_API int ivl_swa_ring_fwd(const void* q, const int64_t* q_strides, const void* ring_k, const void* ring_v,
                             const int32_t* state, void* o, const int64_t* o_strides, int B, int Tq, int Hq, int Hkv,
                             int D, int window, int R, float scale, void* stream);

/* ------------------------------------------------------------------------------------
 * Element-wise pieces of the two mixers (each one coalesced streaming launch).
 * ---------------------------------------------------------------------------------- */

/* Depthwise causal conv (kernel size 4) + optional SiLU with a carried input tail.
 * Replaces ShortConvolution.forward/step -> causal_conv1d_fn / causal_conv1d_update
 * (fla/modules/convolution.py:195-293).  x,y bf16 [B,T,D]; w bf16 [D,4] (the module's [D,1,4]);
 * cache_in / cache_out bf16 [B,D,4] = the last 4 inputs, newest in column 3; either may be NULL.
 * cache_in is left context (pip-fla semantics); cache_out must not alias cache_in.  D % 8 == 0. */
IVL_API int ivl_short_conv_fwd(const void* x, const void* w, const void* cache_in, void* y, void* cache_out,
                               int B, int T, int D, int activation_silu, void* stream);

/* Packed variable-length batch (cu_seqlens; fla/modules/convolution.py:224-251): x, y bf16 [1,T,D], no carried
 * tail; left_ctx uint8 [T] = min(3, number of tokens of the token's own sequence before it), so the window never
 * reaches into an earlier sequence. */
IVL_API int ivl_short_conv_fwd_varlen(const void* x, const void* w, void* y, const uint8_t* left_ctx, int T, int D,
                                      int activation_silu, void* stream);

/* g = -exp(A_log[h]) * softplus(a + dt_bias[h]) (fp32), beta = sigmoid(b) (bf16) -- std:1293-1294.
 * a, b bf16 [n_tokens, H] (outputs of a_proj / b_proj); A_log, dt_bias fp32 [H]. */
IVL_API int ivl_gdn_gate_fwd(const void* a, const void* b, const float* A_log, const float* dt_bias, float* g,
                             void* beta, int64_t n_tokens, int H, void* stream);

/* y = x * rsqrt(mean(x^2) + eps) * w * gate * sigmoid(gate) over rows of 256 (FusedRMSNormGated,
 * fla/modules/fused_norm_gate.py:26-95,735-797).  x, gate, y bf16 [rows,256]; w bf16 [256]. */
IVL_API int ivl_rmsnorm_gated_fwd(const void* x, const void* gate, const void* w, void* y, int64_t rows, int dim,
                                  float eps, void* stream);

/* In-place rotary embedding of x [B,T,Hn,128] (strides {batch,time,head} in elements) with merged
 * per-token cos/sin bf16 [B,T,128]; bit-exact with the reference's bf16 expression
 * q*cos + rotate_half(q)*sin (std:949-984). */
IVL_API int ivl_mrope_apply(void* x, const int64_t* x_strides, const void* cos, const void* sin, int B, int T,
                            int Hn, int D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IVL_B200_H_ */
