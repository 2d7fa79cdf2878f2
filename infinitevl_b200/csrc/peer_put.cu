// Neighbour hand-off of the sequence-sharded prefill through PEER MEMORY (no reference
// counterpart -- the reference never shards a sequence).
//
// peer_put: copy a contiguous buffer into memory that lives on ANOTHER GPU of the node (opened in this process
// through CUDA IPC, reached over NVLink with ordinary stores) and then publish a 32-bit flag next to it, in one
// launch.  The receiver's stream waits for the flag with a stream memory operation (ivl_stream_wait_value32), so
// neither side runs a rendezvous and the receiver spends no SM on the transfer.  NCCL's send / recv pair costs the
// two-GPU step 12 % (profiles/r02_summary.md); torch's cross-device copy_ onto IPC memory takes 2.2 ms per call
// (tools/exp_p2p.py), hence this kernel.
//
// Ordering: every block stores its share, fences at system scope and bumps a block counter; the last block to
// arrive resets the counter and writes the flag with a system-scope release -- the split-K semaphore pattern with
// the flag on the other GPU.  bytes == 0 publishes the flag only (the "consumed" acknowledgement going back).
#include <stdint.h>

#include <cuda_runtime.h>

namespace ivl {

namespace {

__global__ void __launch_bounds__(256)
peer_put_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16, uint32_t* flag, uint32_t value,
                uint32_t* counter) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(counter, 1u) == gridDim.x - 1) {
      *counter = 0;
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
    }
  }
}

}  // namespace

cudaError_t launch_peer_put(void* dst, const void* src, size_t bytes, uint32_t* flag, uint32_t value, uint32_t* counter,
                            cudaStream_t stream) {
  const size_t n16 = bytes / 16;
  // enough blocks to keep a few hundred GB/s of NVLink stores in flight, few enough to slip in next to the
  // compute kernels (no shared memory, 256 threads)
  unsigned blocks = (unsigned)((n16 + 256 * 8 - 1) / (256 * 8));
  if (blocks < 1) blocks = 1;
  if (blocks > 64) blocks = 64;
  peer_put_kernel<<<blocks, 256, 0, stream>>>(static_cast<uint4*>(dst), static_cast<const uint4*>(src), n16, flag, value,
                                              counter);
  return cudaGetLastError();
}

}  // namespace ivl
