// The one exchange step of data-parallel RENI training on NVLink / NVSwitch: in-place all-reduce (mean) of the flat fp32
// weight-gradient buffer across the ranks of one node, written against peer memory instead of calling NCCL so that it
// can sit INSIDE the captured step graph, directly behind the kernels that produce the gradients.
//
// Reference semantics: Lightning's DDPStrategy (run.py:97) averages every gradient over the ranks.
//
// Every rank passes the same symmetric allocation (torch.distributed._symmetric_memory supplies the mapping: an array
// of the W peer pointers, and, on NVSwitch systems, one multicast address that aliases all W buffers):
//   barrier A   every rank's gradient kernels are complete (flag exchange in a second symmetric buffer)
//   reduce      rank r owns the r-th 1/W slice:
//                 multicast : v = multimem.ld_reduce.add [mc + i]   (the switch adds the W copies: NVLS)
//                             multimem.st [mc + i], v * scale       (the switch writes all W copies)
//                 peer ptrs : v = sum_p ld [buf_p + i] ; st [buf_p + i], v * scale for every p   (two-shot over P2P)
//   barrier B   every rank's slice has been written everywhere
// Flags are monotone: every block keeps its own call counter (epoch[b], incremented by the block itself, so a captured
// graph replays correctly); block b of rank r signals slot [b][r] of every peer with 2 * epoch + 1 (+ 2 for barrier B)
// and waits for its own slots.  All waits are bounded: a missing peer sets *status instead of hanging the device.
// One launch; up to one block per SM with four 16-byte multimem / peer loads in flight per thread (the first version
// -- 32 blocks, one load in flight -- took 66 us for 2.7 MB on two GPUs, latency-bound; NCCL takes 53 us).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace reni {

constexpr int kArThreads = 512;
constexpr int kArMaxBlocks = 160;
constexpr int kArMaxWorld = 16;
constexpr int kArUnroll = 4;

struct AllReduceParams {
  float* const* bufs;        // device array of W pointers to the ranks' buffers (own buffer at [rank])
  uint32_t* const* flags;    // device array of W pointers to the ranks' flag blocks (kArMaxBlocks * kArMaxWorld words)
  float* mc;                 // multicast address of the buffers, or null
  int64_t n4;                // float4 elements
  int rank, world;
  float scale;               // 1 / world for the mean
  uint32_t* epoch;           // kArMaxBlocks device words: per-block call counters
  uint32_t* status;          // set to 1 if a wait gave up
};

__device__ __forceinline__ void ar_signal(uint32_t* addr, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ar_peek(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
// all ranks' block b meet: thread t < world signals peer t and waits for peer t's signal
__device__ __forceinline__ void ar_barrier(const AllReduceParams& p, uint32_t target) {
  __syncthreads();
  if ((int)threadIdx.x < p.world) {
    const int peer = threadIdx.x;
    __threadfence_system();
    ar_signal(p.flags[peer] + blockIdx.x * kArMaxWorld + p.rank, target);
    const uint32_t* mine = p.flags[p.rank] + blockIdx.x * kArMaxWorld + peer;
    uint32_t spins = 0;
    while ((int32_t)(ar_peek(mine) - target) < 0) {
      if (++spins > (1u << 24)) {  // ~ seconds
        *p.status = 1u;
        break;
      }
      if (spins > 64) __nanosleep(32);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kArThreads) reni_allreduce_kernel(const AllReduceParams p) {
  const uint32_t e = p.epoch[blockIdx.x] + 1;  // this block's call number (every rank launches the same grid)
  ar_barrier(p, 2 * e + 1);
  const int64_t per = (p.n4 + p.world - 1) / p.world;
  const int64_t lo = (int64_t)p.rank * per;
  const int64_t hi = lo + per < p.n4 ? lo + per : p.n4;
  const int64_t stride = (int64_t)gridDim.x * kArThreads;
  if (p.mc != nullptr) {
    for (int64_t i0 = lo + (int64_t)blockIdx.x * kArThreads + threadIdx.x; i0 < hi; i0 += kArUnroll * stride) {
      float4 v[kArUnroll];
#pragma unroll
      for (int u = 0; u < kArUnroll; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < hi)
          asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                       : "l"(p.mc + 4 * i)
                       : "memory");
      }
#pragma unroll
      for (int u = 0; u < kArUnroll; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < hi)
          asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.mc + 4 * i),
                       "f"(v[u].x * p.scale), "f"(v[u].y * p.scale), "f"(v[u].z * p.scale), "f"(v[u].w * p.scale)
                       : "memory");
      }
    }
  } else {
    for (int64_t i = lo + (int64_t)blockIdx.x * kArThreads + threadIdx.x; i < hi; i += stride) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int q = 0; q < p.world; ++q) {
        float4 v;  // (peer memory, written by another device before barrier A: bypass L1)
        asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "l"(reinterpret_cast<const float4*>(p.bufs[q]) + i)
                     : "memory");
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      acc.x *= p.scale; acc.y *= p.scale; acc.z *= p.scale; acc.w *= p.scale;
      for (int q = 0; q < p.world; ++q) reinterpret_cast<float4*>(p.bufs[q])[i] = acc;
    }
  }
  ar_barrier(p, 2 * e + 2);  // (bar.sync + one system-scope fence per signalling thread publish the block's stores)
  if (threadIdx.x == 0) p.epoch[blockIdx.x] = e;
}

}  // namespace reni
