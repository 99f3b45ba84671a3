// Micro-benchmark: shared -> global store paths of one SM while all 148 SMs stream (the delta-stash write of the
// backward chain).  Per variant: clocks to issue, until the sources have been read, until the writes are complete.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_s2g tools/ubench/s2g.cu && build/ubench_s2g
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_s2g(void* g, const void* s, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(s)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

constexpr int kTile = 65536;
// mode: piece size in bytes for bulk copies (512..65536), or 0 = st.global.v4 from registers (256 threads x 16 x 16 B)
__global__ void __launch_bounds__(256, 1) k(uint8_t* dst, int reps, int piece, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < kTile / 16; i += 256) reinterpret_cast<uint4*>(smem)[i] = make_uint4(i, 1, 2, 3);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  long long t_issue = 0, t_read = 0, t_done = 0;
  const int npieces = piece ? kTile / piece : 0;
  for (int r = 0; r < reps; ++r) {
    uint8_t* d = dst + ((size_t)r * gridDim.x + blockIdx.x) * kTile;
    __syncthreads();
    const long long t0 = clock64();
    if (piece) {
      for (int i = threadIdx.x; i < npieces; i += 256) bulk_s2g(d + (size_t)i * piece, smem + (size_t)i * piece, piece);
      if (threadIdx.x < npieces) commit();
      const long long t1 = clock64();
      if (threadIdx.x < npieces) wait_read0();
      __syncthreads();
      const long long t2 = clock64();
      if (threadIdx.x < npieces) wait0();
      __syncthreads();
      const long long t3 = clock64();
      t_issue += t1 - t0; t_read += t2 - t0; t_done += t3 - t0;
    } else {
      uint4 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = reinterpret_cast<const uint4*>(smem)[i * 256 + threadIdx.x];
#pragma unroll
      for (int i = 0; i < 16; ++i) reinterpret_cast<uint4*>(d)[i * 256 + threadIdx.x] = v[i];
      const long long t1 = clock64();
      __syncthreads();
      const long long t2 = clock64();
      t_issue += t1 - t0; t_read += t2 - t0; t_done += t2 - t0;
    }
  }
  if (threadIdx.x == 0) { out[blockIdx.x * 3] = t_issue / reps; out[blockIdx.x * 3 + 1] = t_read / reps; out[blockIdx.x * 3 + 2] = t_done / reps; }
}
int main() {
  const int grid = 148, reps = 64;
  uint8_t* dst; long long* out;
  cudaMalloc(&dst, (size_t)grid * reps * kTile);
  cudaMalloc(&out, grid * 3 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kTile);
  const int pieces[] = {512, 1024, 2048, 4096, 16384, 65536, 0};
  for (int g : {148, 8}) {
    for (int piece : pieces) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      k<<<g, 256, kTile>>>(dst, reps, piece, out);  // warm
      cudaEventRecord(a);
      k<<<g, 256, kTile>>>(dst, reps, piece, out);
      cudaEventRecord(b);
      cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, a, b);
      long long h[148 * 3]; cudaMemcpy(h, out, g * 3 * sizeof(long long), cudaMemcpyDeviceToHost);
      double s0 = 0, s1 = 0, s2 = 0; for (int i = 0; i < g; ++i) { s0 += h[3 * i]; s1 += h[3 * i + 1]; s2 += h[3 * i + 2]; }
      printf("grid %3d piece %6d: issue %7.0f  read-done %7.0f  complete %7.0f clk per 64 KB | %.1f B/clk/SM | %.2f TB/s chip (%s)\n",
             g, piece, s0 / g, s1 / g, s2 / g, kTile / (s2 / g), (double)g * reps * kTile / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
