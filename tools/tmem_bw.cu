// tmem_bw.cu -- microbenchmark of tcgen05.ld throughput / latency on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bw tools/tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(addr, r)                                                                                   \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                 \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23," \
               "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),     \
                 "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), \
                 "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),           \
                 "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),           \
                 "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])            \
               : "r"(addr) : "memory")
#define LD16(addr, r)                                                                                   \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                 \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                         \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),     \
                 "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), \
                 "=r"(r[14]), "=r"(r[15])                                                                \
               : "r"(addr) : "memory")
#define WAITLD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

// MODE 0: x32 load + wait each (latency chain).  MODE 1: two x32 in flight.  MODE 2: four x32 in flight.
// MODE 3: x16 + wait each.  MODE 4: eight x32 in flight (256 regs -> uses 4 live + xor-fold to bound regs)
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int iters, int nwarps, long long* cyc, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    uint32_t a[32], b[32], c[32], d[32];
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t col = (uint32_t)((i * 128 + (warp >> 2) * 256) & 511);
      if (MODE == 0) {
        LD32(base + (col & 480), a); WAITLD(); acc ^= a[0] ^ a[31];
      } else if (MODE == 1) {
        LD32(base + (col & 448), a); LD32(base + (col & 448) + 32, b); WAITLD(); acc ^= a[0] ^ b[31];
      } else if (MODE == 2) {
        LD32(base + (col & 384), a); LD32(base + (col & 384) + 32, b); LD32(base + (col & 384) + 64, c);
        LD32(base + (col & 384) + 96, d); WAITLD();
        acc ^= a[0] ^ b[31] ^ c[7] ^ d[13];
      } else if (MODE == 3) {
        LD16(base + (col & 496), a); WAITLD(); acc ^= a[0] ^ a[15];
      }
    }
    t1 = clock64();
  }
  if ((threadIdx.x & 31) == 0 && warp < nwarps) cyc[blockIdx.x * 8 + warp] = t1 - t0;
  sink[blockIdx.x * 256 + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int MODE>
void run(const char* name, int bytes_per_iter_per_warp, int nwarps) {
  const int iters = 20000, blocks = 148;
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, blocks * 8 * sizeof(long long)); cudaMalloc(&sink, blocks * 256 * 4);
  cudaMemset(cyc, 0, blocks * 8 * sizeof(long long));
  k<MODE><<<blocks, 256>>>(iters, nwarps, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[148 * 8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0; for (int i = 0; i < nwarps; ++i) mx = h[i] > mx ? h[i] : mx;
  double per_iter = mx / iters;
  printf("%-28s warps/SM %d: %7.1f clk/iter  -> %6.1f B/clk/warp  %7.1f B/clk/SM\n", name, nwarps, per_iter,
         bytes_per_iter_per_warp / per_iter, bytes_per_iter_per_warp * (double)nwarps / per_iter);
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  for (int nw : {1, 4, 8}) {
    run<3>("x16 + wait", 16 * 32 * 4, nw);
    run<0>("x32 + wait", 32 * 32 * 4, nw);
    run<1>("2 x32 in flight + wait", 2 * 32 * 32 * 4, nw);
    run<2>("4 x32 in flight + wait", 4 * 32 * 32 * 4, nw);
  }
  return 0;
}
