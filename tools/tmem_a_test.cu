// tmem_a_test.cu -- micro-test: tcgen05.mma with the A operand in TENSOR MEMORY (written with tcgen05.st,
// lane = row, 2 bf16 per 32-bit column) and B in shared memory (SWIZZLE_128B, K-major), M=128, N=NT, K=128.
// Prints the max error against a CPU reference.   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tmem_a_test tools/tmem_a_test.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../efficient-visual-document-retrieval_b200/csrc/ptx.cuh"
using namespace evdr::ptx;

constexpr int NT = 192;   // N of the MMA (page tokens per tile)

__device__ __forceinline__ void tmem_st64(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, "
      "%33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, "
      "%49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, %64};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]),
        "r"(r[32]), "r"(r[33]), "r"(r[34]), "r"(r[35]), "r"(r[36]), "r"(r[37]), "r"(r[38]), "r"(r[39]),
        "r"(r[40]), "r"(r[41]), "r"(r[42]), "r"(r[43]), "r"(r[44]), "r"(r[45]), "r"(r[46]), "r"(r[47]),
        "r"(r[48]), "r"(r[49]), "r"(r[50]), "r"(r[51]), "r"(r[52]), "r"(r[53]), "r"(r[54]), "r"(r[55]),
        "r"(r[56]), "r"(r[57]), "r"(r[58]), "r"(r[59]), "r"(r[60]), "r"(r[61]), "r"(r[62]), "r"(r[63])
      : "memory");
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(128, 1)
test_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sB = base;                                   // 2 K sub-tiles of NT rows x 128 B
  const uint32_t bar = base + 2 * NT * 128;
  const uint32_t slot = bar + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + (slot - base));
  // B -> shared memory in the SWIZZLE_128B K-major layout TMA would produce
  for (int i = tid; i < NT * 16; i += 128) {
    const int r = i / 16, c16 = i % 16;                        // 16-byte chunk c16 of row r (256 B per row)
    const int sub = c16 / 8, c = c16 % 8;
    const uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)r * 128 + c16 * 8);
    *reinterpret_cast<uint4*>(sm + sub * (NT * 128) + (r / 8) * 1024 + (r % 8) * 128 + ((c ^ (r % 8)) * 16)) = v;
  }
  // A row `tid` -> TMEM lanes [32*warp, +32), columns [384, 448): 2 bf16 per column, K ascending
  uint32_t a[64];
  const uint4* arow = reinterpret_cast<const uint4*>(A + (size_t)tid * 128);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint4 v = arow[i];
    a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
  }
  const uint32_t a_col = 384;
  tmem_st64(tmem + ((uint32_t)(warp * 32) << 16) + a_col, a);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> async proxy (MMA)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    constexpr uint32_t IDESC = make_idesc_bf16(128, NT);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint64_t db = make_sw128_desc(sB + (k >> 2) * (NT * 128) + (k & 3) * 32);
      mma_bf16_ts(tmem, tmem + a_col + k * 8, db, IDESC, k > 0);
    }
    mma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int g = 0; g < NT / 32; ++g) {
    float v[32];
    tmem_ld32_issue(tmem + ((uint32_t)(warp * 32) << 16) + g * 32, v);
    tmem_wait_ld_tied(v);
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)tid * NT + g * 32 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main() {
  std::vector<__nv_bfloat16> hA(128 * 128), hB(NT * 128);
  std::vector<float> fA(128 * 128), fB(NT * 128), hD(128 * NT);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(x); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(x); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, hD.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, hD.size() * 4);
  const int smem = 2 * NT * 128 + 64 + 1024;
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  test_kernel<<<1, 128, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < NT; ++n) {
      double ref = 0;
      for (int k = 0; k < 128; ++k) ref += (double)fA[m * 128 + k] * fB[n * 128 + k];
      double err = fabs(ref - hD[m * NT + n]);
      if (err > maxerr) maxerr = err;
      if (err > 1e-2) ++bad;
    }
  printf("A-in-TMEM MMA M=128 N=%d K=128: max |err| = %.3e, mismatches %d / %d\n", NT, maxerr, bad, 128 * NT);
  printf("D[0][0..3] = %f %f %f %f\n", hD[0], hD[1], hD[2], hD[3]);
  return 0;
}
