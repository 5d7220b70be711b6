// tma_order.cu -- does cp.async.bulk.wait_group 0 order a bulk store (shared -> global) before a later bulk load /
// bulk store / generic load of the same global region issued by the same thread?  One CTA per region, many rounds:
//   test 0: store pattern k from buffer A, wait_group 0, bulk-load the region into buffer B, compare
//   test 1: store pattern k from A, wait_group 0, store pattern k+1 from B, wait_group 0, generic ld.cg read-back
//   test 2: as test 1 but wait_group.read between the two stores (what a pipeline that only recycles shared memory does)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_order tma_order.cu && ./tma_order
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void *g, const void *s, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(s)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void *s, const void *g, uint32_t bytes, uint64_t *bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s)),
               "l"(g), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!ok);
}

constexpr int BYTES = 65536, NQ = BYTES / 8;

__global__ void __launch_bounds__(512) order_kernel(double *G, int rounds, int test, unsigned long long *bad) {
  extern __shared__ __align__(128) unsigned char raw[];
  double *A = reinterpret_cast<double *>(raw), *B = A + NQ;
  __shared__ uint64_t bar;
  double *g = G + (size_t)blockIdx.x * NQ;
  const int tid = threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    fence_async();
  }
  __syncthreads();
  uint32_t ph = 0;
  unsigned long long nbad = 0;
  for (int k = 0; k < rounds; ++k) {
    for (int i = tid; i < NQ; i += blockDim.x) A[i] = (double)(k * 3 + 1) + i * 1e-6;
    fence_async();
    __syncthreads();
    if (test == 0) {
      if (tid == 0) {
        bulk_store(g, A, BYTES);
        wait_all();
        bulk_load(B, g, BYTES, &bar);
      }
      mbar_wait(&bar, ph);
      ph ^= 1;
      for (int i = tid; i < NQ; i += blockDim.x) nbad += B[i] != (double)(k * 3 + 1) + i * 1e-6;
    } else {
      for (int i = tid; i < NQ; i += blockDim.x) B[i] = (double)(k * 3 + 2) + i * 1e-6;
      fence_async();
      __syncthreads();
      if (tid == 0) {
        bulk_store(g, A, BYTES);
        if (test == 1) wait_all(); else wait_read();
        bulk_store(g, B, BYTES);
        wait_all();
      }
      __syncthreads();
      for (int i = tid; i < NQ; i += blockDim.x) nbad += __ldcg(g + i) != (double)(k * 3 + 2) + i * 1e-6;
    }
    __syncthreads();
  }
  if (nbad) atomicAdd(bad, nbad);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int blocks = prop.multiProcessorCount;
  double *G;
  unsigned long long *bad, h;
  cudaMalloc(&G, (size_t)blocks * BYTES);
  cudaMalloc(&bad, 8);
  cudaFuncSetAttribute(order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * BYTES);
  for (int test = 0; test < 3; ++test) {
    cudaMemset(bad, 0, 8);
    order_kernel<<<blocks, 512, 2 * BYTES>>>(G, 2000, test, bad);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost);
    printf("test %d: %s, mismatching doubles %llu of %llu\n", test, cudaGetErrorString(e), h, (unsigned long long)blocks * 2000 * NQ);
  }
  return 0;
}
