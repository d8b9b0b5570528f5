// Probe: issue rate of tcgen05.mma.cta_group::1.kind::f16 (M = 128, K = 16, both operands in shared memory,
// K-major SWIZZLE_128B) as a function of N, measured with clock64 around NREP back-to-back MMAs + one commit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate_probe mma_rate_probe.cu ; run on a B200.
// Prints cycles per MMA next to the nominal floor 128*N/256 (8192 dense FLOP/clk/SM).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {   // f16 x f16 -> f32, K-major
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int N, int DISTINCT>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int nrep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < (2 * 16384 + 2 * N * 128) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;  // fp16 1.0
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, N);
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 2 * 16384;    // A: 2 x 16 KB tiles, B: 2 x (N*128 B)
    long long t0 = clock64();
    for (int r = 0; r < nrep; ++r) {
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const uint32_t sel = DISTINCT ? (r & 1) : 0;
        const uint64_t a = make_desc(a_base + sel * 16384 + k4 * 32), b = make_desc(b_base + sel * (N * 128) + k4 * 32);
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
            "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)(r | k4))
            : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();   // issue done
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    long long t2 = clock64();   // all MMAs retired
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// The fp16x2 split sequence of conv_tc_kernel per K16 step: a0 x [w0; w1] (N = 2*BN -> MAIN | CORR) and a1 x w0 (N = BN -> CORR),
// optionally with a commit after every 8 MMAs (one weight tile) as the kernel does.
template <int BN, int COMMIT>
__global__ void __launch_bounds__(128, 1) probe_split(long long* out, int nrep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < (2 * 16384 + 2 * BN * 128) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc2 = make_idesc(128, 2 * BN), idesc1 = make_idesc(128, BN);
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 2 * 16384;
    long long t0 = clock64();
    for (int r = 0; r < nrep; ++r) {
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const uint64_t a0 = make_desc(a_base + k4 * 32), a1 = make_desc(a_base + 16384 + k4 * 32), w0 = make_desc(b_base + k4 * 32);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
                     "l"(a0), "l"(w0), "r"(idesc2), "r"((uint32_t)(r | k4)) : "memory");
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem + BN),
                     "l"(a1), "l"(w0), "r"(idesc1), "r"(1u) : "memory");
      }
      if (COMMIT) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int BN, int COMMIT>
void run_split(long long* d, int nrep) {
  const size_t smem = 1024 + 2 * 16384 + 2 * BN * 128;
  cudaFuncSetAttribute(probe_split<BN, COMMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long h[2];
  for (int it = 0; it < 2; ++it) {
    probe_split<BN, COMMIT><<<1, 128, smem>>>(d, nrep);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  }
  cudaError_t e = cudaGetLastError();
  const double n = 4.0 * nrep;   // K16 steps
  printf("{\"pattern\": \"fp16x2 split, BN=%d\", \"commit_per_8_mma\": %d, \"issue_clk_per_k16\": %.1f, \"retire_clk_per_k16\": %.1f, \"floor_clk\": %.1f, \"err\": \"%s\"}\n",
         BN, COMMIT, h[0] / n, h[1] / n, 128.0 * (3 * BN) / 256.0, cudaGetErrorString(e));
}

template <int N, int DISTINCT>
void run(long long* d, int nrep) {
  const size_t smem = 1024 + 2 * 16384 + 2 * N * 128;
  cudaFuncSetAttribute(probe<N, DISTINCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long h[2];
  for (int it = 0; it < 2; ++it) {
    probe<N, DISTINCT><<<1, 128, smem>>>(d, nrep);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  }
  cudaError_t e = cudaGetLastError();
  const double n = 4.0 * nrep;
  printf("{\"N\": %d, \"alternate_tiles\": %d, \"mmas\": %d, \"issue_clk_per_mma\": %.1f, \"retire_clk_per_mma\": %.1f, \"floor_clk\": %.1f, \"err\": \"%s\"}\n",
         N, DISTINCT, (int)n, h[0] / n, h[1] / n, 128.0 * N / 256.0, cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int nrep = 512;
  run<256, 0>(d, nrep); run<256, 1>(d, nrep);
  run<128, 0>(d, nrep); run<128, 1>(d, nrep);
  run<64, 0>(d, nrep);  run<64, 1>(d, nrep);
  run<32, 0>(d, nrep);  run<16, 0>(d, nrep);
  run_split<128, 0>(d, nrep); run_split<128, 1>(d, nrep);
  run_split<64, 0>(d, nrep);  run_split<64, 1>(d, nrep);
  return 0;
}
