// Microbenchmark: the memory phase of the GRU q epilogue (3 fp32 maps read, 1 written + 2 half planes written per
// output) with different lane -> (row, column group) mappings, at the epilogue's occupancy (8 warps per SM, one block
// per SM, every warp walks 32-row x 128-column tiles in 16- or 32-column steps).
//   A: 4 lanes x 16 B per row, 8 rows per instruction (the kernel's layout)
//   B: 8 lanes x 16 B per row, 4 rows per instruction (32-column steps, full 128-byte lines)
//   C: 2 lanes x 32 B per row, 16 rows per instruction (256-bit accesses)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o epi_probe epi_pattern_probe.cu && ./epi_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int LPR>   // lanes per row (16 B each)
__global__ void __launch_bounds__(256, 1) probe(const float* __restrict__ pre, float* __restrict__ h, const float* __restrict__ z,
                                                __half* __restrict__ pl, int rows, int ld, long long pl_stride) {
  constexpr int RPI = 32 / LPR;          // rows per instruction
  constexpr int COLS = LPR * 4;          // columns per step
  constexpr int NIT = 32 / RPI;          // instructions per array per 32-row step
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = rows / 32;           // 32-row x ld-column tiles, dealt to (block, warp)
  for (int t = blockIdx.x * 8 + warp; t < tiles; t += gridDim.x * 8) {
    for (int c = 0; c < ld; c += COLS) {
      float4 a[NIT], b[NIT], g[NIT];
      const int col = c + (lane % LPR) * 4;
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        const size_t off = (size_t)(t * 32 + i * RPI + lane / LPR) * ld + col;
        a[i] = __ldg(reinterpret_cast<const float4*>(pre + off));
        b[i] = *reinterpret_cast<const float4*>(h + off);
        g[i] = __ldg(reinterpret_cast<const float4*>(z + off));
      }
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        const size_t off = (size_t)(t * 32 + i * RPI + lane / LPR) * ld + col;
        float4 o = make_float4(fmaf(g[i].x, a[i].x - b[i].x, b[i].x), fmaf(g[i].y, a[i].y - b[i].y, b[i].y),
                               fmaf(g[i].z, a[i].z - b[i].z, b[i].z), fmaf(g[i].w, a[i].w - b[i].w, b[i].w));
        *reinterpret_cast<float4*>(h + off) = o;
        const __half2 h01 = __floats2half2_rn(o.x, o.y), h23 = __floats2half2_rn(o.z, o.w);
        *reinterpret_cast<uint2*>(pl + off) = make_uint2(*reinterpret_cast<const unsigned*>(&h01), *reinterpret_cast<const unsigned*>(&h23));
        *reinterpret_cast<uint2*>(pl + pl_stride + off) = make_uint2(*reinterpret_cast<const unsigned*>(&h23), *reinterpret_cast<const unsigned*>(&h01));
      }
    }
  }
}

int main() {
  const int rows = 18 * 4096, ld = 128;
  const size_t n = (size_t)rows * ld;
  float *pre, *h, *z; __half* pl;
  cudaMalloc(&pre, n * 4); cudaMalloc(&h, n * 4); cudaMalloc(&z, n * 4); cudaMalloc(&pl, n * 4);
  cudaMemset(pre, 0, n * 4); cudaMemset(h, 0, n * 4); cudaMemset(z, 0, n * 4);
  float* flush; cudaMalloc(&flush, 256 << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto kern) {
    float best = 1e9, sum = 0;
    for (int r = 0; r < 12; ++r) {
      cudaMemsetAsync(flush, r, 256 << 20);
      cudaEventRecord(e0);
      kern<<<148, 256>>>(pre, h, z, pl, rows, ld, (long long)n);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r >= 2) { sum += ms; if (ms < best) best = ms; }
    }
    printf("{\"pattern\": \"%s\", \"us_mean\": %.1f, \"us_best\": %.1f, \"GBps\": %.0f}\n", name, sum / 10 * 1e3, best * 1e3, n * 20.0 / (sum / 10 * 1e-3) / 1e9);
  };
  run("A: 8 rows x 64 B per instruction (kernel layout)", probe<4>);
  run("B: 4 rows x 128 B per instruction", probe<8>);
  run("D: 2 rows x 256 B per instruction", probe<16>);
  run("E: 1 row x 512 B per instruction", probe<32>);
  run("C': 16 rows x 32 B (2 lanes x 16 B)", probe<2>);
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
