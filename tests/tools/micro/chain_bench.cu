// Microbenchmark (GPU): cycles per term of a dependent fp32 add chain fed from shared memory, as the strict sums of the
// matcher run it (lane k < 9*GS owns one chain). build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false
#include <cstdio>
#include <cuda_runtime.h>
#define PITCH 388
template <int MODE, int AHEAD>
__global__ void k(float* out, long long* cyc, int n, int lanes) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 18 * PITCH + 64; i += blockDim.x) sm[i] = 1.0f + i * 1e-7f;
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  float acc = 0.f;
  long long t0 = clock64();
  for (int rep = 0; rep < 16; ++rep) {
    if (lane < lanes) {
      const float* q = sm + lane * PITCH;
      if (MODE == 0) {          // as compiled in the product: float4 loads, 4 adds each
        const float4* q4 = reinterpret_cast<const float4*>(q);
        const int n4 = n >> 2;
        for (int i = 0; i < n4; ++i) { const float4 v = q4[i]; acc += v.x; acc += v.y; acc += v.z; acc += v.w; }
      } else if (MODE == 1) {   // explicit register ring: AHEAD float4 loads in flight
        const float4* q4 = reinterpret_cast<const float4*>(q);
        const int n4 = n >> 2;
        float4 r[AHEAD];
#pragma unroll
        for (int u = 0; u < AHEAD; ++u) r[u] = q4[u];
        for (int i = 0; i < n4; i += AHEAD) {
#pragma unroll
          for (int u = 0; u < AHEAD; ++u) {
            const float4 v = r[u];
            if (i + u + AHEAD < n4 + AHEAD) r[u] = q4[i + u + AHEAD];   // (reads a few words past the row: padded)
            acc += v.x; acc += v.y; acc += v.z; acc += v.w;
          }
        }
      } else if (MODE == 2) {   // no loads at all: the FADD chain alone
        float a = q[0];
        for (int i = 0; i < n; ++i) acc += a;
      }
    }
    acc = __shfl_sync(0xffffffffu, acc, (lane + 1) & 31);
  }
  long long t1 = clock64();
  if (lane == 0) *cyc = t1 - t0;
  out[lane] = acc;
}
template <int MODE, int AHEAD>
void run(const char* name, int lanes, int nblk) {
  float* out; long long* cyc;
  cudaMalloc(&out, 32 * 4); cudaMalloc(&cyc, 8);
  const int smem = (18 * PITCH + 64 + 64) * 4;
  k<MODE, AHEAD><<<nblk, 128, smem>>>(out, cyc, 360, lanes);
  k<MODE, AHEAD><<<nblk, 128, smem>>>(out, cyc, 360, lanes);
  long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-40s lanes %2d blocks/SM %d: %.2f cycles per term (%s)\n", name, lanes, nblk / 148, (double)h / (16.0 * 360.0), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int lanes : {9, 18, 27}) {
    run<0, 1>("float4 loads as compiled", lanes, 148);
    run<1, 2>("register ring, 2 loads ahead", lanes, 148);
    run<1, 4>("register ring, 4 loads ahead", lanes, 148);
    run<1, 8>("register ring, 8 loads ahead", lanes, 148);
    run<2, 1>("no loads", lanes, 148);
  }
  run<0, 1>("float4 loads as compiled, 4 CTAs/SM", 9, 148 * 4);
  run<0, 1>("float4 loads as compiled, 7 CTAs/SM", 9, 148 * 7);
  return 0;
}
