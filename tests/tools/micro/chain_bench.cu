// Microbenchmark (GPU): cycles per term of a dependent fp32 add chain fed from shared memory, as the strict sums of the
// matcher run it (lane k < 9*GS owns one chain), for NW chain warps per SM placed on one scheduler or on different ones.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false
#include <cstdio>
#include <cuda_runtime.h>
#define PITCH 388
#define STRIDE 5508
template <int MODE, int AHEAD>
__global__ void k(float* out, long long* cyc, int n, int lanes, int nw, int same_sched) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 7 * STRIDE + 64; i += blockDim.x) sm[i] = 1.0f + i * 1e-7f;
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int g = -1;
  if (same_sched) { if ((w & 3) == 3 && (w >> 2) < nw) g = w >> 2; }
  else if (w < nw) g = w;
  if (g < 0) return;
  float acc = 0.f;
  long long t0 = clock64();
  for (int rep = 0; rep < 16; ++rep) {
    if (lane < lanes) {
      const float* q = sm + (g % 3) * 2 * STRIDE + (lane / 9) * STRIDE + (lane % 9) * PITCH;
      const float4* q4 = reinterpret_cast<const float4*>(q);
      const int n4 = n >> 2;
      if (MODE == 0) {          // as compiled in the product: float4 loads, 4 adds each
        for (int i = 0; i < n4; ++i) { const float4 v = q4[i]; acc += v.x; acc += v.y; acc += v.z; acc += v.w; }
      } else if (MODE == 1) {   // explicit register ring: AHEAD float4 loads in flight
        float4 r[AHEAD];
#pragma unroll
        for (int u = 0; u < AHEAD; ++u) r[u] = q4[u];
        for (int i = 0; i < n4; i += AHEAD) {
#pragma unroll
          for (int u = 0; u < AHEAD; ++u) {
            const float4 v = r[u];
            r[u] = q4[i + u + AHEAD];   // (reads a few words past the row: padded)
            acc += v.x; acc += v.y; acc += v.z; acc += v.w;
          }
        }
      }
    }
    acc = __shfl_sync(0xffffffffu, acc, (lane + 1) & 31);
  }
  long long t1 = clock64();
  if (lane == 0 && g == 0) *cyc = t1 - t0;
  out[threadIdx.x & 31] = acc;
}
template <int MODE, int AHEAD>
void run(const char* name, int lanes, int nw, int same) {
  float* out; long long* cyc;
  cudaMalloc(&out, 32 * 4); cudaMalloc(&cyc, 8);
  const int smem = (7 * STRIDE + 64 + 64) * 4;
  cudaFuncSetAttribute(k<MODE, AHEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<MODE, AHEAD><<<148, 896, smem>>>(out, cyc, 360, lanes, nw, same);
  k<MODE, AHEAD><<<148, 896, smem>>>(out, cyc, 360, lanes, nw, same);
  long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-32s lanes %2d, %d chain warps on %s: %.2f cycles per term (%s)\n", name, lanes, nw, same ? "ONE scheduler" : "different schedulers",
         (double)h / (16.0 * 360.0), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int same = 0; same < 2; ++same)
    for (int nw : {1, 2, 4, 7}) {
      if (!same && nw > 4) continue;
      run<0, 1>("float4 loads as compiled", 9, nw, same);
      run<0, 1>("float4 loads as compiled", 18, nw, same);
      run<1, 4>("register ring, 4 loads ahead", 18, nw, same);
    }
  return 0;
}
