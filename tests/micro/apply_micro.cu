// Micro-benchmark of the staged-apply access pattern (why is k_apply_staged 5x slower than its traffic predicts?)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o apply_micro apply_micro.cu && ./apply_micro
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct P {
  float *tab, *acc, *stg; const unsigned* slot_row; unsigned* meta; int n; int lds; int* counter;
};

template <int VAR>
__global__ void __launch_bounds__(256) k(P p) {
  const int lane = threadIdx.x & 31;
  const long long ng = (long long)gridDim.x * blockDim.x / 32;
  int n = p.n;
  if (VAR == 7) n = min(__ldcg(p.counter), p.n);
  for (long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 32; s < n; s += ng) {
    unsigned code = (VAR == 6) ? __ldg(p.slot_row + s) : __ldcg(p.slot_row + s);
    long long r = code & 0x7fffffffu;
    float4 g = make_float4(0, 0, 0, 0);
    float* st = p.stg + s * p.lds + 4 * lane;
    if (VAR != 3) { g = __ldcg((const float4*)st); if (VAR != 2 && VAR != 8) __stcg((float4*)st, make_float4(0, 0, 0, 0)); }
    float* pr = p.tab + r * 128 + 4 * lane;
    float* ar = p.acc + r * 128 + 4 * lane;
    float4 c, a;
    if (VAR == 4) { c = *(const float4*)pr; a = *(const float4*)ar; }
    else { c = __ldcg((const float4*)pr); a = __ldcg((const float4*)ar); }
    a.x += g.x * g.x; a.y += g.y * g.y; a.z += g.z * g.z; a.w += g.w * g.w;
    c.x -= 0.1f * g.x * rsqrtf(a.x); c.y -= 0.1f * g.y * rsqrtf(a.y); c.z -= 0.1f * g.z * rsqrtf(a.z); c.w -= 0.1f * g.w * rsqrtf(a.w);
    if (VAR != 5) { __stcg((float4*)ar, a); __stcg((float4*)pr, c); }
    if (lane == 0 && VAR != 5) __stcg(p.meta + r, 0u);
    if (VAR == 8) __stcg((float4*)st, make_float4(0.f * c.x, 0, 0, 0));   // zero the slot only after its data was consumed
  }
}

int main() {
  const long long rows = 1500000; const int n = 100000;
  P p; int lds_list[2] = {132, 128};
  CK(cudaMalloc(&p.tab, rows * 128 * 4)); CK(cudaMalloc(&p.acc, rows * 128 * 4));
  CK(cudaMalloc(&p.stg, (size_t)n * 132 * 4)); CK(cudaMalloc(&p.meta, rows * 4)); CK(cudaMalloc(&p.counter, 4));
  CK(cudaMemset(p.tab, 0, rows * 128 * 4)); CK(cudaMemset(p.acc, 0x3f, rows * 128 * 4)); CK(cudaMemset(p.stg, 0, (size_t)n * 132 * 4));
  CK(cudaMemcpy(p.counter, &n, 4, cudaMemcpyHostToDevice));
  std::vector<unsigned> h(rows); for (long long i = 0; i < rows; ++i) h[i] = (unsigned)i;
  std::mt19937 rng(1); std::shuffle(h.begin(), h.end(), rng);
  unsigned* d; CK(cudaMalloc(&d, n * 4)); CK(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice));
  p.slot_row = d; p.n = n;
  float* flush; CK(cudaMalloc(&flush, 512 << 20));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[] = {"V0 current (lds=132)", "V1 aligned staging (lds=128)", "V2 no zeroing", "V3 no staging access", "V4 plain ld for param/acc",
                         "V5 loads only (no stores)", "V6 ldg slot_row", "V7 n from device counter", "V8 zero slot after use"};
  for (int grid : {1184, 148 * 32, 12500}) {
    for (int var : {0, 2, 8}) {
      p.lds = (var == 1) ? 128 : 132;
      float best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        CK(cudaMemsetAsync(flush, rep, 512 << 20));
        CK(cudaEventRecord(e0));
        switch (var) {
          case 0: k<0><<<grid, 256>>>(p); break; case 1: k<1><<<grid, 256>>>(p); break; case 2: k<2><<<grid, 256>>>(p); break;
          case 3: k<3><<<grid, 256>>>(p); break; case 4: k<4><<<grid, 256>>>(p); break; case 5: k<5><<<grid, 256>>>(p); break;
          case 6: k<6><<<grid, 256>>>(p); break; case 7: k<7><<<grid, 256>>>(p); break; default: k<8><<<grid, 256>>>(p); break;
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
      }
      printf("grid %5d  %-34s %8.1f us  (%.0f GB/s of 3.1 KB/row)\n", grid, names[var], best * 1e3, n * 3100.0 / (best * 1e-3) / 1e9);
    }
  }
  return 0;
}
