// Micro-benchmark of NVLink peer-memory access patterns used (or considered) by the multi-GPU exchange: 512-byte rows
// read / written / red.added on ANOTHER GPU's memory from a kernel, random vs sequential rows.  One process, two GPUs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peer_micro peer_micro.cu && ./peer_micro [rows=2000000]
// Answers for DESIGN.md section 10: is pushing gradients with remote red.add (no compact buffer, no owner-side pass)
// competitive with the owner-pull of rows (remote LDG.128), and how much does random vs sequential row order matter.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// MODE 0: read rows (ld.global.cg 16 B per lane) and reduce into a local sink; 1: store rows; 2: red.add.v4.f32 rows
template <int MODE>
__global__ void __launch_bounds__(256) k_rows(float* remote, const int* __restrict__ ids, long long n, float* sink) {
  const int lane = threadIdx.x & 31;
  const long long ng = (long long)gridDim.x * blockDim.x / 32;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long k = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 32; k < n; k += ng) {
    float4* p = reinterpret_cast<float4*>(remote + (long long)__ldg(ids + k) * 128) + lane;
    if (MODE == 0) {
      const float4 v = __ldcg(p);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    } else if (MODE == 1) {
      __stcg(p, make_float4(1.f, 2.f, 3.f, (float)k));
    } else {
      atomicAdd(p, make_float4(1.f, 1.f, 1.f, 1.f));
    }
  }
  if (MODE == 0 && acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
}

template <int MODE>
float run(float* buf, const int* ids, long long n, float* sink, int sms) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    k_rows<MODE><<<sms * 8, 256>>>(buf, ids, n, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 2000000;        // rows touched per launch (1 GB at 512 B)
  const long long table_rows = 4 * n;                              // 4 GB table: rows rarely repeat
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  std::vector<int> h_rand(n), h_seq(n);
  std::mt19937_64 g(1);
  for (long long k = 0; k < n; ++k) { h_rand[k] = (int)(g() % table_rows); h_seq[k] = (int)k; }
  for (int where = 0; where < (ndev > 1 ? 2 : 1); ++where) {       // 0: local memory (baseline), 1: the peer's memory
    float* buf;
    CK(cudaSetDevice(where));
    CK(cudaMalloc(&buf, (size_t)table_rows * 512));
    CK(cudaMemset(buf, 0, (size_t)table_rows * 512));
    CK(cudaSetDevice(0));
    if (where == 1) {
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, 0, 1));
      if (!can) { printf("GPU 0 cannot access GPU 1\n"); return 0; }
      CK(cudaDeviceEnablePeerAccess(1, 0));
    }
    int *d_rand, *d_seq;
    float* sink;
    CK(cudaMalloc(&d_rand, n * 4));
    CK(cudaMalloc(&d_seq, n * 4));
    CK(cudaMalloc(&sink, 16));
    CK(cudaMemcpy(d_rand, h_rand.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_seq, h_seq.data(), n * 4, cudaMemcpyHostToDevice));
    const double gb = (double)n * 512 / 1e9;
    const char* names[3] = {"read  (ld.cg 128-bit)", "store (st.cg 128-bit)", "red.add.v4.f32"};
    for (int order = 0; order < 2; ++order) {
      const int* ids = order ? d_seq : d_rand;
      const float t0 = run<0>(buf, ids, n, sink, sms), t1 = run<1>(buf, ids, n, sink, sms), t2 = run<2>(buf, ids, n, sink, sms);
      const float t[3] = {t0, t1, t2};
      for (int m = 0; m < 3; ++m)
        printf("%-6s %-10s rows  %-22s %8.3f ms  %7.1f GB/s\n", where ? "peer" : "local", order ? "sequential" : "random", names[m],
               t[m], gb / (t[m] * 1e-3));
    }
    CK(cudaFree(d_rand)); CK(cudaFree(d_seq)); CK(cudaFree(sink));
    CK(cudaSetDevice(where));
    CK(cudaFree(buf));
    CK(cudaSetDevice(0));
  }
  return 0;
}
