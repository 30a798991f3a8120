// Micro-benchmark of the tcgen05 + TMA skeleton of the tensor top-K kernel (k_topk_tc) without its epilogue: what feeds the
// tensor pipe fastest on B200?  One CTA per SM, 256 query rows (2 M-tiles of 128) resident, item tiles of 128 streamed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../collaborativefilteringusingtensorflow_b200/csrc \
//        -I ../../include -o mma_micro mma_micro.cu -lcuda && ./mma_micro
// Modes: 0 SS static   A and B in shared memory, no TMA traffic (pure MMA issue / operand-read rate)
//        1 SS + TMA    A in shared memory, B tiles streamed by TMA through a ring (= k_topk_tc today)
//        2 TS static   A in TENSOR memory (tcgen05.st once), B in shared memory, no TMA traffic
//        3 TS + TMA    A in tensor memory, B streamed by TMA (what k_topk_tc would become)
//        4 SS N=256 static   one MMA per 256 items (N = 256: the A operand is read once per 256 items)
//        5 SS N=256 + TMA    (two stages of 64 KB)
//        6..11 = 0..5 with the two M-tiles' MMAs INTERLEAVED (k outer, M-tile inner: consecutive instructions accumulate
//        into different TMEM tiles instead of forming one dependent chain)
// Also checks that the TS path computes the same accumulators as the SS path (TMEM layout of the A operand).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int KC = 2;                     // Kp = 128
constexpr int MT = 2;
constexpr int MAXST = 6;

__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

struct Params {
  int n_tiles, stages, mode, nB;          // nB: items per B tile (128 or 256)
  const __half* Q;                        // [256, 128]
  float* out;                             // [gridDim, 256, 128] accumulators of the LAST tile (check)
  long long* cycles;                      // [gridDim]
};

__global__ void __launch_bounds__(192, 1) k_mma(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmV,
                                                const __grid_constant__ CUtensorMap tmV256, const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bmode = P.mode % 6;
  const bool inter = P.mode >= 6;
  const bool ts = bmode == 2 || bmode == 3;
  const bool stream = bmode == 1 || bmode == 3 || bmode == 5;
  const int a_bytes = MT * KC * CHUNK_BYTES;
  const int b_stage_bytes = KC * CHUNK_BYTES * (P.nB / 128);
  uint8_t* sA = smem;
  uint8_t* sB = smem + a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)P.stages * b_stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAXST;
  uint64_t* a_full = bars + 2 * MAXST;
  uint64_t* done = a_full + 1;
  uint64_t* a_tm = done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_tm + 1);
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(a_full, 1);
    mbar_init(done, 1);
    mbar_init(a_tm, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM map: accumulators at columns [0, 384) (ring of 3 x 128; N = 256: [0, 512) with A overlapping is avoided by using
  // columns [0, 256) + A at [384, 512)), A operand (TS) at columns [384, 512): 2 M-tiles x 64 columns
  const uint32_t a_col = 384;
  const int n_static = P.stages;          // static modes cycle over the stages loaded once

  if (warp == 0 && lane == 0) {
    // ---- TMA producer
    mbar_arrive_expect_tx(a_full, (uint32_t)a_bytes);
    for (int mt = 0; mt < MT; ++mt)
      for (int kc = 0; kc < KC; ++kc) tma_load_2d(&tmQ, a_full, sA + (size_t)(mt * KC + kc) * CHUNK_BYTES, kc * KCH, mt * 128);
    int st = 0;
    uint32_t ph = 0u;
    const int n_load = stream ? P.n_tiles : n_static;
    for (int t = 0; t < n_load; ++t) {
      mbar_wait(empty + st, ph ^ 1u);
      mbar_arrive_expect_tx(full + st, (uint32_t)b_stage_bytes);
      for (int kc = 0; kc < KC; ++kc) {
        if (P.nB == 128) tma_load_2d(&tmV, full + st, sB + (size_t)st * b_stage_bytes + (size_t)kc * CHUNK_BYTES, kc * KCH, (t % 512) * 128);
        else tma_load_2d(&tmV256, full + st, sB + (size_t)st * b_stage_bytes + (size_t)kc * 2 * CHUNK_BYTES, kc * KCH, (t % 256) * 256);
      }
      if (++st == P.stages) { st = 0; ph ^= 1u; }
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer
    const uint32_t idesc = umma_idesc_fp16(128, P.nB);
    uint64_t a0[MT];
    for (int mt = 0; mt < MT; ++mt) a0[mt] = umma_desc_sw128(smem_u32(sA + (size_t)mt * KC * CHUNK_BYTES));
    const uint64_t b00 = umma_desc_sw128(smem_u32(sB));
    mbar_wait(a_full, 0u);
    if (ts) mbar_wait(a_tm, 0u);
    tc_fence_after();
    if (!stream) for (int s = 0; s < n_static; ++s) mbar_wait(full + s, 0u);
    const long long c0 = clock64();
    int st = 0;
    uint32_t ph = 0u;
    int slot = 0;
    for (int t = 0; t < P.n_tiles; ++t) {
      if (stream) {
        mbar_wait(full + st, ph);
        tc_fence_after();
      }
      const uint64_t b0 = b00 + (uint64_t)((st * b_stage_bytes) >> 4);
      uint32_t d_tm[MT];
      for (int mt = 0; mt < MT; ++mt) {
        if (P.nB == 128) {
          d_tm[mt] = tmem_base + (uint32_t)(slot * 128);        // ring of 3 accumulator slots
          if (++slot == 3) slot = 0;
        } else {
          d_tm[mt] = tmem_base + (uint32_t)(mt * 256);          // N = 256: one 256-column accumulator per M tile
        }
      }
      const int chunk_bytes = CHUNK_BYTES * (P.nB / 128);       // one K chunk of the B stage
      if (!inter) {
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int kc = 0; kc < KC; ++kc) {
#pragma unroll
            for (int k = 0; k < KCH / 16; ++k) {
              const uint64_t offa = (uint64_t)((kc * CHUNK_BYTES + k * 32) >> 4);
              const uint64_t offb = (uint64_t)((kc * chunk_bytes + k * 32) >> 4);
              if (ts) tc_mma_ts(d_tm[mt], tmem_base + a_col + (uint32_t)(mt * 64 + kc * 32 + k * 8), b0 + offb, idesc, (kc | k) ? 1u : 0u);
              else tc_mma_bf16(d_tm[mt], a0[mt] + offa, b0 + offb, idesc, (kc | k) ? 1u : 0u);
            }
          }
        }
      } else {
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
#pragma unroll
          for (int k = 0; k < KCH / 16; ++k) {
            const uint64_t offa = (uint64_t)((kc * CHUNK_BYTES + k * 32) >> 4);
            const uint64_t offb = (uint64_t)((kc * chunk_bytes + k * 32) >> 4);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              if (ts) tc_mma_ts(d_tm[mt], tmem_base + a_col + (uint32_t)(mt * 64 + kc * 32 + k * 8), b0 + offb, idesc, (kc | k) ? 1u : 0u);
              else tc_mma_bf16(d_tm[mt], a0[mt] + offa, b0 + offb, idesc, (kc | k) ? 1u : 0u);
            }
          }
        }
      }
      if (stream) tc_commit(empty + st);
      if (++st == P.stages) { st = 0; ph ^= 1u; }
    }
    tc_commit(done);
    mbar_wait(done, 0u);
    const long long c1 = clock64();
    P.cycles[blockIdx.x] = c1 - c0;
  } else if (warp >= 2) {
    // ---- warps 2..5: TMEM lane quadrant = warp % 4; load A into tensor memory (TS), read the accumulators back at the end
    const int q = warp & 3;
    if (ts) {
      for (int mt = 0; mt < MT; ++mt) {
        const __half* row = P.Q + (size_t)(mt * 128 + q * 32 + lane) * 128;
        for (int c8 = 0; c8 < 8; ++c8) {      // 8 x 8 columns of 2 halves = 128 halves
          uint32_t r[8];
          const uint4 lo = *reinterpret_cast<const uint4*>(row + c8 * 16), hi = *reinterpret_cast<const uint4*>(row + c8 * 16 + 8);
          r[0] = lo.x; r[1] = lo.y; r[2] = lo.z; r[3] = lo.w; r[4] = hi.x; r[5] = hi.y; r[6] = hi.z; r[7] = hi.w;
          tc_st8(tmem_base + ((uint32_t)(q * 32) << 16) + a_col + (uint32_t)(mt * 64 + c8 * 8), r);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      mbar_arrive(a_tm);
    }
    mbar_wait(done, 0u);
    tc_fence_after();
    if (P.out != nullptr && P.nB == 128) {
      // the LAST tile's accumulators: tile n_tiles-1, M-tile mt sits in ring slot (2 (n_tiles-1) + mt) % 3
      for (int mt = 0; mt < MT; ++mt) {
        const int slot = (2 * (P.n_tiles - 1) + mt) % 3;
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 128 + c * 32), r);
          tc_wait_ld();
          float* o = P.out + ((size_t)blockIdx.x * 256 + mt * 128 + q * 32 + lane) * 128 + c * 32;
          for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- TMEM read throughput: 8 warps (two per lane quadrant) read 32 lanes x NCOL columns per instruction, back to back
template <int NCOL>
__device__ __forceinline__ void ld_cols(uint32_t taddr, uint32_t& sink) {
  if constexpr (NCOL == 32) {
    uint32_t r[32];
    tc_ld32(taddr, r);
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) sink ^= r[j];
  } else if constexpr (NCOL == 16) {
    uint32_t r[16];
    tc_ld16(taddr, r);
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) sink ^= r[j];
  } else {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
        "%25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, "
        "%49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]),
          "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]),
          "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]),
          "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr)
        : "memory");
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 64; ++j) sink ^= r[j];
  }
}

template <int NCOL>
__global__ void __launch_bounds__(256, 1) k_ldtm(int iters, int nwarps, long long* cycles, uint32_t* out) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t sink = 0;
  __syncthreads();
  const long long c0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
      for (int c = 0; c < 512; c += NCOL) ld_cols<NCOL>(base + (uint32_t)c, sink);
    }
  }
  __syncthreads();
  const long long c1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = c1 - c0;
  out[blockIdx.x * 256 + threadIdx.x] = sink;
  (void)lane;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
  }
}

typedef CUresult (*encode_tiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_map(encode_tiled_t enc, CUtensorMap* tm, void* base, long long rows, int box_rows = 128) {
  const cuuint64_t dims[2] = {128, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {256};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
}

int main(int argc, char** argv) {
  const int n_tiles = argc > 1 ? atoi(argv[1]) : 20000;
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  encode_tiled_t enc = reinterpret_cast<encode_tiled_t>(fn);
  const long long v_rows = 512 * 128;                 // 16 MB of item tiles, L2-resident: isolates the SM-side feed
  std::vector<__half> hq(256 * 128), hv((size_t)v_rows * 128);
  srand(1);
  for (auto& x : hq) x = __float2half((rand() % 2001 - 1000) / 1000.f);
  for (auto& x : hv) x = __float2half((rand() % 2001 - 1000) / 1000.f);
  __half *dq, *dv;
  float* dout;
  long long* dcyc;
  CK(cudaMalloc(&dq, hq.size() * 2));
  CK(cudaMalloc(&dv, hv.size() * 2));
  CK(cudaMalloc(&dout, (size_t)sms * 256 * 128 * 4));
  CK(cudaMalloc(&dcyc, sms * 8));
  CK(cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, hv.data(), hv.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tmQ, tmV, tmV256;
  make_map(enc, &tmQ, dq, 256);
  make_map(enc, &tmV, dv, v_rows);
  make_map(enc, &tmV256, dv, v_rows, 256);
  CK(cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));

  // ---- correctness of the TS operand layout: same accumulators as SS on the same tiles (one tile, streamed)
  std::vector<float> ref((size_t)256 * 128), got((size_t)256 * 128);
  for (int mode : {1, 3}) {
    Params P = {3, 4, mode, 128, dq, dout, dcyc};
    const size_t smem = MT * KC * CHUNK_BYTES + 4 * KC * CHUNK_BYTES + 512 + 1024;
    k_mma<<<1, 192, smem>>>(tmQ, tmV, tmV256, P);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(mode == 1 ? ref.data() : got.data(), dout, ref.size() * 4, cudaMemcpyDeviceToHost));
  }
  double maxd = 0.0, maxv = 0.0;
  for (size_t i = 0; i < ref.size(); ++i) { maxd = fmax(maxd, fabs((double)ref[i] - got[i])); maxv = fmax(maxv, fabs((double)ref[i])); }
  // and against the host: tile 2 (the last of 3), rows of Q x rows 256..383 of V
  double maxh = 0.0;
  for (int r = 0; r < 256; r += 37)
    for (int c = 0; c < 128; c += 11) {
      double s = 0.0;
      for (int k = 0; k < 128; ++k) s += (double)__half2float(hq[r * 128 + k]) * (double)__half2float(hv[(size_t)(2 * 128 + c) * 128 + k]);
      maxh = fmax(maxh, fabs(s - ref[(size_t)r * 128 + c]));
    }
  printf("TS vs SS accumulators: max |diff| = %.3g (max |value| %.3g); SS vs host fp64: max |diff| = %.3g\n", maxd, maxv, maxh);

  {   // TMEM read throughput (no MMA running): bytes per SM clock
    uint32_t* dsink;
    CK(cudaMalloc(&dsink, (size_t)sms * 256 * 4));
    const int iters = 2000;
    for (int nw : {1, 4, 8}) {
      for (int ncol : {16, 32, 64}) {
        if (ncol == 16) k_ldtm<16><<<sms, 256>>>(iters, nw, dcyc, dsink);
        else if (ncol == 32) k_ldtm<32><<<sms, 256>>>(iters, nw, dcyc, dsink);
        else k_ldtm<64><<<sms, 256>>>(iters, nw, dcyc, dsink);
        CK(cudaDeviceSynchronize());
        std::vector<long long> cyc(sms);
        CK(cudaMemcpy(cyc.data(), dcyc, sms * 8, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (auto c : cyc) avg += c;
        avg /= sms;
        const double bytes = (double)nw * iters * 512.0 * 32 * 4;
        printf("LDTM 32x32b.x%-2d  %d warps: %7.1f bytes per cycle per SM (%.1f cycles per instruction per warp)\n", ncol, nw, bytes / avg,
               avg / (iters * (512.0 / ncol)));
      }
    }
  }
  const char* names[6] = {"SS static", "SS + TMA ring", "TS static (A in TMEM)", "TS + TMA ring", "SS N=256 static", "SS N=256 + TMA ring"};
  for (int mode = 0; mode < 12; ++mode) {
    for (int stages : {2, 4, 6}) {
      const int nB = (mode % 6) >= 4 ? 256 : 128;
      if ((nB == 256) != (stages == 2)) continue;
      const size_t smem = MT * KC * CHUNK_BYTES + (size_t)stages * KC * CHUNK_BYTES * (nB / 128) + 512 + 1024;
      if (smem > 227 * 1024) continue;
      Params P = {n_tiles / (nB / 128), stages, mode, nB, dq, nullptr, dcyc};
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k_mma<<<sms, 192, smem>>>(tmQ, tmV, tmV256, P);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms < best) best = ms;
      }
      CK(cudaGetLastError());
      std::vector<long long> cyc(sms);
      CK(cudaMemcpy(cyc.data(), dcyc, sms * 8, cudaMemcpyDeviceToHost));
      double avg = 0;
      for (auto c : cyc) avg += c;
      avg /= sms;
      const double flops = (double)sms * n_tiles * 2.0 * 256 * 128 * 128;
      printf("mode %2d %-22s%s stages %d: %8.3f ms  %7.1f TFLOP/s  %6.1f cycles per 128x128x16 MMA (SM clock)\n", mode, names[mode % 6], mode >= 6 ? " interleaved" : "            ", stages,
             best, flops / (best * 1e-3) / 1e12, avg / ((double)n_tiles * 16));
    }
  }
  return 0;
}
