// umma_probe.cu — stand-alone check of the tcgen05 descriptor conventions used by conv_umma.cuh:
//   * K-major SWIZZLE_NONE operands laid out [k-chunk][row][8 bf16] (SBO = 128 B, LBO = rows*16 B)
//   * row-shifted A start address (conv tap trick): start += shift*16 B for arbitrary shift
//   * bulk async copy of B + mbarrier, TMEM alloc/ld, commit
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../larynx_b200/csrc/umma.cuh"
using namespace b2::umma;

constexpr int M = 128, N = 64, KC = 4 /* k-chunks of 8 -> K = 32 */, R = 160;

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                             float* __restrict__ D, int shift) {
  extern __shared__ __align__(128) unsigned char smem[];
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem);                 // [KC][R][8]
  __nv_bfloat16* sB = sA + KC * R * 8;                                        // [KC][N][8]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + KC * N * 8);              // [0]=B landed, [1]=mma done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  // A: generic-proxy writes
  for (int i = tid; i < KC * R * 8 / 8; i += 128) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(A)[i];
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], KC * N * 16);
    bulk_g2s(sB, B, KC * N * 16, &bars[0]);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16_f32(M, N);
    for (int k = 0; k < KC / 2; ++k) {
      const uint64_t ad = make_smem_desc(smem_u32(sA) + (2 * k) * R * 16 + shift * 16, R * 16, 128);
      const uint64_t bd = make_smem_desc(smem_u32(sB) + (2 * k) * N * 16, N * 16, 128);
      mma_bf16_ss(tmem, ad, bd, idesc, k > 0);
    }
    mma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  uint32_t r[32];
  for (int c0 = 0; c0 < N; c0 += 32) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<float> a(R * KC * 8), b(N * KC * 8);
  srand(1);
  for (auto& v : a) v = (rand() % 2001 - 1000) / 1000.f;
  for (auto& v : b) v = (rand() % 2001 - 1000) / 1000.f;
  // logical A[row][k], B[n][k], k in [0, 32)
  std::vector<__nv_bfloat16> hA(KC * R * 8), hB(KC * N * 8);
  for (int kc = 0; kc < KC; ++kc)
    for (int r = 0; r < R; ++r)
      for (int e = 0; e < 8; ++e) hA[(kc * R + r) * 8 + e] = __float2bfloat16(a[r * KC * 8 + kc * 8 + e]);
  for (int kc = 0; kc < KC; ++kc)
    for (int n = 0; n < N; ++n)
      for (int e = 0; e < 8; ++e) hB[(kc * N + n) * 8 + e] = __float2bfloat16(b[n * KC * 8 + kc * 8 + e]);
  __nv_bfloat16 *dA, *dB;
  float* dD;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = KC * R * 16 + KC * N * 16 + 64;
  int fails = 0;
  for (int shift : {0, 8, 1, 3, 5, 13, 31}) {
    cudaMemset(dD, 0xff, M * N * 4);
    probe<<<1, 128, smem>>>(dA, dB, dD, shift);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shift %d: CUDA error %s\n", shift, cudaGetErrorString(e)); return 2; }
    std::vector<float> hD(M * N);
    cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < KC * 8; ++k)
          ref += (double)__bfloat162float(hA[((k / 8) * R + m + shift) * 8 + k % 8]) * (double)__bfloat162float(hB[((k / 8) * N + n) * 8 + k % 8]);
        maxerr = fmax(maxerr, fabs(ref - hD[m * N + n]));
      }
    printf("shift %2d: max |err| = %.3e %s\n", shift, maxerr, maxerr < 1e-4 ? "OK" : "MISMATCH");
    if (!(maxerr < 1e-4)) ++fails;
  }
  printf(fails ? "PROBE FAILED\n" : "PROBE PASSED\n");
  return fails ? 1 : 0;
}
