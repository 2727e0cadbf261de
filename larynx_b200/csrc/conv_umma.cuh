// conv_umma.cuh — Conv1d as an implicit GEMM on the 5th-gen tensor cores (tcgen05, accumulators in TMEM)
// with fp32-class accuracy from a bf16 split:  x = hi + lo (+ lo2),  D += Ahi*Bhi + Alo*Bhi + Ahi*Blo (+ ...)
//
//   D[t][co] = sum_k sum_ci  X[ci][t + k*dil - pad_left] * W[co][ci][k]
//   M = time (128 rows per MMA, MT tiles per CTA), N = output channels (NT <= 256 per CTA), K = input channels
//
// One CTA:
//   warps 0..7  : A producers — read fp32 activations [C][L] from global (coalesced along time), apply the fused
//                 pre-op (MRF mean of 3 inputs, LeakyReLU, zero padding by valid length), split to bf16 hi/lo and
//                 write the UMMA operand image [k-chunk][row][8] (K-major, SWIZZLE_NONE, SBO = 128 B) to smem;
//                 after the main loop the same warps run the epilogue: TMEM -> registers (tcgen05.ld 32x32b) ->
//                 bias / act / residual / gate / WN res-skip / coupling+invconv+actnorm / pixel-shuffle -> global.
//   warp 8      : MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::f16, tap k = same A image with the
//                 descriptor start address advanced by k*dil rows (16 B each) — no im2col, no per-tap copies.
//   warp 9      : weight producer: one cp.async.bulk (TMA engine) per (ci-chunk, tap) stage from the pre-packed bf16
//                 hi/lo image, completion on an mbarrier.
// Pipelines: A ring (2 stages, per ci-chunk of 32 channels, reused by all K taps), W ring (4 stages, per tap),
// accumulator barrier (MMA -> epilogue).  All waits are mbarrier based; tcgen05.commit releases stages.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace b2 {

constexpr int UM_BKC = 32;               // input channels per A stage
constexpr int UM_KCH = UM_BKC / 8;       // 16-byte k-chunks per stage
constexpr int UM_HALO = 80;              // max (K-1)*dil
constexpr int UM_WSTAGES = 3;
constexpr int UM_ASTAGES = 2;
constexpr int UM_NPROD = 256;            // producer / epilogue threads (8 warps)
constexpr int UM_NT_THREADS = UM_NPROD + 64;

template <int MT>
struct UmmaCfg {
  static constexpr int RA = MT * 128 + UM_HALO;  // rows of the A image
  static size_t a_stage_bytes(int nsplit) { return (size_t)nsplit * UM_KCH * RA * 16; }
  static size_t w_stage_bytes(int nsplit, int NT) { return (size_t)nsplit * UM_KCH * NT * 16; }
  static size_t smem_bytes(int nsplit, int NT) {
    return UM_ASTAGES * a_stage_bytes(nsplit) + UM_WSTAGES * w_stage_bytes(nsplit, NT) + 128 + 4 * (256 + 16 + 2 * 256);
  }
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define UM_TRACE(slot)                                                                         \
  do {                                                                                         \
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.dbg[slot] = gtimer(); \
  } while (0)

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

template <int MT, int NSPLIT>
__global__ void __launch_bounds__(UM_NT_THREADS, MT == 1 ? 2 : 1) conv_umma_kernel(const ConvParams p) {
  using namespace umma;
  using C = UmmaCfg<MT>;
  constexpr int RA = C::RA;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int NT = p.um_nt;
  const uint32_t a_split_bytes = UM_KCH * RA * 16;          // one split plane (hi or lo) of an A stage
  const uint32_t a_stage = NSPLIT * a_split_bytes;
  const uint32_t w_split_bytes = UM_KCH * NT * 16;
  const uint32_t w_stage = NSPLIT * w_split_bytes;
  unsigned char* sA = smem_raw;
  unsigned char* sW = smem_raw + UM_ASTAGES * a_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + UM_WSTAGES * w_stage);
  uint64_t* a_full = bars;                      // [2]
  uint64_t* a_empty = bars + 2;                 // [2]
  uint64_t* w_full = bars + 4;                  // [UM_WSTAGES]
  uint64_t* w_empty = bars + 8;                 // [UM_WSTAGES]
  uint64_t* acc_full = bars + 12;               // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  float* s_bias = reinterpret_cast<float*>(bars + 16);  // [NT] bias of this N tile (+ [16+2*half] coupling constants)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * (MT * 128);
  const int ntile = blockIdx.y;
  const int n0 = ntile * NT;
  const int Lb = p.lens ? p.lens[b] * p.len_mul : p.Lin;
  if (t0 >= Lb) return;
  const int K = p.K, dil = p.dil;
  const int nch = (p.Cin + UM_BKC - 1) / UM_BKC;
  uint32_t tmem_cols = 32;
  // one fp32 accumulator per magnitude class of split products (hi*hi | hi*lo+lo*hi | ...): the tensor core
  // truncates when it adds into the accumulator, so the long hi*hi chain must not absorb the small terms
  constexpr int NACC = NSPLIT;
  while (tmem_cols < (uint32_t)(NACC * MT * NT)) tmem_cols <<= 1;

  if (tid == 0) UM_TRACE(0);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], UM_NPROD / 32);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < UM_WSTAGES; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, tmem_cols);
  // epilogue constants -> smem (global loads in the epilogue would serialise on L2 latency)
  for (int i = tid; i < NT; i += UM_NT_THREADS) s_bias[i] = (n0 + i) < p.Cout ? p.bias[n0 + i] : 0.f;
  float* s_inv = s_bias + 256;        // [16]
  float* s_anb = s_inv + 16;          // [2*half] ActNorm bias
  float* s_ans = s_anb + 256;         // [2*half] ActNorm exp(-logs)
  if (p.epi == EPI_COUPLING) {
    if (tid < 16) s_inv[tid] = p.inv_w[tid];
    for (int i = tid; i < 2 * p.half; i += UM_NT_THREADS) {
      s_anb[i] = p.an_bias[i];
      s_ans[i] = p.an_scale[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) UM_TRACE(1);

  if (warp == 9) {
    // ===================== weight producer =====================
    if (lane == 0) {
      const unsigned char* gw = reinterpret_cast<const unsigned char*>(p.wu) + (size_t)ntile * nch * K * w_stage;
      int n = 0;
      for (int c = 0; c < nch; ++c)
        for (int k = 0; k < K; ++k, ++n) {
          const int s = n % UM_WSTAGES;
          mbar_wait(&w_empty[s], ((n / UM_WSTAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&w_full[s], w_stage);
          bulk_g2s(sW + (size_t)s * w_stage, gw + (size_t)n * w_stage, w_stage, &w_full[s]);
        }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16_f32(128, NT);
      const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(sW);
      int n = 0;
      for (int c = 0; c < nch; ++c) {
        const int as = c & 1;
        mbar_wait(&a_full[as], (c >> 1) & 1);
        tc_fence_after();
        if (c == 0) UM_TRACE(3);
        for (int k = 0; k < K; ++k, ++n) {
          const int s = n % UM_WSTAGES;
          mbar_wait(&w_full[s], (n / UM_WSTAGES) & 1);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < UM_BKC / 16; ++ks) {
            const uint32_t wb = sW_u + s * w_stage + (2 * ks) * NT * 16;
            uint64_t bd[NSPLIT];
#pragma unroll
            for (int sp = 0; sp < NSPLIT; ++sp) bd[sp] = make_smem_desc(wb + sp * w_split_bytes, NT * 16, 128);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint32_t ab = sA_u + as * a_stage + (2 * ks) * RA * 16 + (mt * 128 + k * dil) * 16;
              uint64_t ad[NSPLIT];
#pragma unroll
              for (int sp = 0; sp < NSPLIT; ++sp) ad[sp] = make_smem_desc(ab + sp * a_split_bytes, RA * 16, 128);
              const uint32_t d0 = tmem + (0 * MT + mt) * NT, d1 = tmem + (1 * MT + mt) * NT;
              const uint32_t first = (c == 0 && k == 0 && ks == 0) ? 0u : 1u;
              mma_bf16_ss(d0, ad[0], bd[0], idesc, first);   // hi*hi
              mma_bf16_ss(d1, ad[0], bd[1], idesc, first);   // hi*lo
              mma_bf16_ss(d1, ad[1], bd[0], idesc, 1u);      // lo*hi
              if (NSPLIT == 3) {
                const uint32_t d2 = tmem + (2 * MT + mt) * NT;
                mma_bf16_ss(d2, ad[1], bd[1], idesc, first); // lo*lo
                mma_bf16_ss(d2, ad[0], bd[2], idesc, 1u);    // hi*lo2
                mma_bf16_ss(d2, ad[2], bd[0], idesc, 1u);    // lo2*hi
              }
            }
          }
          mma_commit(&w_empty[s]);  // frees the weight stage when these MMAs have completed
        }
        mma_commit(&a_empty[as]);   // frees the activation stage
      }
      mma_commit(acc_full);
      UM_TRACE(4);
    }
  } else {
    // ===================== A producers (warps 0..7) =====================
    const float* xb = p.x + (size_t)b * p.x_bs;
    const float* xb2 = p.x2 ? p.x2 + (size_t)b * p.x_bs : nullptr;
    const float* xb3 = p.x3 ? p.x3 + (size_t)b * p.x_bs : nullptr;
    const int rows_used = MT * 128 + (K - 1) * dil;
    const bool three = xb2 != nullptr && xb3 != nullptr;  // MRF mean of three inputs (uniform)
    for (int c = 0; c < nch; ++c) {
      const int as = c & 1;
      mbar_wait(&a_empty[as], ((c >> 1) & 1) ^ 1);
      unsigned char* dst = sA + (size_t)as * a_stage;
      constexpr int NTASK = (UM_KCH * RA + UM_NPROD - 1) / UM_NPROD;
      constexpr int TB = MT == 1 ? 4 : 3;  // tasks per batch: TB x 8 independent global loads in flight per thread
#pragma unroll 1
      for (int i0 = 0; i0 < NTASK; i0 += TB) {
        float v[TB][8];
        int so[TB];
        // branch-free gather: every load is unconditional (clamped to a safe address) and masked afterwards,
        // so all TB*8 loads of a batch are in flight together
#pragma unroll
        for (int i = 0; i < TB; ++i) {
          const int task = tid + (i0 + i) * UM_NPROD;
          const int kc = task / RA, r = task - kc * RA;
          const int tin = t0 - p.pad_left + r;
          const bool row_ok = task < UM_KCH * RA && r < rows_used && tin >= 0 && tin < Lb;
          const int ci0 = c * UM_BKC + kc * 8;
          so[i] = task < UM_KCH * RA ? (kc * RA + r) * 16 : -1;
          const int tin_c = row_ok ? tin : 0;
          if (!three) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const bool ok = row_ok && (ci0 + e) < p.Cin;
              const size_t off = ok ? (size_t)(ci0 + e) * p.x_ld + tin_c : 0;
              const float a = __ldg(xb + off);
              v[i][e] = ok ? a : 0.f;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const bool ok = row_ok && (ci0 + e) < p.Cin;
              const size_t off = ok ? (size_t)(ci0 + e) * p.x_ld + tin_c : 0;
              const float a = __ldg(xb + off), a2 = __ldg(xb2 + off), a3 = __ldg(xb3 + off);
              v[i][e] = ok ? (a + a2) + a3 : 0.f;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < TB; ++i) {
          if (so[i] < 0) continue;
          __nv_bfloat16 h[8], l[8], l2[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float a = v[i][e];
            if (p.in_div != 1.f) a = a / p.in_div;
            if (p.pre == PRE_LRELU) a = a > 0.f ? a : a * p.pre_slope;
            h[e] = __float2bfloat16_rn(a);
            const float r1 = a - __bfloat162float(h[e]);
            l[e] = __float2bfloat16_rn(r1);
            if (NSPLIT == 3) l2[e] = __float2bfloat16_rn(r1 - __bfloat162float(l[e]));
          }
          unsigned char* o = dst + so[i];
          *reinterpret_cast<uint4*>(o) =
              make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
          *reinterpret_cast<uint4*>(o + a_split_bytes) =
              make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
          if (NSPLIT == 3)
            *reinterpret_cast<uint4*>(o + 2 * a_split_bytes) = make_uint4(
                pack_bf16x2(l2[0], l2[1]), pack_bf16x2(l2[2], l2[3]), pack_bf16x2(l2[4], l2[5]), pack_bf16x2(l2[6], l2[7]));
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[as]);
      if (tid == 0 && c == 0) UM_TRACE(2);
      if (tid == 0 && c == nch - 1) UM_TRACE(5);
    }

    // ===================== epilogue (same warps) =====================
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (tid == 0) UM_TRACE(6);
    const int quad = warp & 3;
    const int grp = warp >> 2;  // MT == 2: the M tile; MT == 1: which half of the N columns
    const int mt = MT == 2 ? grp : 0;
    const int m = quad * 32 + lane;
    const int t = t0 + mt * 128 + m;
    const bool t_ok = t < Lb;
    const int ngroups = (NT + 31) / 32;
    const int g_lo = MT == 2 ? 0 : (grp == 0 ? 0 : (ngroups + 1) / 2);
    const int g_hi = MT == 2 ? ngroups : (grp == 0 ? (ngroups + 1) / 2 : ngroups);
    for (int g = g_lo; g < g_hi; ++g) {
      float acc[32];
      {
        uint32_t raw[32];
        const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * 32);
        // smallest magnitude class first
        tmem_ld32(lane_addr + (uint32_t)(((NACC - 1) * MT + mt) * NT), raw);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(raw[j]);
#pragma unroll
        for (int a = NACC - 2; a >= 0; --a) {
          tmem_ld32(lane_addr + (uint32_t)((a * MT + mt) * NT), raw);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(raw[j]);
        }
      }
      const int nb = n0 + g * 32;  // packed channel of acc[0]
      const int jmax = min(32, min(NT - g * 32, p.Cout - nb));
      if (!t_ok || jmax <= 0) continue;
      if (p.epi == EPI_PLAIN && p.pix <= 1) {
        float* __restrict__ yb = p.y + (size_t)b * p.y_bs + t;
        float rv[32];
        if (p.res) {
          const float* __restrict__ rb = p.res + (size_t)b * p.res_bs + t;
#pragma unroll
          for (int j = 0; j < 32; ++j) rv[j] = rb[(size_t)(nb + (j < jmax ? j : 0)) * p.res_ld];
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) rv[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < jmax) {
            float v = acc[j] + s_bias[g * 32 + j];
            if (p.act == ACT_RELU) v = fmaxf(v, 0.f);
            else if (p.act == ACT_TANH) v = tanhf(v);
            yb[(size_t)(nb + j) * p.y_ld] = v + rv[j];
          }
        }
      } else if (p.epi == EPI_PLAIN) {
        // pixel shuffle (transposed conv): packed channel n = co*u + r -> y[co][t*u + r]
        float* __restrict__ yb = p.y + (size_t)b * p.y_bs;
        const int u = p.pix;
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += j < jmax ? s_bias[g * 32 + j] : 0.f;
        if (u == 8 && jmax == 32) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4* d = reinterpret_cast<float4*>(yb + (size_t)(nb / 8 + q) * p.y_ld + (size_t)t * 8);
            d[0] = make_float4(acc[q * 8 + 0], acc[q * 8 + 1], acc[q * 8 + 2], acc[q * 8 + 3]);
            d[1] = make_float4(acc[q * 8 + 4], acc[q * 8 + 5], acc[q * 8 + 6], acc[q * 8 + 7]);
          }
        } else if (u == 4 && jmax == 32) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(yb + (size_t)(nb / 4 + q) * p.y_ld + (size_t)t * 4) =
                make_float4(acc[q * 4 + 0], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
        } else if (u == 2 && jmax == 32) {
#pragma unroll
          for (int q = 0; q < 16; ++q)
            *reinterpret_cast<float2*>(yb + (size_t)(nb / 2 + q) * p.y_ld + (size_t)t * 2) = make_float2(acc[q * 2], acc[q * 2 + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < jmax) {
              const int n = nb + j, co = n / u, r = n - co * u;
              yb[(size_t)co * p.y_ld + (size_t)t * u + r] = acc[j];
            }
          }
        }
      } else if (p.epi == EPI_GATE) {
        float* __restrict__ yb = p.y + (size_t)b * p.y_bs + t;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          if (j < jmax) {
            const float a = acc[j] + s_bias[g * 32 + j];
            const float gt = acc[j + 1] + s_bias[g * 32 + j + 1];
            yb[(size_t)((nb + j) >> 1) * p.y_ld] = tanhf(a) * (1.f / (1.f + expf(-gt)));
          }
        }
      } else if (p.epi == EPI_RESSKIP) {
        float* yb = p.y + (size_t)b * p.y_bs + t;
        float* ab = p.aux + (size_t)b * p.aux_bs + t;
        float old[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {  // branch-free so the 32 loads are in flight together
          const int co = nb + j;
          const bool in_y = co < p.split;
          const bool need = j < jmax && (in_y || !p.first);
          const float* src = in_y ? yb + (size_t)co * p.y_ld : ab + (size_t)(co - p.split) * p.aux_ld;
          const float o = *(need ? src : yb);
          old[j] = need ? o : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int co = nb + j;
          if (j < jmax) {
            const float v = old[j] + (acc[j] + s_bias[g * 32 + j]);
            if (co < p.split) yb[(size_t)co * p.y_ld] = v;
            else ab[(size_t)(co - p.split) * p.aux_ld] = v;
          }
        }
      } else {  // EPI_COUPLING (+ InvConvNear^-1 + ActNorm^-1), see conv1d.cuh for the channel algebra
        float* sb = p.y + (size_t)b * p.y_bs + t;
        const int half = p.half;
        float xin[32];  // per group of 4 packed channels: x0[2q], x0[2q+1], x1[2q], x1[2q+1]
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int i0 = j < jmax ? (nb + j) >> 1 : 0;
          xin[j + 0] = sb[(size_t)i0 * p.y_ld];
          xin[j + 1] = sb[(size_t)(i0 + 1) * p.y_ld];
          xin[j + 2] = sb[(size_t)(half + i0) * p.y_ld];
          xin[j + 3] = sb[(size_t)(half + i0 + 1) * p.y_ld];
        }
        float wi[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) wi[o] = s_inv[o];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (j < jmax) {
            const int co = nb + j;
            const int i0 = co >> 1;
            const int ch[4] = {i0, i0 + 1, half + i0, half + i0 + 1};
            float vin[4];
            vin[0] = xin[j + 0];
            vin[1] = xin[j + 1];
            vin[2] = (xin[j + 2] - (acc[j] + s_bias[g * 32 + j])) * expf(-(acc[j + 1] + s_bias[g * 32 + j + 1]));
            vin[3] = (xin[j + 3] - (acc[j + 2] + s_bias[g * 32 + j + 2])) * expf(-(acc[j + 3] + s_bias[g * 32 + j + 3]));
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              float z = wi[o * 4 + 0] * vin[0];
              z = fmaf(wi[o * 4 + 1], vin[1], z);
              z = fmaf(wi[o * 4 + 2], vin[2], z);
              z = fmaf(wi[o * 4 + 3], vin[3], z);
              sb[(size_t)ch[o] * p.y_ld] = (z - s_anb[ch[o]]) * s_ans[ch[o]];
            }
          }
        }
      }
    }
  }
  if (tid == 0) UM_TRACE(7);
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, tmem_cols);
  if (tid == 0) UM_TRACE(8);
}

}  // namespace b2
