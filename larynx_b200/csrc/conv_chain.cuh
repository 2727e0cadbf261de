// conv_chain.cuh — a CHAIN of dependent convs in ONE cooperative launch (GlowTTS flow decoder at small batch).
//
// STATUS: written at the end of round 1, compiled, NOT yet run on a GPU (the round's GPU budget was spent); it is only
// reachable with B200TTS_FLOW_CHAIN=1 and is the starting point of round 2.
//
// Why: at batch 1 the flow decoder is 120 dependent conv launches of 6-24 CTAs each; a launch costs ~12 us of which
// ~7 us are launch gap, TMEM allocation, barrier init and pipeline fill/drain (profiles/r1_tcgen05_summary.md).  Here the
// CTAs stay resident: grid = (time tiles, max n-tiles, batch); CTA (x, y, z) computes tile (x, y) of item z in every
// layer that has such a tile; layers are separated by a grid-wide barrier, one counter per layer (release: epilogue ->
// atomicAdd, acquire: the activation-copy lane before it loads the next layer's A tiles).  The weight-copy lane does not wait for
// the barrier — weights are static — so the weight stream of layer l+1 is in flight while layer l is in its epilogue.
//
// The conv itself is the plane-mode path of conv_umma_kernel<1, 2, EW, WST, AST, G, 1> (same operand images, same MMA
// sequence, the very same epilogue text: conv_umma_epilogue.inc).  Restrictions (checked by the launcher): every layer
// reads and writes operand images (no fp32 gather), 2-way split, NT <= 64, one 128-row tile per CTA and layer.
#pragma once
#include "conv_umma.cuh"

namespace b2 {

struct ChainParams {
  const ConvParams* layers;  // device array, n_layers entries, in execution order
  int n_layers;
  unsigned int* counter;     // [n_layers], zeroed before the launch; counter[l] += 1 per CTA that is done with layer l
  unsigned int n_ctas;       // gridDim.x * gridDim.y * gridDim.z
};

constexpr int CH_NT_MAX = 64;  // widest N tile of a chained layer (weight ring slots and TMEM are sized for it)

template <int EW, int WST, int AST, int G>
struct ChainCfg {
  static constexpr int RA = 128 + UM_HALO;
  static constexpr uint32_t a_split_bytes = UM_KCH * RA * 16;
  static constexpr uint32_t a_stage = 2 * a_split_bytes;
  static constexpr uint32_t w_slot = (uint32_t)G * 2 * UM_KCH * CH_NT_MAX * 16;  // G taps at the widest N tile
  static constexpr size_t smem_bytes = (size_t)AST * a_stage + (size_t)WST * w_slot + 384 + 4 * (256 + 16 + 2 * 256);
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// epilogue warps only (ids 0..EW-1): named barrier 1
template <int EW>
__device__ __forceinline__ void epi_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
}

template <int EW, int WST, int AST, int G>
__global__ void __launch_bounds__(EW * 32 + 96, 1) conv_chain_kernel(const __grid_constant__ ChainParams cp) {
  using namespace umma;
  using CC = ChainCfg<EW, WST, AST, G>;
  constexpr int MT = 1, NSPLIT = 2, NACC = 2;
  constexpr int RA = CC::RA;
  constexpr int UM_NPROD = EW * 32;
  constexpr int W_MMA = EW, W_LOAD = EW + 1, W_ALOAD = EW + 2;
  constexpr uint32_t a_split_bytes = CC::a_split_bytes, a_stage = CC::a_stage, w_slot = CC::w_slot;
  static_assert(WST <= UM_WSTAGES_MAX && AST <= 4, "ring depths");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sA = smem_raw;
  unsigned char* sW = smem_raw + AST * a_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + WST * w_slot);
  uint64_t* a_full = bars;          // [AST]
  uint64_t* a_empty = bars + 4;     // [AST]
  uint64_t* w_full = bars + 8;      // [WST]
  uint64_t* w_empty = bars + 24;    // [WST]
  uint64_t* acc_full = bars + 40;   // MMA -> epilogue
  uint64_t* acc_empty = bars + 42;  // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 44);
  float* s_bias = reinterpret_cast<float*>(bars + 46);
  float* s_inv = s_bias + 256;
  float* s_anb = s_inv + 16;
  float* s_ans = s_anb + 256;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, mtile = blockIdx.x, ntile = blockIdx.y;
  const int t0 = mtile * 128;
  constexpr uint32_t tmem_cols = 128;  // NACC * CH_NT_MAX

  if (tid == 0) {
    for (int i = 0; i < AST; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < WST; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, EW);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // a CTA takes part in layer l iff the layer has its (time tile, n-tile); every CTA takes part in every barrier
  auto active = [&](const ConvParams& p) {
    const int Lb = p.lens ? p.lens[b] * p.len_mul : p.Lin;
    return t0 < Lb && ntile * p.um_nt < p.Cout;
  };

  if (warp == W_LOAD) {
    // ===================== weight producer: runs ahead of the layer barrier (weights are static) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int l = 0; l < cp.n_layers; ++l) {
        const ConvParams& p = cp.layers[l];
        if (!active(p)) continue;
        const int NT = p.um_nt, K = p.K, nch = (p.Cin + UM_BKC - 1) / UM_BKC;
        const uint32_t w_stage = (uint32_t)NSPLIT * UM_KCH * NT * 16;
        const unsigned char* gw = reinterpret_cast<const unsigned char*>(p.wu) + (size_t)ntile * nch * K * w_stage;
        for (int c = 0; c < nch; ++c)
          for (int k0 = 0; k0 < K; k0 += G) {
            const uint32_t bytes = (uint32_t)min(G, K - k0) * w_stage;
            mbar_wait(&w_empty[s], ph);
            mbar_arrive_expect_tx(&w_full[s], bytes);
            bulk_g2s(sW + (size_t)s * w_slot, gw + ((size_t)c * K + k0) * w_stage, bytes, &w_full[s]);
            if (++s == WST) { s = 0; ph ^= 1; }
          }
      }
    }
  } else if (warp == W_ALOAD) {
    // ===================== activation producer: waits for the layer barrier, then bulk-copies the A tiles ==============
    if (lane == 0) {
      int as = 0;
      uint32_t aph = 1;
      for (int l = 0; l < cp.n_layers; ++l) {
        // every CTA has finished (and published) layer l-1 — and with it, in order, all earlier layers.  One counter per
        // layer: a CTA without work (tile beyond the item, n-tile beyond the layer) runs ahead and must not be counted
        // towards an earlier barrier.
        if (l > 0)
          while (ld_acquire_gpu(cp.counter + (l - 1)) < cp.n_ctas) __nanosleep(20);
        asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes of other CTAs -> our async-proxy reads
        const ConvParams& p = cp.layers[l];
        if (!active(p)) continue;
        const int K = p.K, dil = p.dil, nch = (p.Cin + UM_BKC - 1) / UM_BKC;
        const unsigned char* xp = reinterpret_cast<const unsigned char*>(p.xp) + (size_t)b * p.xp_bs;
        const uint32_t kc_bytes = (uint32_t)(128 + (K - 1) * dil) * 16;
        const long long row0 = (long long)t0 - p.pad_left + UM_PADL;
        for (int c = 0; c < nch; ++c, as = (as + 1 == AST ? 0 : as + 1), aph ^= (as == 0)) {
          mbar_wait(&a_empty[as], aph);
          mbar_arrive_expect_tx(&a_full[as], NSPLIT * UM_KCH * kc_bytes);
          unsigned char* dst = sA + (size_t)as * a_stage;
          for (int sp = 0; sp < NSPLIT; ++sp)
#pragma unroll
            for (int kc = 0; kc < UM_KCH; ++kc)
              bulk_g2s(dst + sp * a_split_bytes + kc * RA * 16,
                       xp + (((size_t)sp * p.xp_kc + c * UM_KCH + kc) * p.xp_rows + row0) * 16, kc_bytes, &a_full[as]);
        }
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (same sequence as conv_umma_kernel, NSPLIT = 2) =====================
    int ws = 0, as = 0;
    uint32_t wph = 0, aph = 0, use = 0;
    const uint64_t a_desc0 = make_smem_desc(smem_u32(sA), RA * 16, 128);
    const uint32_t a_ks = (2 * RA * 16) >> 4, a_sp = a_split_bytes >> 4;
    for (int l = 0; l < cp.n_layers; ++l) {
      const ConvParams& p = cp.layers[l];
      if (!active(p)) continue;  // uniform across the warp
      const int NT = p.um_nt, K = p.K, dil = p.dil, nch = (p.Cin + UM_BKC - 1) / UM_BKC;
      const uint32_t w_stage = (uint32_t)NSPLIT * UM_KCH * NT * 16;
      const uint32_t idesc1 = make_idesc_bf16_f32(128, NT), idesc2 = make_idesc_bf16_f32(128, 2 * NT);
      const uint64_t b_desc0 = make_smem_desc(smem_u32(sW), NSPLIT * NT * 16, 128);
      const uint32_t b_ks = (2 * NSPLIT * NT * 16) >> 4;
      mbar_wait(acc_empty, (use & 1) ^ 1);  // the epilogue of this CTA's previous layer has drained the accumulator
      tc_fence_after();
      for (int c = 0; c < nch; ++c) {
        mbar_wait(&a_full[as], aph);
        tc_fence_after();
        const uint64_t a_c = a_desc0 + ((as * a_stage) >> 4);
        for (int k0 = 0; k0 < K; k0 += G) {
          mbar_wait(&w_full[ws], wph);
          tc_fence_after();
          const int cnt = min(G, K - k0);
          if (elect_one()) {
            for (int kk = 0; kk < cnt; ++kk) {
              const int k = k0 + kk;
              const uint64_t a_k = a_c + (uint32_t)(k * dil);
              const uint64_t b_k = b_desc0 + ((ws * w_slot + kk * w_stage) >> 4);
              const uint32_t first = (c == 0 && k == 0) ? 0u : 1u;
#pragma unroll
              for (int ks = 0; ks < UM_BKC / 16; ++ks) {
                const uint64_t a0 = a_k + ks * a_ks, b0 = b_k + ks * b_ks;
                const uint32_t acc0 = ks == 0 ? first : 1u;
                mma_bf16_ss(tmem, a0, b0, idesc2, acc0);            // hi x [hi|lo] -> [c0|c1]
                mma_bf16_ss(tmem + NT, a0 + a_sp, b0, idesc1, 1u);  // lo x hi      -> c1
              }
            }
            mma_commit(&w_empty[ws]);
          }
          __syncwarp();
          if (++ws == WST) { ws = 0; wph ^= 1; }
        }
        if (elect_one()) mma_commit(&a_empty[as]);
        __syncwarp();
        if (++as == AST) { as = 0; aph ^= 1; }
      }
      if (elect_one()) mma_commit(acc_full);
      __syncwarp();
      ++use;
    }
  } else {
    // ===================== epilogue warps =====================
    uint32_t use = 0;
    const int it = 1;  // (tracing in the shared epilogue text is tied to it == 0: off)
    for (int l = 0; l < cp.n_layers; ++l) {
      const ConvParams& p = cp.layers[l];
      const int NT = p.um_nt, n0 = ntile * NT;
      const int Lb = p.lens ? p.lens[b] * p.len_mul : p.Lin;
      if (active(p)) {
        // epilogue constants of this layer -> smem (the previous layer's readers passed the barrier below)
        for (int i = tid; i < NT; i += UM_NPROD) s_bias[i] = (n0 + i) < p.Cout ? p.bias[n0 + i] : 0.f;
        if (p.epi == EPI_COUPLING) {
          if (tid < 16) s_inv[tid] = p.inv_w[tid];
          for (int i = tid; i < 2 * p.half; i += UM_NPROD) {
            s_anb[i] = p.an_bias[i];
            s_ans[i] = p.an_scale[i];
          }
        }
        epi_bar_sync<EW>();
        mbar_wait(acc_full, use & 1);
        tc_fence_after();
        // Everything this epilogue reads from global memory (residual / skip / flow state) was written before the layer
        // barrier this CTA's activation-copy lane has already passed (the accumulator exists only after its copies);
        // an acquire of our own makes that visible to these threads.
        if (lane == 0 && l > 0) (void)ld_acquire_gpu(cp.counter + (l - 1));
        __syncwarp();
        const uint32_t tm = tmem;
#include "conv_umma_epilogue.inc"
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
        ++use;
      }
      // publish this CTA's part of layer l: all epilogue warps' global writes, then one release increment
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
      epi_bar_sync<EW>();
      if (tid == 0) atomicAdd(cp.counter + l, 1u);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace b2
