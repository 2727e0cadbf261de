// flow_chain.cuh — the whole GlowTTS flow decoder (glow_tts/models.py:191-209: 12 x CouplingBlock reverse, each = start ->
// 4 x (WN in_layer + gate -> res_skip) -> end + affine coupling + InvConvNear^-1 + ActNorm^-1) in ONE persistent launch.
//
// Why: at batch 1 the decoder is 120 dependent convs of 6-24 CTAs each; one launch per conv costs ~12 us of which ~7 us
// are launch gap, TMEM allocation, barrier init and pipeline fill/drain (profiles/r1_tcgen05_summary.md), and the
// res/skip epilogue paid two dependent global round trips for the fp32 hidden state and skip sum.
//
// Here a group of (time tiles x 6) CTAs owns one utterance for all layers of all blocks:
//   * CTA (x, y) computes the 128-row time tile x of output-channel tile y of every layer that has such a tile;
//   * layers are separated by a per-utterance barrier in global memory (one counter per layer: release = epilogue
//     warps -> red.release.gpu, acquire = the activation-copy lane before it bulk-loads the next layer's A tiles); only the
//     CTAs of ONE utterance synchronise, groups of different utterances never wait for each other;
//   * the weight-copy lane does not wait for the barrier — weights are static — so the weights of layer l+1 stream into the
//     ring while layer l is still in its epilogue;
//   * the WN hidden state h and the skip sum never exist in fp32 memory: CTAs y = 0..2 keep their 64 channels of h, CTAs
//     y = 3..5 their 64 channels of the skip sum, as a PERSISTENT TMEM accumulator — `start` writes it, every res_skip
//     conv accumulates onto it with the MMA itself (h += rs[:, :192], skip += rs[:, 192:], glow_tts/layers.py:151-160) and
//     the epilogue only reads it to emit the bf16 operand image of the new h (the next in_layer's A operand);
//   * barriers, TMEM (256 columns: working + persistent accumulator) and the operand rings are set up once per launch.
// The conv itself is the plane-mode path of conv_umma_kernel<1, 2, ...>: same operand images, same MMA sequence per
// k-chunk, the very same epilogue text (conv_umma_epilogue.inc) for the gate and the coupling.
#pragma once
#include "conv_umma.cuh"
#include "flow_chain_params.cuh"

namespace b2 {

constexpr int FC_NT = 64;      // N tile of start / in / res_skip (pack.py: B200TTS_GLOW_NT)
constexpr int FC_YT = 6;       // channel-tile CTAs per time tile = 2 * Hd / FC_NT
constexpr int FC_EW = 8;       // epilogue warps
constexpr int FC_KMAX = 5;     // taps of an in_layer (bounded so that the geometry check is explicit)
#ifndef FC_G1_V
#define FC_G1_V 3
#endif
#ifndef FC_WST_V
#define FC_WST_V 3
#endif
#ifndef FC_AST_V
#define FC_AST_V 4
#endif
constexpr int FC_G1 = FC_G1_V;  // 1x1 convs: 32-channel chunks per weight slot
constexpr int FC_WST = FC_WST_V;  // weight ring slots of FC_KMAX blocks (40 KB): one chunk with all its taps (in_layer) or FC_G1
                               // chunks of a 1x1 conv.  Per slot the MMA warp pays one wait and one commit; with one block
                               // per slot that fixed cost (~150 clk) exceeded the 4 MMAs it paid for (measured: 6.9 vs 4.8 us)
constexpr int FC_AST = FC_AST_V;
constexpr int FC_HALO = 16;    // max (K-1)*dil
constexpr int FC_RA = 128 + FC_HALO;
constexpr int FC_THREADS = FC_EW * 32 + 96;
constexpr uint32_t FC_A_SPLIT = UM_KCH * FC_RA * 16;
constexpr uint32_t FC_A_STAGE = 2 * FC_A_SPLIT;
constexpr uint32_t FC_W_SLOT = (uint32_t)FC_KMAX * 2 * UM_KCH * FC_NT * 16;
constexpr size_t FC_SMEM = (size_t)FC_AST * FC_A_STAGE + (size_t)FC_WST * FC_W_SLOT + 512 + 4 * (64 + 64 + 16 + 2 * 256);

// The fields conv_umma_epilogue.inc reads from `p`, for this kernel: register-resident, and everything the flow never uses
// is a compile-time constant so that the shared epilogue text collapses to the one mode it is instantiated for.
template <int EPI>
struct FlowEpiP {
  static constexpr int epi = EPI, pix = 1, act = ACT_NONE, yp_nsplit = 2, split = 0, first = 0;
  static constexpr float yp_slope = 1.f, rp_inv_slope = 1.f;
  static constexpr const void* rp = nullptr;
  static constexpr long long rp_bs = 0;
  static constexpr int rp_kc = 0, rp_rows = 0;
  static constexpr const float* res = nullptr;
  static constexpr long long res_bs = 0;
  static constexpr int res_ld = 0;
  static constexpr float* aux = nullptr;
  static constexpr long long aux_bs = 0;
  static constexpr int aux_ld = 0;
  static constexpr void* yp2 = nullptr;
  static constexpr unsigned long long* dbg = nullptr;
  void* yp;
  long long yp_bs;
  int yp_kc, yp_rows;
  float* y;
  long long y_bs;
  int y_ld;
  float* tap;
  int half, Cout;
};

enum FlowStepKind { FS_START = 0, FS_IN = 1, FS_RS = 2, FS_RS_LAST = 3, FS_END = 4 };

// what CTA row y does in one step of a coupling block
struct FlowStep {
  int kind, l;
  bool active;     // this CTA has a tile in this step
  int wt;          // weight n-tile it multiplies with
  int NT, nch, K, dil, pad;
  int acc;         // 0: working accumulator (TMEM columns [0, 128)), 1: persistent accumulator ([128, 256))
  int init;        // 1: the first MMA overwrites the accumulator, 0: accumulates onto it
  int epi;         // -1: no epilogue (the persistent skip sum just grows), else EpiMode
  int n_active;    // CTA rows taking part (for the barrier count)
};

__device__ __forceinline__ FlowStep flow_step(const FlowChainParams& fp, int step, int y) {
  FlowStep s;
  const int last = 1 + 2 * fp.n_wn;
  s.l = step > 0 ? (step - 1) >> 1 : 0;
  s.NT = FC_NT; s.K = 1; s.dil = 1; s.pad = 0; s.acc = 0; s.init = 1; s.epi = EPI_PLAIN; s.wt = y;
  s.nch = (fp.Hd + UM_BKC - 1) / UM_BKC;
  const int hd_tiles = fp.Hd / FC_NT;
  if (step == 0) {
    s.kind = FS_START;
    s.nch = (fp.half + UM_BKC - 1) / UM_BKC;
    s.n_active = hd_tiles;
    s.active = y < hd_tiles;
    s.acc = 1;
  } else if (step == last) {
    s.kind = FS_END;
    s.NT = fp.nt_end;
    s.n_active = (2 * fp.half + fp.nt_end - 1) / fp.nt_end;
    s.active = y < s.n_active;
    s.epi = EPI_COUPLING;
  } else if ((step - 1) & 1) {
    const bool is_last = s.l == fp.n_wn - 1;
    s.kind = is_last ? FS_RS_LAST : FS_RS;
    s.acc = 1;
    if (!is_last) {
      s.n_active = 2 * hd_tiles;
      s.active = true;
      const bool skip = y >= hd_tiles;
      s.init = skip && s.l == 0;
      s.epi = skip ? -1 : EPI_PLAIN;
    } else {  // Cout = Hd, all of it skip: the skip owners (y >= hd_tiles) take weight tile y - hd_tiles
      s.n_active = hd_tiles;
      s.active = y >= hd_tiles;
      s.wt = y - hd_tiles;
      s.init = s.l == 0;
    }
  } else {
    s.kind = FS_IN;
    s.K = fp.K;
    int d = 1;
    for (int j = 0; j < s.l; ++j) d *= fp.dil_rate;
    s.dil = d;
    s.pad = (fp.K * d - d) / 2;
    s.n_active = 2 * hd_tiles;
    s.active = true;
    s.epi = EPI_GATE;
  }
  return s;
}

__device__ __forceinline__ unsigned int fc_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fc_red_release(unsigned int* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void fc_epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(FC_EW * 32) : "memory"); }

#define FC_TRACE(L, ev)                                                                              \
  do {                                                                                               \
    if (fp.trace && blockIdx.x == 0 && blockIdx.z == 0 && b == 0) fp.trace[((size_t)(L)*FC_YT + y) * 8 + (ev)] = gtimer(); \
  } while (0)

__global__ void __launch_bounds__(FC_THREADS, 1) flow_chain_kernel(const __grid_constant__ FlowChainParams fp) {
  using namespace umma;
  constexpr int MT = 1, NSPLIT = 2, NACC = 2, EW = FC_EW;
  constexpr int RA = FC_RA;
  constexpr int W_MMA = EW, W_LOAD = EW + 1, W_ALOAD = EW + 2;
  constexpr uint32_t a_split_bytes = FC_A_SPLIT, a_stage = FC_A_STAGE, w_slot = FC_W_SLOT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sA = smem_raw;
  unsigned char* sW = smem_raw + FC_AST * a_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + FC_WST * w_slot);
  uint64_t* a_full = bars;          // [AST]
  uint64_t* a_empty = bars + 4;     // [AST]
  uint64_t* w_full = bars + 8;      // [WST]
  uint64_t* w_empty = bars + 24;    // [WST]
  uint64_t* acc_full = bars + 40;   // MMA -> epilogue
  uint64_t* acc_empty = bars + 41;  // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 42);
  static_assert(FC_WST <= 16 && FC_AST <= 4, "barrier slots");
  float* s_wbias = reinterpret_cast<float*>(bars + 64);
  float* s_pbias = s_wbias + 64;    // running bias of the persistent accumulator (sum of the biases accumulated so far)
  float* s_inv = s_pbias + 64;
  float* s_anb = s_inv + 16;
  float* s_ans = s_anb + 256;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mtile = blockIdx.x, y = blockIdx.y;
  const int t0 = mtile * 128;
  constexpr uint32_t tmem_cols = 256;
  const int n_steps = 2 + 2 * fp.n_wn;

  if (tid == 0) {
    for (int i = 0; i < FC_AST; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < FC_WST; ++i) {
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

  if (warp == W_LOAD) {
    // ===================== weight producer: runs ahead of the layer barriers (weights are static) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int b = blockIdx.z; b < fp.batch; b += gridDim.z) {
        if (t0 >= fp.lens[b]) continue;
        for (int blk = 0; blk < fp.n_blocks; ++blk) {
          const FlowBlockDev& fb = fp.blocks[blk];
          for (int step = 0; step < n_steps; ++step) {
            const FlowStep st = flow_step(fp, step, y);
            if (!st.active) continue;
            const uint32_t w_stage = (uint32_t)NSPLIT * UM_KCH * st.NT * 16;
            const unsigned char* gw = reinterpret_cast<const unsigned char*>(fb.w[step]) + (size_t)st.wt * st.nch * st.K * w_stage;
            const int n_blocks = st.nch * st.K;  // (chunk, tap) blocks are consecutive in the image
            const int G = st.K > 1 ? st.K : FC_G1;
            for (int i = 0; i < n_blocks; i += G) {
              const uint32_t bytes = (uint32_t)min(G, n_blocks - i) * w_stage;
              mbar_wait(&w_empty[s], ph);
              mbar_arrive_expect_tx(&w_full[s], bytes);
              bulk_g2s(sW + (size_t)s * w_slot, gw + (size_t)i * w_stage, bytes, &w_full[s]);
              if (++s == FC_WST) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == W_ALOAD) {
    // ===================== activation producer: waits for the previous layer of THIS utterance, then bulk-copies ========
    // The whole warp walks the layers; a 32-channel chunk is 2 planes x 4 channel groups = 8 copies of ~2 KB, and one lane
    // needs ~100 ns per copy (48 copies per layer = most of the 4.6-5.3 us "mma" phase in profiles/r2a_chain_trace.log), so
    // the copies of up to four chunks (= the four ring stages) are issued by 32 lanes at once.
    {
      int as = 0;
      uint32_t aph = 1;
      const int q = lane >> 3, j = lane & 7;
      for (int b = blockIdx.z; b < fp.batch; b += gridDim.z) {
        const int Lb = fp.lens[b];
        if (t0 >= Lb) continue;
        const unsigned int m_tiles = (unsigned int)((Lb + 127) / 128);
        const unsigned int* cnt = fp.counter + (size_t)b * fp.n_blocks * n_steps;
        unsigned int prev_expect = 0;
        for (int blk = 0; blk < fp.n_blocks; ++blk) {
          for (int step = 0; step < n_steps; ++step) {
            const FlowStep st = flow_step(fp, step, y);
            const int L = blk * n_steps + step;
            if (st.active) {
              if (L > 0 && lane == 0) {
                // every CTA of this utterance has finished (and published) layer L-1 — and with it, in order, all earlier ones
                unsigned int spins = 0;
                while (fc_ld_acquire(cnt + (L - 1)) < prev_expect) {
                  if (++spins > (1u << 25)) {  // seconds: a lost arrival must fail loudly, not hang the device
                    printf("flow_chain: barrier %d of item %d timed out (CTA %d,%d,%d)\n", L - 1, b, blockIdx.x, blockIdx.y, blockIdx.z);
                    __trap();
                  }
                }
              }
              __syncwarp();
              asm volatile("fence.proxy.async;" ::: "memory");  // (each issuing lane) other CTAs' generic-proxy writes -> our async-proxy reads
              if (lane == 0) FC_TRACE(L, 0);
              const unsigned char* img;
              long long bs;
              int kc_total;
              if (st.kind == FS_START) { img = fp.S_p; bs = fp.Sp_bs; kc_total = fp.Sp_kc; }
              else if (st.kind == FS_IN) { img = fp.H_p; bs = fp.Hp_bs; kc_total = fp.Hp_kc; }
              else if (st.kind == FS_END) { img = fp.K_p; bs = fp.Hp_bs; kc_total = fp.Hp_kc; }
              else { img = fp.A_p; bs = fp.Hp_bs; kc_total = fp.Hp_kc; }
              const unsigned char* xp = img + (size_t)b * bs;
              // copied from the 8-row (128-byte) boundary below the first row: unaligned bulk copies run ~3x slower
              const long long row0u = (long long)t0 - st.pad + UM_PADL;
              const uint32_t roff = a_row_off(row0u);
              const long long row0 = row0u - roff;
              const uint32_t kc_bytes = (uint32_t)((128 + (st.K - 1) * st.dil + (int)roff + 7) & ~7) * 16;
              static_assert(FC_AST >= 1 && FC_AST <= 4 && NSPLIT * UM_KCH == 8, "lane -> (stage, copy) mapping");
              for (int c0 = 0; c0 < st.nch; c0 += FC_AST) {
                const int c = c0 + q;
                int sq = as + q;
                uint32_t pq = aph;
                if (sq >= FC_AST) { sq -= FC_AST; pq ^= 1; }
                if (q < FC_AST && c < st.nch) {
                  mbar_wait(&a_empty[sq], pq);
                  if (j == 0) mbar_arrive_expect_tx(&a_full[sq], NSPLIT * UM_KCH * kc_bytes);
                  const int sp = j >> 2, kc = j & 3;
                  bulk_g2s(sA + (size_t)sq * a_stage + sp * a_split_bytes + kc * RA * 16,
                           xp + (((size_t)sp * kc_total + c * UM_KCH + kc) * fp.rows + row0) * 16, kc_bytes, &a_full[sq]);
                }
                as += min(FC_AST, st.nch - c0);
                if (as >= FC_AST) { as -= FC_AST; aph ^= 1; }
                __syncwarp();
              }
            }
            prev_expect = (unsigned int)st.n_active * m_tiles;
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (same sequence per k-chunk as conv_umma_kernel, NSPLIT = 2) =====================
    int ws = 0, as = 0;
    uint32_t wph = 0, aph = 0, use = 0;
    const uint64_t a_desc0 = make_smem_desc(smem_u32(sA), RA * 16, 128);
    const uint32_t a_ks = (2 * RA * 16) >> 4, a_sp = a_split_bytes >> 4;
    for (int b = blockIdx.z; b < fp.batch; b += gridDim.z) {
      if (t0 >= fp.lens[b]) continue;
      for (int blk = 0; blk < fp.n_blocks; ++blk)
        for (int step = 0; step < n_steps; ++step) {
          const FlowStep st = flow_step(fp, step, y);
          if (!st.active) continue;  // uniform across the warp
          const int NT = st.NT, K = st.K, dil = st.dil, nch = st.nch;
          const uint32_t idesc1 = make_idesc_bf16_f32(128, NT), idesc2 = make_idesc_bf16_f32(128, 2 * NT);
          const uint64_t b_desc0 = make_smem_desc(smem_u32(sW), NSPLIT * NT * 16, 128);
          const uint32_t b_ks = (2 * NSPLIT * NT * 16) >> 4;
          const uint32_t tm = tmem + (st.acc ? 128u : 0u);
          const uint32_t roff = a_row_off((long long)t0 - st.pad + UM_PADL);
          mbar_wait(acc_empty, (use & 1) ^ 1);  // the epilogue of this CTA's previous step is done with the accumulators
          tc_fence_after();
          const uint32_t w_stage = (uint32_t)NSPLIT * UM_KCH * NT * 16;
          auto issue = [&](uint64_t a_k, uint64_t b_k, uint32_t first) {  // both K=16 steps of one (chunk, tap) block
#pragma unroll
            for (int ks = 0; ks < UM_BKC / 16; ++ks) {
              const uint64_t a0 = a_k + ks * a_ks, b0 = b_k + ks * b_ks;
              const uint32_t acc0 = ks == 0 ? first : 1u;
              mma_bf16_ss(tm, a0, b0, idesc2, acc0);            // hi x [hi|lo] -> [c0|c1]
              mma_bf16_ss(tm + NT, a0 + a_sp, b0, idesc1, 1u);  // lo x hi      -> c1
            }
          };
          if (K > 1) {  // in_layer: a weight slot = one chunk with all its taps
            for (int c = 0; c < nch; ++c) {
              const bool mt_ = blk == 0 && step == 1 && lane == 0;  // debug micro-trace of the first in_layer
              if (mt_) FC_TRACE(210 + c, 0);
              mbar_wait(&a_full[as], aph);
              tc_fence_after();
              if (mt_) FC_TRACE(210 + c, 1);
              if (c == 0 && lane == 0) FC_TRACE(blk * n_steps + step, 1);
              mbar_wait(&w_full[ws], wph);
              tc_fence_after();
              if (mt_) FC_TRACE(210 + c, 2);
              if (c == 0 && lane == 0) FC_TRACE(blk * n_steps + step, 2);
              if (elect_one()) {
                const uint64_t a_c = a_desc0 + ((as * a_stage) >> 4) + roff;
                for (int k = 0; k < K; ++k)
                  issue(a_c + (uint32_t)(k * dil), b_desc0 + ((ws * w_slot + k * w_stage) >> 4), (c == 0 && k == 0 && st.init) ? 0u : 1u);
                mma_commit(&w_empty[ws]);
                mma_commit(&a_empty[as]);
              }
              __syncwarp();
              if (blk == 0 && step == 1 && lane == 0) FC_TRACE(210 + c, 3);
              if (++ws == FC_WST) { ws = 0; wph ^= 1; }
              if (++as == FC_AST) { as = 0; aph ^= 1; }
            }
          } else {  // 1x1 conv: a weight slot = FC_G1 chunks
            for (int c0 = 0; c0 < nch; c0 += FC_G1) {
              mbar_wait(&w_full[ws], wph);
              tc_fence_after();
              if (c0 == 0 && lane == 0) FC_TRACE(blk * n_steps + step, 2);
              const int cnt = min(FC_G1, nch - c0);
              for (int j = 0; j < cnt; ++j) {
                const bool mt_ = blk == 0 && step == 2 && lane == 0;  // debug micro-trace of the first res_skip layer
                if (mt_) FC_TRACE(200 + c0 + j, 0);
                mbar_wait(&a_full[as], aph);
                tc_fence_after();
                if (mt_) FC_TRACE(200 + c0 + j, 1);
                if (c0 + j == 0 && lane == 0) FC_TRACE(blk * n_steps + step, 1);
                if (elect_one()) {
                  issue(a_desc0 + ((as * a_stage) >> 4) + roff, b_desc0 + ((ws * w_slot + j * w_stage) >> 4), (c0 + j == 0 && st.init) ? 0u : 1u);
                  mma_commit(&a_empty[as]);
                  if (j == cnt - 1) mma_commit(&w_empty[ws]);
                }
                __syncwarp();
                if (mt_) FC_TRACE(200 + c0 + j, 2);
                if (++as == FC_AST) { as = 0; aph ^= 1; }
              }
              if (++ws == FC_WST) { ws = 0; wph ^= 1; }
            }
          }
          if (elect_one()) mma_commit(acc_full);
          __syncwarp();
          if (lane == 0) FC_TRACE(blk * n_steps + step, 3);
          ++use;
        }
    }
  } else {
    // ===================== epilogue warps =====================
    uint32_t use = 0;
    const int it = 1;  // (tracing in the shared epilogue text is tied to it == 0: off)
    for (int b = blockIdx.z; b < fp.batch; b += gridDim.z) {
      const int Lb = fp.lens[b];
      if (t0 >= Lb) continue;
      unsigned int* cnt = fp.counter + (size_t)b * fp.n_blocks * n_steps;
      for (int blk = 0; blk < fp.n_blocks; ++blk) {
        const FlowBlockDev& fb = fp.blocks[blk];
        for (int step = 0; step < n_steps; ++step) {
          const FlowStep st = flow_step(fp, step, y);
          if (!st.active) continue;
          const int NT = st.NT;
          const int n0 = st.wt * NT;  // first packed output channel of this CTA's tile
          // ---- this step's epilogue constants -> smem (the previous step's readers are past fc_epi_bar below)
          const float* bias = fb.b[step];
          if (st.acc == 0) {
            for (int i = tid; i < NT; i += EW * 32) s_wbias[i] = bias[n0 + i];
          } else if (tid < NT) {  // persistent accumulator: its bias is the sum of the biases of everything accumulated
            s_pbias[tid] = st.init ? bias[n0 + tid] : s_pbias[tid] + bias[n0 + tid];
          }
          if (st.kind == FS_END) {
            if (tid < 16) s_inv[tid] = fb.inv[tid];
            for (int i = tid; i < 2 * fp.half; i += EW * 32) {
              s_anb[i] = fb.an_bias[i];
              s_ans[i] = fb.an_scale[i];
            }
          }
          fc_epi_bar();
          mbar_wait(acc_full, use & 1);
          tc_fence_after();
          if (tid == 0) FC_TRACE(blk * n_steps + step, 4);
          const uint32_t tm = tmem + (st.acc ? 128u : 0u);
          const float* s_bias = st.acc ? s_pbias : s_wbias;
          if (st.epi == EPI_GATE) {  // in_layer: tanh * sigmoid -> operand image of the gate output
            FlowEpiP<EPI_GATE> p;
            p.yp = fp.A_p; p.yp_bs = fp.Hp_bs; p.yp_kc = fp.Hp_kc; p.yp_rows = fp.rows;
            p.y = nullptr; p.y_bs = 0; p.y_ld = 0; p.tap = nullptr; p.half = fp.half; p.Cout = 2 * fp.Hd;
#include "conv_umma_epilogue.inc"
          } else if (st.epi == EPI_COUPLING) {  // end: affine coupling + InvConvNear^-1 + ActNorm^-1 on the fp32 state
            FlowEpiP<EPI_COUPLING> p;
            p.yp = fp.S_p; p.yp_bs = fp.Sp_bs; p.yp_kc = fp.Sp_kc; p.yp_rows = fp.rows;
            p.y = fp.S; p.y_bs = fp.S_bs; p.y_ld = fp.S_ld;
            p.tap = fp.taps ? fp.taps + (size_t)fb.tap_index * fp.batch * fp.S_bs : nullptr;
            p.half = fp.half; p.Cout = 2 * fp.half;
#include "conv_umma_epilogue.inc"
          } else if (st.epi == EPI_PLAIN) {  // persistent accumulator + running bias -> operand image of h / of the skip sum
            FlowEpiP<EPI_PLAIN> p;
            p.yp = st.kind == FS_RS_LAST ? fp.K_p : fp.H_p; p.yp_bs = fp.Hp_bs; p.yp_kc = fp.Hp_kc; p.yp_rows = fp.rows;
            p.y = nullptr; p.y_bs = 0; p.y_ld = 0; p.tap = nullptr; p.half = fp.half; p.Cout = fp.Hd;
#include "conv_umma_epilogue.inc"
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
          ++use;
          if (tid == 0) FC_TRACE(blk * n_steps + step, 5);
          // publish this CTA's part of the layer: the named barrier orders every epilogue thread's global writes before
          // thread 0, whose gpu-scope release is cumulative over them (the pattern of a cooperative-groups grid barrier)
          fc_epi_bar();
          if (tid == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");
            fc_red_release(cnt + (blk * n_steps + step));
          }
          if (tid == 0) FC_TRACE(blk * n_steps + step, 6);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace b2
