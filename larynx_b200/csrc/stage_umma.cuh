// stage_umma.cuh — ALL ResBlock convs of one HiFi-GAN stage (3 MRF branches x 3 dilations x (c1, c2) = 18 convs,
// hifi_gan/models.py:91-98, 191-197) in ONE persistent launch with tile-level dataflow between the convs.
//
// Why: with one launch per conv position (6 per stage, 3 branches each) every launch walks only 5-7 tiles per CTA and pays
// its pipeline fill, the drain of its last epilogue and the imbalance between the k = 3 / 7 / 11 branches before the next
// launch may start (griddepcontrol.wait = the WHOLE previous grid): ~8 us of ~35-50 us per launch in the narrow stages
// (profiles/r2c_voc_trace.log, r2a_launches.csv).  A conv tile only needs the three neighbouring tiles of its producer.
//
// Here every CTA owns a contiguous range of time tiles of one branch (sometimes the end of one branch and the start of the
// next: the host cuts the ranges so that every CTA gets the same modelled cost) for ALL six conv levels and walks them
// level by level.  Before the activation warp bulk-loads tile i of conv q it waits (ld.acquire.gpu) until tiles i-1, i, i+1
// of the producing conv have been published; the epilogue publishes a tile (red.release.gpu) after its stores.  Inside a
// CTA's own range the producer tiles were finished one level earlier, so only the range borders ever wait, the operand rings
// and the two TMEM accumulator buffers run straight through level boundaries, and the 148 CTAs finish together.
// The same waits also order the buffer reuse (T and R images ping-pong through the dilations): a tile is only rewritten
// after every reader of its old contents has been published, because readers and writers are the same three neighbours.
// All CTAs must be resident (they spin on each other): cooperative launch, grid <= #SMs.
//
// The per-tile conv is conv_umma_kernel's persistent path (NBUF = 2): same operand images, same MMA sequence, the same
// epilogue text.
#pragma once
#include "conv_umma.cuh"
#include "rbpair_umma.cuh"

namespace b2 {

struct StageSeg {
  int col;      // branch = column of the conv table
  int b;        // batch item
  int ntile;    // output-channel tile
  int t_begin, t_end;  // time tiles [t_begin, t_end)
};

struct StageParams {
  const ConvParams* convs;  // device [n_levels][n_cols]: conv (level, col); pair mode: the second conv (c2) of the ResBlock step
  const ConvParams* convs1; // pair mode (stage_pair_kernel): the first conv (c1) of each step; nullptr otherwise
  const int* dep;           // [n_levels * n_cols]: flat index of the conv that writes this conv's input image, -1 = before the launch
  const StageSeg* segs;     // segments of all CTAs
  const int* seg_off;       // [gridDim.x + 1]
  unsigned int* flags;      // [n_levels * n_cols][zb][tiles_max] published n-tiles per (conv, item, time tile); zeroed before the launch
  int n_levels, n_cols, zb, tiles_max, ntiles;
  unsigned long long* trace;  // debug: 8 stamps per item of CTA trace_cta (item order = walk order), or nullptr
  int trace_cta, trace_items;
};

constexpr int SG_EW = 16;

#ifndef SG_WST1
#define SG_WST1 6  // weight ring slots (16 KB) of the 128-channel-tile variant
#endif
#ifndef SG_AST1
#define SG_AST1 4  // activation ring stages (one 32-channel chunk, 26.6 KB) of the 128-channel-tile variant: a whole item of the
                   // 128-channel stage; 4 + 6 slots measured 1 % faster than 2 + 8 (profiles/r2d_notes.md)
#endif

template <int MT>
struct StageCfg {  // ring geometry per tile height (the persistent variants of conv_umma.cu)
  static constexpr int WST = MT == 1 ? SG_WST1 : (MT == 2 ? 6 : 3);
  static constexpr int AST = MT == 1 ? SG_AST1 : 2;
  static constexpr int G = MT == 1 ? 1 : (MT == 2 ? 2 : 4);
  static constexpr int RA = MT * 128 + UM_HALO;
  static size_t smem_bytes(int NT, int Cout) {
    return (size_t)AST * 2 * UM_KCH * RA * 16 + (size_t)WST * G * 2 * UM_KCH * NT * 16 + 512 + 4 * (size_t)SG_MAX_CONVS * Cout;
  }
};

__device__ __forceinline__ unsigned int sg_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void sg_red_release(unsigned int* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

// epilogue warps -> publisher lane: a monotonic counter in shared memory (an mbarrier phase can be missed by a waiter that
// falls two phases behind; a counter cannot)
__device__ __forceinline__ void sg_done_add(unsigned int* p) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(umma::smem_u32(p)) : "memory");
}
__device__ __forceinline__ unsigned int sg_done_ld(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(umma::smem_u32(p)) : "memory");
  return v;
}

#define SG_TRACE(item, ev)                                                                                   \
  do {                                                                                                       \
    if (sp.trace && (int)blockIdx.x == sp.trace_cta && (item) < sp.trace_items) sp.trace[(size_t)(item)*8 + (ev)] = gtimer(); \
  } while (0)

template <int MT>
__global__ void __launch_bounds__(SG_EW * 32 + 128, 1) stage_umma_kernel(const __grid_constant__ StageParams sp) {
  using namespace umma;
  using Cfg = StageCfg<MT>;
  constexpr int NSPLIT = 2, NACC = 2, EW = SG_EW, WST = Cfg::WST, AST = Cfg::AST, G = Cfg::G, RA = Cfg::RA;
  constexpr int NTHREADS = EW * 32 + 128;
  constexpr int W_MMA = EW, W_LOAD = EW + 1, W_ALOAD = EW + 2, W_PUB = EW + 3;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const ConvParams& p0 = sp.convs[0];  // geometry shared by all convs of the stage
  const int NT = p0.um_nt, C = p0.Cin, Cout = p0.Cout;
  const int nch = (C + UM_BKC - 1) / UM_BKC;
  const uint32_t a_split_bytes = UM_KCH * RA * 16, a_stage = NSPLIT * a_split_bytes;
  const uint32_t w_stage = NSPLIT * UM_KCH * NT * 16, w_ring = G * w_stage;
  unsigned char* sA = smem_raw;
  unsigned char* sW = smem_raw + AST * a_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + WST * w_ring);
  uint64_t* a_full = bars;          // [AST <= 4]
  uint64_t* a_empty = bars + 4;     // [AST <= 4]
  uint64_t* w_full = bars + 8;      // [WST <= 9]
  uint64_t* w_empty = bars + 17;    // [WST <= 9]
  uint64_t* acc_full = bars + 26;   // [2]
  uint64_t* acc_empty = bars + 28;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);
  unsigned int* s_done = reinterpret_cast<unsigned int*>(bars + 31);  // epilogue warps that have finished an item (monotonic)
  static_assert(WST <= 9 && AST <= 4, "barrier slots");
  float* s_bias_all = reinterpret_cast<float*>(bars + 32);  // [n_convs][Cout]
  const float *s_inv = nullptr, *s_anb = nullptr, *s_ans = nullptr;  // named by the (dead here) coupling branch of the epilogue text

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_convs = sp.n_levels * sp.n_cols;
  const int seg0 = sp.seg_off[blockIdx.x], seg1 = sp.seg_off[blockIdx.x + 1];
  if (seg0 == seg1) return;  // (no work: never waited for by anybody)
  constexpr uint32_t ACC_COLS = NACC * MT;  // x NT

  if (tid == 0) {
    for (int i = 0; i < AST; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < WST; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], EW);
    }
    *s_done = 0;
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
  for (int cv = warp; cv < n_convs; cv += NTHREADS / 32) {  // one warp per conv: its bias pointer is loaded once, the values coalesced
    const float* bp = sp.convs[cv].bias;
    for (int i = lane; i < Cout; i += 32) s_bias_all[cv * Cout + i] = bp[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // every role walks the same item sequence: level by level over this CTA's segments
  auto walk = [&](auto&& body) {
    for (int lv = 0; lv < sp.n_levels; ++lv)
      for (int sg = seg0; sg < seg1; ++sg) {
        const StageSeg seg = sp.segs[sg];
        const int conv = lv * sp.n_cols + seg.col;
        const ConvParams& p = sp.convs[conv];
        for (int mtile = seg.t_begin; mtile < seg.t_end; ++mtile) body(conv, p, seg.b, seg.ntile, mtile);
      }
  };

  if (warp == W_LOAD) {
    // ===================== weight producer (weights are static: runs ahead of every dependency) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      walk([&](int conv, const ConvParams& p, int b, int ntile, int mtile) {
        (void)conv; (void)b; (void)mtile;
        const int K = p.K;
        const unsigned char* gw = reinterpret_cast<const unsigned char*>(p.wu) + (size_t)ntile * nch * K * w_stage;
        for (int c = 0; c < nch; ++c)
          for (int k0 = 0; k0 < K; k0 += G) {
            const uint32_t bytes = (uint32_t)min(G, K - k0) * w_stage;
            mbar_wait(&w_empty[s], ph);
            mbar_arrive_expect_tx(&w_full[s], bytes);
            bulk_g2s(sW + (size_t)s * w_ring, gw + ((size_t)c * K + k0) * w_stage, bytes, &w_full[s]);
            if (++s == WST) { s = 0; ph ^= 1; }
          }
      });
    }
  } else if (warp == W_ALOAD) {
    // ===================== activation producer: waits for the producer tiles, then bulk-copies =====================
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int as = 0, ita = 0;
    uint32_t aph = 1;
    walk([&](int conv, const ConvParams& p, int b, int ntile, int mtile) {
      (void)ntile;
      if (lane == 0) SG_TRACE(ita, 0);
      const int d = sp.dep[conv];
      if (d >= 0) {
        if (lane == 0) {
          const int Lb = p.lens ? p.lens[b] * p.len_mul : p.Lin;
          const int m_tiles = (Lb + MT * 128 - 1) / (MT * 128);
          const unsigned int* fl = sp.flags + ((size_t)d * sp.zb + b) * sp.tiles_max;
          const int i0 = max(mtile - 1, 0), i1 = min(mtile + 1, m_tiles - 1);
          for (int ii = i0; ii <= i1; ++ii) {
            unsigned int spins = 0;
            while (sg_ld_acquire(fl + ii) < (unsigned int)sp.ntiles) {
              if (++spins > (1u << 26)) {  // seconds: a lost publication must fail loudly, not hang the device
                printf("stage_umma: tile %d of conv %d (item %d) was never published (CTA %d waits for conv %d tile %d)\n", ii, d, b,
                       blockIdx.x, conv, mtile);
                __trap();
              }
            }
          }
        }
        __syncwarp();
        asm volatile("fence.proxy.async;" ::: "memory");  // (each issuing lane) other CTAs' generic-proxy writes -> our async-proxy reads
      }
      if (lane == 0) SG_TRACE(ita, 1);
      const unsigned char* xp = reinterpret_cast<const unsigned char*>(p.xp) + (size_t)b * p.xp_bs;
      const long long row0 = (long long)mtile * (MT * 128) - p.pad_left + UM_PADL;
      issue_a_chunks<NSPLIT, AST>(lane, as, aph, nch, a_empty, a_full, sA, a_stage, a_split_bytes, RA, xp, p.xp_kc, p.xp_rows, row0,
                                  MT * 128 + (p.K - 1) * p.dil);
      ++ita;
    });
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (conv_umma_kernel's sequence, NSPLIT = 2) =====================
    const uint32_t idesc1 = make_idesc_bf16_f32(128, NT), idesc2 = make_idesc_bf16_f32(128, 2 * NT);
    const uint64_t a_desc0 = make_smem_desc(smem_u32(sA), RA * 16, 128);
    const uint64_t b_desc0 = make_smem_desc(smem_u32(sW), NSPLIT * NT * 16, 128);
    const uint32_t a_ks = (2 * RA * 16) >> 4, b_ks = (2 * NSPLIT * NT * 16) >> 4;
    const uint32_t a_sp = a_split_bytes >> 4;
    const uint32_t a_mt = (128 * 16) >> 4;
    int ws = 0, as = 0;
    uint32_t wph = 0, aph = 0;
    int it = 0;
    walk([&](int conv, const ConvParams& p, int b, int ntile, int mtile) {
      (void)conv; (void)b; (void)ntile;
      const int K = p.K, dil = p.dil;
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t tm = tmem + (uint32_t)buf * (ACC_COLS * NT);
      const uint32_t roff = a_row_off((long long)mtile * (MT * 128) - p.pad_left + UM_PADL);
      if (lane == 0) SG_TRACE(it, 3);
      for (int c = 0; c < nch; ++c) {
        mbar_wait(&a_full[as], aph);
        tc_fence_after();
        const uint64_t a_c = a_desc0 + ((as * a_stage) >> 4) + roff;
        if (c == 0 && lane == 0) SG_TRACE(it, 7);
        for (int k0 = 0; k0 < K; k0 += G) {
          mbar_wait(&w_full[ws], wph);
          tc_fence_after();
          if (c == 0 && k0 == 0 && lane == 0) SG_TRACE(it, 2);
          const int cnt = min(G, K - k0);
          if (elect_one()) {
            for (int kk = 0; kk < cnt; ++kk) {
              const int k = k0 + kk;
              const uint64_t a_k = a_c + (uint32_t)(k * dil);
              const uint64_t b_k = b_desc0 + ((ws * w_ring + kk * w_stage) >> 4);
              const uint32_t first = (c == 0 && k == 0) ? 0u : 1u;
#pragma unroll
              for (int ks = 0; ks < UM_BKC / 16; ++ks) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  const uint64_t a0 = a_k + ks * a_ks + mt * a_mt, b0 = b_k + ks * b_ks;
                  const uint32_t acc0 = ks == 0 ? first : 1u;
                  const uint32_t d0 = tm + (uint32_t)(mt * NACC) * NT;
                  mma_bf16_ss(d0, a0, b0, idesc2, acc0);            // hi x [hi|lo] -> [c0|c1]
                  mma_bf16_ss(d0 + NT, a0 + a_sp, b0, idesc1, 1u);  // lo x hi      -> c1
                }
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
      if (elect_one()) mma_commit(&acc_full[buf]);
      __syncwarp();
      if (lane == 0) SG_TRACE(it, 4);
      ++it;
    });
  } else if (warp < EW) {
    // ===================== epilogue warps =====================
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int it = 0;
    walk([&](int conv, const ConvParams& p, int b, int ntile, int mtile) {
      const int n0 = ntile * NT;
      const int Lb = p.lens ? p.lens[b] * p.len_mul : p.Lin;
      const int t0 = mtile * (MT * 128);
      const float* s_bias = s_bias_all + conv * Cout + n0;
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(&acc_full[buf], use & 1);
      tc_fence_after();
      if (tid == 0) SG_TRACE(it, 5);
      {
        const uint32_t tm = tmem + (uint32_t)buf * (ACC_COLS * NT);
        // images rewritten by other CTAs during this launch must not be read through a stale L1 line
#define UM_EPI_LD16(ptr) __ldcg(reinterpret_cast<const uint4*>(ptr))
#include "conv_umma_epilogue.inc"
#undef UM_EPI_LD16
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&acc_empty[buf]);  // accumulator drained
        sg_done_add(s_done);           // ... and this warp's stores issued (the publisher counts the warps)
      }
      if (tid == 0) SG_TRACE(it, 6);
      ++it;
    });
  } else if (warp == W_PUB) {
    // ===================== publisher: makes a finished tile visible to the other CTAs =====================
    // Waiting for the tile's stores to be performed (the gpu-scope release below) costs ~1.3 us; on the epilogue warps that
    // was 17 % of an epilogue-bound item (profiles/r2d_stage_trace.log).  Here one extra lane pays it: the epilogue warps'
    // stores happen-before their release increments of a shared-memory counter, those before this lane's acquire load of it,
    // and its gpu-scope release is cumulative.
    if (lane == 0) {
      int it = 0;
      walk([&](int conv, const ConvParams& p, int b, int ntile, int mtile) {
        (void)p; (void)ntile;
        while (sg_done_ld(s_done) < (unsigned int)(EW * (it + 1))) {
        }
        asm volatile("fence.proxy.async;" ::: "memory");
        sg_red_release(sp.flags + ((size_t)conv * sp.zb + b) * sp.tiles_max + mtile);
        ++it;
      });
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// stage_pair_kernel — the same dataflow launch with a whole ResBlock step (c1 -> LeakyReLU -> c2 + residual) per item,
// for the 32-channel stage, which is bound by its epilogues' stores (3.3-5.4 us per 16 K-element tile against 3.2 us of
// MMAs, profiles/r2d_stage_trace.log): the inner activation T goes to SHARED memory (rbpair_umma.cuh: halo recompute, tiles
// advance by MT*128 - (K-1) rows), so only x_new is stored and nothing is re-read.  Two epilogue groups of 8 warps work at the
// same time: group 1 turns c1's accumulators into the T image (fast: no global traffic), group 2 runs the shared epilogue
// text on c2's accumulators; the MMA warp issues c1(i), c2(i), c1(i+1), ...  Items = (dilation step, branch, tile); a tile
// waits for tiles i-1, i, i+1 of the previous step of its branch; every image is written once per launch (x ping-pongs
// U -> R -> T), so the waits only have to order reads after writes.
// ------------------------------------------------------------------------------------------------
template <int MT>
__global__ void __launch_bounds__(SG_EW * 32 + 128, 1) stage_pair_kernel(const __grid_constant__ StageParams sp) {
  using namespace umma;
  using Cfg = RbPairCfg<MT>;
  constexpr int NSPLIT = 2, NACC = 2, AST = Cfg::AST, WST = Cfg::WST, G = Cfg::G, RA = Cfg::RA, RT = Cfg::RT;
  constexpr int EWG = SG_EW / 2;  // warps per epilogue group
  constexpr int W_MMA = SG_EW, W_LOAD = SG_EW + 1, W_ALOAD = SG_EW + 2, W_PUB = SG_EW + 3;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const ConvParams& p0 = sp.convs[0];
  const int NT = p0.um_nt, C = p0.Cin;
  const int nch = (C + UM_BKC - 1) / UM_BKC, KC = C / 8;
  const uint32_t a_split_bytes = UM_KCH * RA * 16, a_stage = NSPLIT * a_split_bytes;
  const uint32_t t_split_bytes = (uint32_t)KC * RT * 16;
  const uint32_t w_stage = NSPLIT * UM_KCH * NT * 16, w_ring = G * w_stage;
  unsigned char* sA = smem_raw;
  unsigned char* sT = sA + AST * a_stage;
  unsigned char* sW = sT + NSPLIT * t_split_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + WST * w_ring);
  uint64_t* a_full = bars;           // [AST <= 2]
  uint64_t* a_empty = bars + 2;      // [AST <= 2]
  uint64_t* w_full = bars + 4;       // [WST <= 8]
  uint64_t* w_empty = bars + 12;     // [WST <= 8]
  uint64_t* acc1_full = bars + 20;   // c1 MMAs done -> group 1
  uint64_t* acc1_empty = bars + 21;  // group 1 has read acc1
  uint64_t* acc2_full = bars + 22;   // c2 MMAs done -> group 2
  uint64_t* acc2_empty = bars + 23;  // group 2 has read acc2 and issued its stores (the publisher waits for it too)
  uint64_t* t_full = bars + 24;      // group 1 has written the T image -> c2
  uint64_t* t_empty = bars + 25;     // c2 MMAs have read the T image
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  unsigned int* s_done = reinterpret_cast<unsigned int*>(bars + 27);  // group-2 warps that have finished an item (monotonic)
  float* s_bias1_all = reinterpret_cast<float*>(bars + 28);  // [n_convs][C]
  const int n_convs = sp.n_levels * sp.n_cols;
  float* s_bias2_all = s_bias1_all + n_convs * C;            // [n_convs][C]
  const float *s_inv = nullptr, *s_anb = nullptr, *s_ans = nullptr;  // named by the (dead here) coupling branch of the epilogue text

  const int tid = threadIdx.x, warp_g = tid >> 5, lane = tid & 31;
  const int seg0 = sp.seg_off[blockIdx.x], seg1 = sp.seg_off[blockIdx.x + 1];
  if (seg0 == seg1) return;
  constexpr uint32_t ACC_COLS = 256;

  if (tid == 0) {
    for (int i = 0; i < AST; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < WST; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, EWG);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, EWG);
    mbar_init(t_full, EWG);
    mbar_init(t_empty, 1);
    *s_done = 0;
    fence_barrier_init();
  }
  if (warp_g == W_MMA) tmem_alloc(tmem_slot, 512);
  for (int cv = warp_g; cv < n_convs; cv += (SG_EW * 32 + 128) / 32) {
    const float *b1 = sp.convs1[cv].bias, *b2 = sp.convs[cv].bias;
    for (int i = lane; i < C; i += 32) {
      s_bias1_all[cv * C + i] = b1[i];
      s_bias2_all[cv * C + i] = b2[i];
    }
  }
  // the RP_TPAD rows behind the computed T rows are only ever read for output rows that are discarded; keep them finite
  for (int i = tid; i < NSPLIT * KC * RP_TPAD; i += SG_EW * 32 + 128) {
    const int pl_kc = i / RP_TPAD, r = i - pl_kc * RP_TPAD;
    *reinterpret_cast<uint4*>(sT + ((size_t)pl_kc * RT + MT * 128 + r) * 16) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  auto walk = [&](auto&& body) {
    for (int lv = 0; lv < sp.n_levels; ++lv)
      for (int sg = seg0; sg < seg1; ++sg) {
        const StageSeg seg = sp.segs[sg];
        const int conv = lv * sp.n_cols + seg.col;
        for (int mtile = seg.t_begin; mtile < seg.t_end; ++mtile) body(conv, sp.convs1[conv], sp.convs[conv], seg.b, mtile);
      }
  };

  if (warp_g == W_LOAD) {
    // ===================== weight producer: per item the (chunk, tap) blocks of c1, then those of c2 =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      walk([&](int conv, const ConvParams& p1, const ConvParams& p, int b, int mtile) {
        (void)conv; (void)b; (void)mtile;
        for (int cv = 0; cv < 2; ++cv) {
          const ConvParams& q = cv ? p : p1;
          const unsigned char* gw = reinterpret_cast<const unsigned char*>(q.wu);
          const int Kq = q.K;
          for (int c = 0; c < nch; ++c)
            for (int k0 = 0; k0 < Kq; k0 += G) {
              const uint32_t bytes = (uint32_t)min(G, Kq - k0) * w_stage;
              mbar_wait(&w_empty[s], ph);
              mbar_arrive_expect_tx(&w_full[s], bytes);
              bulk_g2s(sW + (size_t)s * w_ring, gw + ((size_t)c * Kq + k0) * w_stage, bytes, &w_full[s]);
              if (++s == WST) { s = 0; ph ^= 1; }
            }
        }
      });
    }
  } else if (warp_g == W_ALOAD) {
    // ===================== x tile: waits for the producer tiles, then bulk-copies =====================
    int as = 0;
    uint32_t aph = 1;
    walk([&](int conv, const ConvParams& p1, const ConvParams& p, int b, int mtile) {
      const int S = MT * 128 - (p.K - 1);
      const int d = sp.dep[conv];
      if (d >= 0) {
        if (lane == 0) {
          const int Lb = p.lens[b] * p.len_mul;
          const int m_tiles = (Lb + S - 1) / S;
          const unsigned int* fl = sp.flags + ((size_t)d * sp.zb + b) * sp.tiles_max;
          const int i0 = max(mtile - 1, 0), i1 = min(mtile + 1, m_tiles - 1);
          for (int ii = i0; ii <= i1; ++ii) {
            unsigned int spins = 0;
            while (sg_ld_acquire(fl + ii) < 1u) {
              if (++spins > (1u << 26)) {  // seconds: a lost publication must fail loudly, not hang the device
                printf("stage_pair: tile %d of step %d (item %d) was never published (CTA %d waits for step %d tile %d)\n", ii, d, b,
                       blockIdx.x, conv, mtile);
                __trap();
              }
            }
          }
        }
        __syncwarp();
        asm volatile("fence.proxy.async;" ::: "memory");  // (each issuing lane) other CTAs' generic-proxy writes -> our async-proxy reads
      }
      const unsigned char* xp = reinterpret_cast<const unsigned char*>(p1.xp) + (size_t)b * p1.xp_bs;
      const long long row0 = (long long)mtile * S - p.pad_left - p1.pad_left + UM_PADL;
      issue_a_chunks<NSPLIT, AST>(lane, as, aph, nch, a_empty, a_full, sA, a_stage, a_split_bytes, RA, xp, p1.xp_kc, p1.xp_rows, row0,
                                  MT * 128 + (p1.K - 1) * p1.dil);
    });
  } else if (warp_g == W_MMA) {
    // ===================== MMA issuer: per item c1 (x ring -> acc1), then c2 (T image -> acc2) =====================
    const uint32_t idesc1 = make_idesc_bf16_f32(128, NT), idesc2 = make_idesc_bf16_f32(128, 2 * NT);
    const uint64_t a_desc0 = make_smem_desc(smem_u32(sA), RA * 16, 128);
    const uint64_t t_desc0 = make_smem_desc(smem_u32(sT), RT * 16, 128);
    const uint64_t b_desc0 = make_smem_desc(smem_u32(sW), NSPLIT * NT * 16, 128);
    const uint32_t a_ks = (2 * RA * 16) >> 4, t_ks = (2 * RT * 16) >> 4, b_ks = (2 * NSPLIT * NT * 16) >> 4;
    const uint32_t a_sp = a_split_bytes >> 4, t_sp = t_split_bytes >> 4;
    const uint32_t a_mt = (128 * 16) >> 4;
    int ws = 0, as = 0;
    uint32_t wph = 0, aph = 0;
    auto run_conv = [&](bool second, int Kq, int dq, uint32_t tm, uint32_t roff) {
      for (int c = 0; c < nch; ++c) {
        uint64_t a_c;
        if (!second) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          a_c = a_desc0 + ((as * a_stage) >> 4) + roff;
        } else {
          a_c = t_desc0 + (uint32_t)((c * UM_KCH * RT * 16) >> 4);
        }
        const uint32_t ks_step = second ? t_ks : a_ks, sp_step = second ? t_sp : a_sp;
        for (int k0 = 0; k0 < Kq; k0 += G) {
          mbar_wait(&w_full[ws], wph);
          tc_fence_after();
          const int cnt = min(G, Kq - k0);
          if (elect_one()) {
            for (int kk = 0; kk < cnt; ++kk) {
              const int k = k0 + kk;
              const uint64_t a_k = a_c + (uint32_t)(k * dq);
              const uint64_t b_k = b_desc0 + ((ws * w_ring + kk * w_stage) >> 4);
              const uint32_t first = (c == 0 && k == 0) ? 0u : 1u;
#pragma unroll
              for (int ks = 0; ks < UM_BKC / 16; ++ks) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  const uint64_t a0 = a_k + ks * ks_step + mt * a_mt, b0 = b_k + ks * b_ks;
                  const uint32_t acc0 = ks == 0 ? first : 1u;
                  const uint32_t d0 = tm + (uint32_t)(mt * NACC) * NT;
                  mma_bf16_ss(d0, a0, b0, idesc2, acc0);               // hi x [hi|lo] -> [c0|c1]
                  mma_bf16_ss(d0 + NT, a0 + sp_step, b0, idesc1, 1u);  // lo x hi      -> c1
                }
              }
            }
            mma_commit(&w_empty[ws]);
          }
          __syncwarp();
          if (++ws == WST) { ws = 0; wph ^= 1; }
        }
        if (!second) {
          if (elect_one()) mma_commit(&a_empty[as]);
          __syncwarp();
          if (++as == AST) { as = 0; aph ^= 1; }
        }
      }
    };
    int it = 0;
    walk([&](int conv, const ConvParams& p1, const ConvParams& p, int b, int mtile) {
      (void)conv; (void)b;
      const int S = MT * 128 - (p.K - 1);
      mbar_wait(acc1_empty, (it & 1) ^ 1);  // group 1 has drained the previous item's c1 accumulators
      tc_fence_after();
      run_conv(false, p1.K, p1.dil, tmem, a_row_off((long long)mtile * S - p.pad_left - p1.pad_left + UM_PADL));
      if (elect_one()) mma_commit(acc1_full);
      __syncwarp();
      mbar_wait(acc2_empty, (it & 1) ^ 1);  // group 2 has drained the previous item's c2 accumulators
      mbar_wait(t_full, it & 1);            // group 1 has written this item's T image
      tc_fence_after();
      run_conv(true, p.K, 1, tmem + ACC_COLS, 0u);
      if (elect_one()) {
        mma_commit(t_empty);
        mma_commit(acc2_full);
      }
      __syncwarp();
      ++it;
    });
  } else if (warp_g < EWG) {
    // ===================== epilogue group 1: T = lrelu(c1 + bias), zero outside [0, len) -> hi/lo image in shared memory =====
    const int warp = warp_g;
    const int quad = warp & 3, wset = warp >> 2;
    constexpr int NWS = EWG / 4;
    const int m = quad * 32 + lane;
    const int ngroups = (NT + 31) / 32;
    int it = 0;
    walk([&](int conv, const ConvParams& p1, const ConvParams& p, int b, int mtile) {
      const int S = MT * 128 - (p.K - 1);
      const int t0 = mtile * S, pad2 = p.pad_left;
      const int Lb = p.lens[b] * p.len_mul;
      const float* s_bias1 = s_bias1_all + conv * C;
      const float sl1 = p1.yp_slope;
      mbar_wait(acc1_full, it & 1);
      mbar_wait(t_empty, (it & 1) ^ 1);  // c2 of the previous item has read the image
      tc_fence_after();
      for (int task = wset; task < MT * ngroups; task += NWS) {
        const int mt = MT == 1 ? 0 : task % MT;
        const int g = MT == 1 ? task : task / MT;
        const int r = mt * 128 + m;       // row of the T image
        const int tT = t0 - pad2 + r;     // its time position
        const bool ok = tT >= 0 && tT < Lb;
        const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * 32);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float acc[16];
          {
            uint32_t r1[16], r0[16];
            tmem_ld16(lane_addr + (uint32_t)((mt * NACC + 1) * NT) + hf * 16, r1);  // smallest magnitude class first
            tmem_ld16(lane_addr + (uint32_t)((mt * NACC) * NT) + hf * 16, r0);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(r1[j]) + __uint_as_float(r0[j]);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float v = acc[j] + s_bias1[g * 32 + hf * 16 + j];
            v = v > 0.f ? v : v * sl1;  // the LeakyReLU in front of c2, exactly as the unfused c1 epilogue stores it
            acc[j] = ok ? v : 0.f;
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (g * 32 + hf * 16 + q * 8 >= NT) continue;
            uint4 h2[2];
            split8<2>(acc + q * 8, h2);
            unsigned char* o = sT + ((size_t)(g * 4 + hf * 2 + q) * RT + r) * 16;
            *reinterpret_cast<uint4*>(o) = h2[0];
            *reinterpret_cast<uint4*>(o + t_split_bytes) = h2[1];
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(acc1_empty);
        mbar_arrive(t_full);
      }
      ++it;
    });
  } else if (warp_g < SG_EW) {
    // ===================== epilogue group 2: the shared epilogue on c2's accumulators =====================
    const int warp = warp_g - EWG;  // the epilogue text distributes its tasks over EW warps numbered from 0
    constexpr int EW = EWG;
    int it = 0;
    walk([&](int conv, const ConvParams& p1, const ConvParams& p, int b, int mtile) {
      (void)p1;
      const int S = MT * 128 - (p.K - 1);
      const int n0 = 0, t0 = mtile * S;
      const int Lb = p.lens[b] * p.len_mul;
      const float* s_bias = s_bias2_all + conv * C;
      mbar_wait(acc2_full, it & 1);
      tc_fence_after();
      {
        const uint32_t tm = tmem + ACC_COLS;
#define UM_EPI_SKIP_ROW(t) ((t) >= t0 + S && (t) < Lb)
#define UM_EPI_LD16(ptr) __ldcg(reinterpret_cast<const uint4*>(ptr))
#include "conv_umma_epilogue.inc"
#undef UM_EPI_LD16
#undef UM_EPI_SKIP_ROW
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(acc2_empty);
        sg_done_add(s_done);
      }
      ++it;
    });
  } else if (warp_g == W_PUB) {
    // ===================== publisher =====================
    if (lane == 0) {
      int it = 0;
      walk([&](int conv, const ConvParams& p1, const ConvParams& p, int b, int mtile) {
        (void)p1; (void)p;
        while (sg_done_ld(s_done) < (unsigned int)(EWG * (it + 1))) {
        }
        asm volatile("fence.proxy.async;" ::: "memory");
        sg_red_release(sp.flags + ((size_t)conv * sp.zb + b) * sp.tiles_max + mtile);
        ++it;
      });
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_g == W_MMA) tmem_dealloc(tmem, 512);
}

}  // namespace b2
