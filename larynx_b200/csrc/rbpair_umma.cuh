// rbpair_umma.cuh — one HiFi-GAN ResBlock1 step  x <- x + c2(lrelu(c1(lrelu(x))))  (hifi_gan/models.py:91-98) as ONE
// persistent tcgen05 kernel for the narrow stages (C <= 64: stages 3 and 4 of universal_large).
//
// Why: with one launch per conv the inner activation T = lrelu(c1(.)) makes a round trip through L2/HBM as a hi/lo operand
// image (4 B per element written by c1's epilogue, 4 B x (1 + halo) read by c2's producer).  In the narrow stages that
// traffic — not the tensor pipe — bounds the launches (stage 4: ~170 MB per launch at 51 us; tensor pipe 13-16 % active,
// profiles/r1j_vocoder_launches_tensor_active.csv).  Here T never leaves the SM:
//   c1 : MT x 128 rows of T (time  t0 - p2 .. t0 - p2 + MT*128) from the x tile (bulk-copied operand image, as conv_umma)
//        -> accumulators in TMEM columns [0, 256)
//   epi1: TMEM -> registers -> + bias, LeakyReLU(0.1), zero outside [0, len) -> hi/lo split -> T operand image in SHARED memory
//   c2 : MT x 128 output rows (time t0 ..) straight from that image (tap k = descriptor start + k rows) -> TMEM [256, 512)
//   epi2: the shared epilogue text (conv_umma_epilogue.inc): + bias + residual (from x's image) -> fp32 and/or the operand
//        image of the next step.
// c2's last K-1 rows of a tile would need T rows this tile did not compute: tiles advance by S = MT*128 - (K-1) rows and
// those rows are left to the next tile (halo recompute: (K-1) / (MT*128) = 2-4 % of the MMAs).
// Because tile i writes x_new[t0 .. t0+S) while other CTAs still read x around it, input and output images must differ
// (engine.cu ping-pongs U -> R -> T images, the T image being free in this mode).
// Pipeline per CTA (persistent over the tiles of its (branch, item)): the MMA warp issues c2(i) then c1(i+1); the 16 epilogue
// warps run epi1(i) -> epi2(i) -> epi1(i+1) ...; c1(i+1) overlaps epi2(i).  Weights of c1 then c2 stream through one ring.
#pragma once
#include "conv_umma.cuh"

namespace b2 {

struct PairBatch {
  int n = 1;   // branches (the three MRF ResBlocks of a stage)
  int zb = 1;  // batch items per branch
  int gx[3] = {0, 0, 0};  // CTAs (along x) that share the time tiles of branch i
  ConvParams c1[3], c2[3];
};

constexpr int RP_EW = 16;
constexpr int RP_THREADS = RP_EW * 32 + 96;
constexpr int RP_TPAD = 16;  // extra rows of the T image read by c2's taps past the last computed row (>= K - 1)

template <int MT>
struct RbPairCfg {
  static constexpr int AST = MT >= 4 ? 1 : 2;  // x ring stages (32-channel chunks)
  static constexpr int WST = 4;                // weight ring slots of G taps
  static constexpr int G = 2;
  static constexpr int RA = MT * 128 + UM_HALO;
  static constexpr int RT = MT * 128 + RP_TPAD;
  static size_t smem_bytes(int NT, int C) {
    const size_t a = (size_t)AST * 2 * UM_KCH * RA * 16;
    const size_t t = (size_t)2 * (C / 8) * RT * 16;
    const size_t w = (size_t)WST * G * 2 * UM_KCH * NT * 16;
    return a + t + w + 512 + 4 * (2 * 64);
  }
};

template <int MT>
__global__ void __launch_bounds__(RP_THREADS, 1) rbpair_umma_kernel(const __grid_constant__ PairBatch pb) {
  using namespace umma;
  using Cfg = RbPairCfg<MT>;
  constexpr int NSPLIT = 2, NACC = 2, EW = RP_EW, AST = Cfg::AST, WST = Cfg::WST, G = Cfg::G;
  constexpr int RA = Cfg::RA, RT = Cfg::RT;
  constexpr int W_MMA = EW, W_LOAD = EW + 1, W_ALOAD = EW + 2;
  const int branch = blockIdx.z / pb.zb;
  const ConvParams& p1 = pb.c1[branch];
  const ConvParams& p = pb.c2[branch];  // named p: the shared epilogue text reads it
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int NT = p.um_nt, C = p.Cin;  // one N tile covers all C output channels of both convs
  const int nch = (C + UM_BKC - 1) / UM_BKC, KC = C / 8;
  const uint32_t a_split_bytes = UM_KCH * RA * 16, a_stage = NSPLIT * a_split_bytes;
  const uint32_t t_split_bytes = (uint32_t)KC * RT * 16;
  const uint32_t w_stage = NSPLIT * UM_KCH * NT * 16, w_ring = G * w_stage;
  unsigned char* sA = smem_raw;
  unsigned char* sT = sA + AST * a_stage;
  unsigned char* sW = sT + NSPLIT * t_split_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + WST * w_ring);
  uint64_t* a_full = bars;         // [AST <= 2]
  uint64_t* a_empty = bars + 2;    // [AST <= 2]
  uint64_t* w_full = bars + 4;     // [WST <= 8]
  uint64_t* w_empty = bars + 12;   // [WST <= 8]
  uint64_t* acc1_full = bars + 20;   // c1 MMAs done -> epi1
  uint64_t* acc1_empty = bars + 21;  // epi1 has read acc1 -> c1 of the next tile
  uint64_t* acc2_full = bars + 22;   // c2 MMAs done -> epi2
  uint64_t* acc2_empty = bars + 23;  // epi2 has read acc2
  uint64_t* t_full = bars + 24;      // epi1 has written the T image -> c2
  uint64_t* t_empty = bars + 25;     // c2 MMAs have read the T image -> epi1 of the next tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  static_assert(WST <= 8 && AST <= 2, "barrier slots");
  float* s_bias1 = reinterpret_cast<float*>(bars + 28);  // [64]
  float* s_bias = s_bias1 + 64;                          // [64] (c2; the shared epilogue text reads s_bias)
  const float *s_inv = nullptr, *s_anb = nullptr, *s_ans = nullptr;  // named by the (dead here) coupling branch of the epilogue text

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z % pb.zb;
  const int n0 = 0;
  const int Lb = p.lens ? p.lens[b] * p.len_mul : p.Lin;
  const int K1 = p1.K, dil1 = p1.dil, pad1 = p1.pad_left, K2 = p.K, pad2 = p.pad_left;
  const int S = MT * 128 - (K2 - 1);  // output rows a tile contributes
  const int m_tiles = (Lb + S - 1) / S;
  const int gstride = pb.gx[branch] > 0 ? pb.gx[branch] : (int)gridDim.x;
  if ((int)blockIdx.x >= m_tiles || (int)blockIdx.x >= gstride) return;
  constexpr uint32_t ACC_COLS = 256;  // columns of one accumulator set (NACC * MT * NT <= 256)

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
    mbar_init(acc1_empty, EW);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, EW);
    mbar_init(t_full, EW);
    mbar_init(t_empty, 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < NT; i += RP_THREADS) {
    s_bias1[i] = i < p1.Cout ? p1.bias[i] : 0.f;
    s_bias[i] = i < p.Cout ? p.bias[i] : 0.f;
  }
  // the RP_TPAD rows behind the computed T rows are only ever read for output rows that are discarded; keep them finite
  for (int i = tid; i < NSPLIT * KC * RP_TPAD; i += RP_THREADS) {
    const int pl_kc = i / RP_TPAD, r = i - pl_kc * RP_TPAD;
    *reinterpret_cast<uint4*>(sT + ((size_t)pl_kc * RT + MT * 128 + r) * 16) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == W_LOAD) {
    // ===================== weight producer: per tile the (chunk, tap) blocks of c1, then those of c2 =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int mtile = blockIdx.x; mtile < m_tiles; mtile += gstride)
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
    }
  } else if (warp == W_ALOAD) {
    // ===================== x tile by bulk copy from the producer's plane image =====================
    asm volatile("griddepcontrol.wait;" ::: "memory");  // activations are written by the previous grid
    const unsigned char* xp = reinterpret_cast<const unsigned char*>(p1.xp) + (size_t)b * p1.xp_bs;
    int as = 0;
    uint32_t aph = 1;
    for (int mtile = blockIdx.x; mtile < m_tiles; mtile += gstride) {
      const long long row0 = (long long)mtile * S - pad2 - pad1 + UM_PADL;
      issue_a_chunks<NSPLIT, AST>(lane, as, aph, nch, a_empty, a_full, sA, a_stage, a_split_bytes, RA, xp, p1.xp_kc, p1.xp_rows, row0,
                                  MT * 128 + (K1 - 1) * dil1);
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer: c1(0), then per tile c2(i), c1(i+1) =====================
    const uint32_t idesc1 = make_idesc_bf16_f32(128, NT), idesc2 = make_idesc_bf16_f32(128, 2 * NT);
    const uint64_t a_desc0 = make_smem_desc(smem_u32(sA), RA * 16, 128);
    const uint64_t t_desc0 = make_smem_desc(smem_u32(sT), RT * 16, 128);
    const uint64_t b_desc0 = make_smem_desc(smem_u32(sW), NSPLIT * NT * 16, 128);
    const uint32_t a_ks = (2 * RA * 16) >> 4, t_ks = (2 * RT * 16) >> 4, b_ks = (2 * NSPLIT * NT * 16) >> 4;
    const uint32_t a_sp = a_split_bytes >> 4, t_sp = t_split_bytes >> 4;
    const uint32_t a_mt = (128 * 16) >> 4;
    int ws = 0, as = 0;
    uint32_t wph = 0, aph = 0;
    // one conv of one tile: A = x ring stages (c1) or the T image (c2)
    auto run_conv = [&](bool second, uint32_t tm, uint32_t roff) {
      const int Kq = second ? K2 : K1, dq = second ? 1 : dil1;
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
                  mma_bf16_ss(d0, a0, b0, idesc2, acc0);             // hi x [hi|lo] -> [c0|c1]
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
    for (int mtile = blockIdx.x; mtile < m_tiles; mtile += gstride, ++it) {
      // c1(it): its accumulators were drained by epi1(it-1)
      mbar_wait(acc1_empty, (it & 1) ^ 1);
      tc_fence_after();
      run_conv(false, tmem, a_row_off((long long)mtile * S - pad2 - pad1 + UM_PADL));
      if (elect_one()) mma_commit(acc1_full);
      __syncwarp();
      // c2(it): needs the T image of epi1(it) and its accumulators drained by epi2(it-1)
      mbar_wait(acc2_empty, (it & 1) ^ 1);
      mbar_wait(t_full, it & 1);
      tc_fence_after();
      run_conv(true, tmem + ACC_COLS, 0u);
      if (elect_one()) {
        mma_commit(t_empty);
        mma_commit(acc2_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue warps =====================
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the residual image is written by the previous grid
    const int quad = warp & 3, wset = warp >> 2;
    constexpr int NWS = EW / 4;
    const int m = quad * 32 + lane;
    const int ngroups = (NT + 31) / 32;
    int it = 0;
    for (int mtile = blockIdx.x; mtile < m_tiles; mtile += gstride, ++it) {
      const int t0 = mtile * S;
      // ---------- epi1: T = lrelu(c1 + bias, slope of c2's input), zero outside [0, len) -> hi/lo image in shared memory
      mbar_wait(acc1_full, it & 1);
      mbar_wait(t_empty, (it & 1) ^ 1);  // c2 of the previous tile has read the image
      tc_fence_after();
      for (int task = wset; task < MT * ngroups; task += NWS) {
        const int mt = MT == 1 ? 0 : task % MT;
        const int g = MT == 1 ? task : task / MT;
        const int r = mt * 128 + m;       // row of the T image
        const int tT = t0 - pad2 + r;     // its time position
        const bool ok = tT >= 0 && tT < Lb;
        float acc[32];
        {
          uint32_t raw[32];
          const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * 32);
          tmem_ld32(lane_addr + (uint32_t)((mt * NACC + 1) * NT), raw);  // smallest magnitude class first
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(raw[j]);
          tmem_ld32(lane_addr + (uint32_t)((mt * NACC) * NT), raw);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(raw[j]);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float v = acc[j] + s_bias1[g * 32 + j];
          v = v > 0.f ? v : v * p1.yp_slope;  // the LeakyReLU in front of c2, exactly as the unfused c1 epilogue stores it
          acc[j] = ok ? v : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (g * 32 + q * 8 >= NT) continue;
          uint4 h2[2];
          split8<2>(acc + q * 8, h2);
          unsigned char* o = sT + ((size_t)(g * 4 + q) * RT + r) * 16;
          *reinterpret_cast<uint4*>(o) = h2[0];
          *reinterpret_cast<uint4*>(o + t_split_bytes) = h2[1];
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(acc1_empty);
        mbar_arrive(t_full);
      }
      // ---------- epi2: the shared epilogue on c2's accumulators; rows [t0 + S, t0 + MT*128) belong to the next tile
      mbar_wait(acc2_full, it & 1);
      tc_fence_after();
      {
        const uint32_t tm = tmem + ACC_COLS;
#define UM_EPI_SKIP_ROW(t) ((t) >= t0 + S && (t) < Lb)
#include "conv_umma_epilogue.inc"
#undef UM_EPI_SKIP_ROW
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, 512);
}

}  // namespace b2
