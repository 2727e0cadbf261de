// enc_chain.cuh — the GlowTTS text encoder + duration-predictor convs (glow_tts/models.py:118-140, 39-49;
// attentions.py:62-74, 214-264, 375-383; layers.py:73-80) as ONE persistent cooperative launch driven by a step table.
//
// Why: at batch 1 the encoder is ~50 dependent launches of 2-12 CTAs (0.96 ms for 128 phonemes, 15 MFLOP per phoneme);
// every launch pays launch gap + TMEM allocation + barrier init + pipeline fill.  Here a group of (time tiles x 12) CTAs
// owns one utterance for the whole encoder; steps are separated by a per-utterance barrier in global memory (as in
// flow_chain.cuh).  Step types:
//   CONV    Conv1d as tcgen05 implicit GEMM on 3-way split bf16 operand images (fp32-equivalent products: the encoder
//           feeds the bit-exact alignment).  Wide layers (ffn1: 12 N tiles, qkv: 9) spread over the CTA rows by output
//           channels; narrow, deep ones (ffn2: K = 768 x 3 taps on 3 N tiles, the k5 prenet convs, the duration convs) are
//           K-SPLIT over the CTA rows and write raw fp32 partial sums;
//   REDUCE  sum of the K-split partials + bias + residual, then the channel LayerNorm (warp-shuffle / shared-memory
//           statistics) -> fp32 residual stream + operand image of the next conv.  Every residual add of the encoder sits
//           in front of a LayerNorm, so conv epilogues never read activations written by other CTAs;
//   ATTN    relative-position attention, one head per CTA, on tcgen05 (the code of attn_umma.cu as a step: S = Q [K; Ek]^T,
//           thread-local softmax, O = [P, P_band] [V; Ev]); its operands overlay the idle conv rings.
// The step table lives in shared memory (copied once at kernel start).
#pragma once
#include "conv_umma.cuh"
#include "enc_chain_params.cuh"

namespace b2 {

constexpr int EC_EW = 8;
constexpr int EC_THREADS = EC_EW * 32 + 96;
constexpr int EC_NT_MAX = 64;
constexpr int EC_HALO = 16;
constexpr int EC_RA = 128 + EC_HALO;
constexpr int EC_AST = 3, EC_WST = 3;
constexpr int EC_G = 3;        // consecutive (chunk, tap) weight blocks per ring slot: one wait + one commit per slot
constexpr uint32_t EC_A_SPLIT = UM_KCH * EC_RA * 16;            // 9216
constexpr uint32_t EC_A_STAGE = 3 * EC_A_SPLIT;                 // 27648
constexpr uint32_t EC_W_BLOCK = 3 * UM_KCH * EC_NT_MAX * 16;    // 12288
constexpr uint32_t EC_W_SLOT = EC_G * EC_W_BLOCK;               // 36864
constexpr uint32_t EC_DATA = 193536;  // = the attention operands (below) >= EC_AST * EC_A_STAGE + EC_WST * EC_W_SLOT
static_assert(EC_AST * EC_A_STAGE + EC_WST * EC_W_SLOT <= EC_DATA, "rings must fit the data region");
// attention operand geometry (attn_umma.cu)
constexpr int EA_ROWS = 128, EA_DK = 96, EA_KC = EA_DK / 8, EA_NS = 3, EA_XR = EA_ROWS + 16;
constexpr uint32_t EA_QPLANE = EA_KC * EA_ROWS * 16;        // 24576
constexpr uint32_t EA_XPLANE = EA_KC * EA_XR * 16;          // 27648
constexpr uint32_t EA_PPLANE = (EA_XR / 8) * EA_ROWS * 16;  // 36864
static_assert(EA_NS * EA_XPLANE + EA_NS * EA_PPLANE <= EC_DATA, "attention operands must fit the conv rings they overlay");
constexpr size_t EC_SMEM = EC_DATA + 512 + 4 * (64 + 8 * 33 + 3 * 128 * 9 + 4 * 128) + sizeof(EncStep) * EC_MAX_STEPS + 16;
static_assert(EC_SMEM <= 227 * 1024, "enc_chain: shared memory budget");

// epilogue parameters of conv_umma_epilogue.inc for this kernel (EPI_PLAIN only; register resident)
struct EncEpiP {
  static constexpr int epi = EPI_PLAIN, pix = 1, yp_nsplit = 3, split = 0, first = 0, half = 0;
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
  static constexpr float* tap = nullptr;
  static constexpr unsigned long long* dbg = nullptr;
  void* yp;
  long long yp_bs;
  int yp_kc, yp_rows;
  float* y;
  long long y_bs;
  int y_ld, act, Cout;
};

__device__ __forceinline__ unsigned int ec_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ec_red_release(unsigned int* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void ec_epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(EC_EW * 32) : "memory"); }

// CTAs of one utterance that take part in a step (the barrier count)
__device__ __forceinline__ unsigned int ec_expect(const EncStep& st, unsigned int m_tiles) {
  if (st.type == ES_CONV) {
    unsigned int n = 0;
    for (int i = 0; i < st.n_conv; ++i) n += (unsigned int)(st.conv[i].ntiles * st.conv[i].ksplit);
    return n * m_tiles;
  }
  if (st.type == ES_REDUCE) return (unsigned int)EC_YT * m_tiles;
  return (unsigned int)st.att.n_heads;
}
// the conv op of a step that CTA row y works on, or nullptr
__device__ __forceinline__ const EncConvOp* ec_my_op(const EncStep& st, int y) {
  for (int i = 0; i < st.n_conv; ++i) {
    const EncConvOp& o = st.conv[i];
    if (y >= o.y0 && y < o.y0 + o.ntiles * o.ksplit) return &o;
  }
  return nullptr;
}

__device__ __forceinline__ void ec_wait_step(const unsigned int* cnt, int s, unsigned int expect, int b) {
  unsigned int spins = 0;
  while (ec_ld_acquire(cnt + s) < expect) {
    if (++spins > (1u << 25)) {  // seconds: a lost arrival must fail loudly, not hang the device
      printf("enc_chain: barrier %d of item %d timed out (CTA %d,%d,%d)\n", s, b, blockIdx.x, blockIdx.y, blockIdx.z);
      __trap();
    }
  }
}

#define EC_TRACE(S, ev)                                                                                         \
  do {                                                                                                          \
    if (fp.trace && blockIdx.x == 0 && blockIdx.z == 0 && b == 0) fp.trace[((size_t)(S)*EC_YT + y) * 8 + (ev)] = gtimer(); \
  } while (0)

__global__ void __launch_bounds__(EC_THREADS, 1) enc_chain_kernel(const __grid_constant__ EncChainParams fp) {
  using namespace umma;
  constexpr int MT = 1, NSPLIT = 3, NACC = 3, EW = EC_EW;
  constexpr int RA = EC_RA;
  constexpr int W_MMA = EW, W_LOAD = EW + 1, W_ALOAD = EW + 2;
  constexpr uint32_t a_split_bytes = EC_A_SPLIT, a_stage = EC_A_STAGE, w_slot = EC_W_SLOT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sA = smem_raw;
  unsigned char* sW = smem_raw + EC_AST * a_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + EC_DATA);
  uint64_t* a_full = bars;           // [AST]
  uint64_t* a_empty = bars + 4;      // [AST]
  uint64_t* w_full = bars + 8;       // [WST <= 12]
  uint64_t* w_empty = bars + 20;     // [WST <= 12]
  uint64_t* acc_full = bars + 32;    // MMA -> epilogue
  uint64_t* acc_empty = bars + 33;   // epilogue -> MMA
  uint64_t* at_qk = bars + 34;       // attention: Q, K' landed
  uint64_t* at_s = bars + 35;        //            S accumulated (Q, K' dead)
  uint64_t* at_v = bars + 36;        //            V' landed
  uint64_t* at_p = bars + 37;        //            P' image written (one arrival per epilogue thread)
  uint64_t* at_o = bars + 38;        //            O accumulated (operands dead: the rings may refill)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);
  static_assert(EC_WST <= 12 && EC_AST <= 4, "barrier slots");
  float* s_bias = reinterpret_cast<float*>(bars + 64);  // [64]
  const float *s_inv = nullptr, *s_anb = nullptr, *s_ans = nullptr;  // named by the (dead here) coupling branch of the epilogue text
  float* s_red = s_bias + 64;                           // [8][33] LayerNorm partial sums
  float* s_relk = s_red + 8 * 33;                       // [2][128][9] relative-key logits of each query (a copy per column half)
  float* s_band = s_relk + 2 * 128 * 9;                 // [128][9] scaled, masked logits on the band |j - i| <= window
  float* s_mx = s_band + 128 * 9;                       // [2][128] row maximum of each column half
  float* s_sum = s_mx + 2 * 128;                        // [2][128] row sum of each column half
  EncStep* s_steps = reinterpret_cast<EncStep*>((reinterpret_cast<uintptr_t>(s_sum + 2 * 128) + 15) & ~(uintptr_t)15);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mtile = blockIdx.x, y = blockIdx.y;
  const int t0 = mtile * 128;
  constexpr uint32_t tmem_cols = 512;
  const int n_steps = fp.n_steps;

  {  // step table -> shared memory
    const uint32_t* src = reinterpret_cast<const uint32_t*>(fp.steps);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_steps);
    const int n_words = (int)(sizeof(EncStep) / 4) * n_steps;
    for (int i = tid; i < n_words; i += EC_THREADS) dst[i] = src[i];
  }
  if (tid == 0) {
    for (int i = 0; i < EC_AST; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < EC_WST; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, EW);
    mbar_init(at_qk, 1);
    mbar_init(at_s, 1);
    mbar_init(at_v, 1);
    mbar_init(at_p, EW * 32);
    mbar_init(at_o, 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem + 256, tO = tmem + 256 + EA_XR + 16;  // attention: S' 144 columns (reads touch 160), O 96

  // does this CTA run the attention of step st?
  auto att_mine = [&](const EncStep& st) { return mtile == 0 && y >= st.att.y0 && y < st.att.y0 + st.att.n_heads; };

  if (warp == W_LOAD) {
    // ===================== weight producer: runs ahead of the step barriers (weights are static) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1, att_use = 0;
      for (int b = blockIdx.z; b < fp.batch; b += gridDim.z) {
        if (t0 >= fp.lens[b]) continue;
        for (int sidx = 0; sidx < n_steps; ++sidx) {
          const EncStep& st = s_steps[sidx];
          if (st.type == ES_ATTN) {
            // the attention operands overlay the rings: no weight may land before its last MMA has read them
            if (att_mine(st)) {
              mbar_wait(at_o, att_use & 1);
              ++att_use;
            }
            continue;
          }
          if (st.type != ES_CONV) continue;
          const EncConvOp* op = ec_my_op(st, y);
          if (!op) continue;
          const int nt = (y - op->y0) % op->ntiles, ks = (y - op->y0) / op->ntiles;
          const int nch = (op->Cin + UM_BKC - 1) / UM_BKC, per = nch / op->ksplit;
          const uint32_t w_stage = (uint32_t)NSPLIT * UM_KCH * op->NT * 16;
          const unsigned char* gw = reinterpret_cast<const unsigned char*>(op->w) + ((size_t)nt * nch + (size_t)ks * per) * op->K * w_stage;
          const int n_blocks = per * op->K;  // (chunk, tap) blocks are consecutive in the image
          for (int i = 0; i < n_blocks; i += EC_G) {
            const uint32_t bytes = (uint32_t)min(EC_G, n_blocks - i) * w_stage;
            mbar_wait(&w_empty[s], ph);
            mbar_arrive_expect_tx(&w_full[s], bytes);
            bulk_g2s(sW + (size_t)s * w_slot, gw + (size_t)i * w_stage, bytes, &w_full[s]);
            if (++s == EC_WST) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == W_ALOAD) {
    // ===================== activation producer: waits for the previous step of THIS utterance, then bulk-copies ========
    // The whole warp walks the step table (uniform control flow).  Conv steps: lane 0 issues the few large copies of a stage.
    // Attention steps: 108 + 72 small copies (one per 8-channel group of a plane) — every lane issues its share (a single
    // lane needed 13 us for Q and K' alone, profiles/r2a_chain_trace.log).
    {
      int as = 0;
      uint32_t aph = 1, att_use = 0;
      for (int b = blockIdx.z; b < fp.batch; b += gridDim.z) {
        const int Lb = fp.lens[b];
        if (t0 >= Lb) continue;
        const unsigned int m_tiles = (unsigned int)((Lb + 127) / 128);
        const unsigned int* cnt = fp.counter + (size_t)b * n_steps;
        for (int sidx = 0; sidx < n_steps; ++sidx) {
          const EncStep& st = s_steps[sidx];
          if (st.type == ES_CONV) {
            const EncConvOp* op = ec_my_op(st, y);
            if (!op) continue;
            const int ks = (y - op->y0) / op->ntiles;
            const int nch = (op->Cin + UM_BKC - 1) / UM_BKC, per = nch / op->ksplit;
            if (lane == 0 && sidx > 0) ec_wait_step(cnt, sidx - 1, ec_expect(s_steps[sidx - 1], m_tiles), b);
            __syncwarp();
            asm volatile("fence.proxy.async;" ::: "memory");  // (each issuing lane) other CTAs' generic-proxy writes -> our async-proxy reads
            if (lane == 0) EC_TRACE(sidx, 0);
            const unsigned char* xp = op->xp + (size_t)b * op->xp_bs;
            // copied from the 8-row (128-byte) boundary below the first row: unaligned bulk copies run ~3x slower
            const long long row0u = (long long)t0 - op->pad + UM_PADL;
            const uint32_t roff = a_row_off(row0u);
            const long long row0 = row0u - roff;
            const uint32_t kc_bytes = (uint32_t)((128 + (op->K - 1) + (int)roff + 7) & ~7) * 16;
            // a chunk = 3 planes x 4 channel groups = 12 copies of ~2 KB; one lane needs ~100 ns per copy (6 chunks = 7 us in
            // the ffn1 / qkv steps of profiles/r2a_chain_trace.log): lanes 0-11 and 12-23 issue two chunks at once
            static_assert(EC_AST >= 2 && NSPLIT * UM_KCH == 12, "lane -> (stage, copy) mapping");
            const int q = lane / 12, j = lane - q * 12;
            for (int c0 = 0; c0 < per; c0 += 2) {
              const int c = ks * per + c0 + q;
              int sq = as + q;
              uint32_t pq = aph;
              if (sq >= EC_AST) { sq -= EC_AST; pq ^= 1; }
              if (q < 2 && c0 + q < per) {
                mbar_wait(&a_empty[sq], pq);
                if (j == 0) mbar_arrive_expect_tx(&a_full[sq], op->planes * UM_KCH * kc_bytes);
                const int sp = j >> 2, kc = j & 3;
                if (sp < op->planes)  // bf16-activation mode reads the hi plane only
                  bulk_g2s(sA + (size_t)sq * a_stage + sp * a_split_bytes + kc * RA * 16,
                         xp + (((size_t)sp * op->xp_kc + c * UM_KCH + kc) * fp.rows + row0) * 16, kc_bytes, &a_full[sq]);
              }
              as += min(2, per - c0);
              if (as >= EC_AST) { as -= EC_AST; aph ^= 1; }
              __syncwarp();
            }
          } else if (st.type == ES_ATTN && att_mine(st)) {
            const EncAttnOp& at = st.att;
            if (lane == 0 && sidx > 0) ec_wait_step(cnt, sidx - 1, ec_expect(s_steps[sidx - 1], m_tiles), b);
            __syncwarp();
            asm volatile("fence.proxy.async;" ::: "memory");  // (each issuing lane) other CTAs' writes -> our async-proxy reads
            if (lane == 0) EC_TRACE(sidx, 0);
            const int h = y - at.y0;
            const unsigned char* src = at.qkv_p + (size_t)b * at.qkv_bs;
            const int hk = at.H / 8;  // channel groups per q / k / v block
            auto img = [&](int opi, int sp, int kc) {
              const size_t g = (size_t)opi * hk + (size_t)h * EA_KC + kc;
              return src + (((size_t)sp * at.qkv_kc + g) * fp.rows + UM_PADL) * 16;
            };
            unsigned char* sQ = smem_raw;
            unsigned char* sK = smem_raw + EA_NS * EA_QPLANE;
            unsigned char* sV = smem_raw;
            // phase 1: Q rows [0,128) and K' = [K rows [0,128); Ek rows] (image rows >= len are zero)
            if (lane == 0) mbar_arrive_expect_tx(at_qk, at.planes * (EA_QPLANE + EA_XPLANE));
            __syncwarp();
            for (int i = lane; i < at.planes * EA_KC; i += 32) {
              const int sp = i / EA_KC, kc = i - sp * EA_KC;
              bulk_g2s(sQ + sp * EA_QPLANE + kc * EA_ROWS * 16, img(0, sp, kc), EA_ROWS * 16, at_qk);
              bulk_g2s(sK + sp * EA_XPLANE + kc * EA_XR * 16, img(1, sp, kc), EA_ROWS * 16, at_qk);
              bulk_g2s(sK + sp * EA_XPLANE + kc * EA_XR * 16 + EA_ROWS * 16, at.relk_img + ((size_t)sp * EA_KC + kc) * 256, 256, at_qk);
            }
            // phase 2: V' = [V; Ev] over the dead Q / K' bytes, as soon as the S MMAs have retired
            mbar_wait(at_s, att_use & 1);
            if (lane == 0) mbar_arrive_expect_tx(at_v, at.planes * EA_XPLANE);
            __syncwarp();
            for (int i = lane; i < at.planes * EA_KC; i += 32) {
              const int sp = i / EA_KC, kc = i - sp * EA_KC;
              bulk_g2s(sV + sp * EA_XPLANE + kc * EA_XR * 16, img(2, sp, kc), EA_ROWS * 16, at_v);
              bulk_g2s(sV + sp * EA_XPLANE + kc * EA_XR * 16 + EA_ROWS * 16, at.relv_img + ((size_t)sp * EA_KC + kc) * 256, 256, at_v);
            }
            ++att_use;
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer =====================
    int ws = 0, as = 0;
    uint32_t wph = 0, aph = 0, use = 0, att_use = 0;
    const uint64_t a_desc0 = make_smem_desc(smem_u32(sA), RA * 16, 128);
    const uint32_t a_ks = (2 * RA * 16) >> 4, a_sp = a_split_bytes >> 4;
    for (int b = blockIdx.z; b < fp.batch; b += gridDim.z) {
      if (t0 >= fp.lens[b]) continue;
      for (int sidx = 0; sidx < n_steps; ++sidx) {
        const EncStep& st = s_steps[sidx];
        if (st.type == ES_CONV) {
          const EncConvOp* op = ec_my_op(st, y);
          if (!op) continue;  // uniform across the warp
          const int NT = op->NT, K = op->K;
          const int nch = (op->Cin + UM_BKC - 1) / UM_BKC, per = nch / op->ksplit;
          const uint32_t idesc1 = make_idesc_bf16_f32(128, NT), idesc2 = make_idesc_bf16_f32(128, 2 * NT);
          const uint32_t idesc3 = make_idesc_bf16_f32(128, 3 * NT);
          const uint64_t b_desc0 = make_smem_desc(smem_u32(sW), NSPLIT * NT * 16, 128);
          const uint32_t b_ks = (2 * NSPLIT * NT * 16) >> 4;
          const uint32_t roff = a_row_off((long long)t0 - op->pad + UM_PADL);
          mbar_wait(acc_empty, (use & 1) ^ 1);  // the epilogue of this CTA's previous conv step has drained the accumulator
          tc_fence_after();
          const uint32_t w_stage = (uint32_t)NSPLIT * UM_KCH * NT * 16;
          const int n_blocks = per * K;
          for (int i0 = 0; i0 < n_blocks; i0 += EC_G) {
            mbar_wait(&w_full[ws], wph);
            tc_fence_after();
            if (i0 == 0 && lane == 0) EC_TRACE(sidx, 2);
            const int cnt = min(EC_G, n_blocks - i0);
            for (int j = 0; j < cnt; ++j) {
              const int i = i0 + j, c = i / K, k = i - c * K;
              if (k == 0) {  // first tap of a chunk: its activation stage must have landed
                mbar_wait(&a_full[as], aph);
                tc_fence_after();
                if (i == 0 && lane == 0) EC_TRACE(sidx, 1);
              }
              if (elect_one()) {
                const uint64_t a_k = a_desc0 + ((as * a_stage) >> 4) + roff + (uint32_t)k;  // conv tap (dilation 1): shift by k rows
                const uint64_t b_k = b_desc0 + ((ws * w_slot + j * w_stage) >> 4);
#pragma unroll
                for (int ks2 = 0; ks2 < UM_BKC / 16; ++ks2) {
                  const uint64_t a0 = a_k + ks2 * a_ks, b0 = b_k + ks2 * b_ks;
                  const uint32_t acc0 = (i == 0 && ks2 == 0) ? 0u : 1u;
                  mma_bf16_ss(tmem, a0, b0, idesc3, acc0);                     // hi  x [hi|lo|lo2] -> [c0|c1|c2]
                  if (op->planes == 3) {
                    mma_bf16_ss(tmem + NT, a0 + a_sp, b0, idesc2, 1u);         // lo  x [hi|lo]     -> [c1|c2]
                    mma_bf16_ss(tmem + 2 * NT, a0 + 2 * a_sp, b0, idesc1, 1u); // lo2 x hi          -> c2
                  }
                }
                if (k == K - 1) mma_commit(&a_empty[as]);
                if (j == cnt - 1) mma_commit(&w_empty[ws]);
              }
              __syncwarp();
              if (k == K - 1 && ++as == EC_AST) { as = 0; aph ^= 1; }
            }
            if (++ws == EC_WST) { ws = 0; wph ^= 1; }
          }
          if (elect_one()) mma_commit(acc_full);
          __syncwarp();
          if (lane == 0) EC_TRACE(sidx, 3);
          ++use;
        } else if (st.type == ES_ATTN && att_mine(st)) {
          unsigned char* sQ = smem_raw;
          unsigned char* sK = smem_raw + EA_NS * EA_QPLANE;
          unsigned char* sV = smem_raw;
          unsigned char* sP = smem_raw + EA_NS * EA_XPLANE;
          const uint32_t idS = make_idesc_bf16_f32(128, EA_XR);
          const uint32_t idO = make_idesc_bf16_f32(128, EA_DK) | (1u << 16);  // b_major = MN
          mbar_wait(at_qk, att_use & 1);
          tc_fence_after();
          if (lane == 0) EC_TRACE(sidx, 1);
          if (elect_one()) {
            const uint64_t qd = make_smem_desc(smem_u32(sQ), EA_ROWS * 16, 128);
            const uint64_t kd = make_smem_desc(smem_u32(sK), EA_XR * 16, 128);
            const uint32_t qks = (2 * EA_ROWS * 16) >> 4, kks = (2 * EA_XR * 16) >> 4, qsp = EA_QPLANE >> 4, ksp = EA_XPLANE >> 4;
#pragma unroll
            for (int k = 0; k < EA_DK / 16; ++k) {
              const uint64_t a = qd + k * qks, bb = kd + k * kks;
              const uint32_t acc = k ? 1u : 0u;
              if (st.att.planes == 3) {
                mma_bf16_ss(tS, a + qsp, bb + ksp, idS, acc);     // lo*lo
                mma_bf16_ss(tS, a, bb + 2 * ksp, idS, 1u);        // hi*lo2
                mma_bf16_ss(tS, a + 2 * qsp, bb, idS, 1u);        // lo2*hi
                mma_bf16_ss(tS, a, bb + ksp, idS, 1u);            // hi*lo
                mma_bf16_ss(tS, a + qsp, bb, idS, 1u);            // lo*hi
              }
              mma_bf16_ss(tS, a, bb, idS, st.att.planes == 3 ? 1u : acc);  // hi*hi
            }
            mma_commit(at_s);
          }
          __syncwarp();
          mbar_wait(at_v, att_use & 1);
          mbar_wait(at_p, att_use & 1);
          tc_fence_after();
          if (elect_one()) {
            // A = P' image [18 key chunks][128 rows][8], K-major;  B = V' [12 channel groups][144 keys][8], MN-major:
            // LBO = 128 B (next 8 keys), SBO = 144*16 B (next 8 channels)
            const uint64_t pd = make_smem_desc(smem_u32(sP), EA_ROWS * 16, 128);
            const uint64_t vd = make_smem_desc(smem_u32(sV), 128, EA_XR * 16);
            const uint32_t pks = (2 * EA_ROWS * 16) >> 4, psp = EA_PPLANE >> 4;
            const uint32_t vks = (16 * 16) >> 4, vsp = EA_XPLANE >> 4;
#pragma unroll
            for (int k = 0; k < EA_XR / 16; ++k) {
              const uint64_t a = pd + k * pks, bb = vd + k * vks;
              const uint32_t acc = k ? 1u : 0u;
              if (st.att.planes == 3) {
                mma_bf16_ss(tO, a + psp, bb + vsp, idO, acc);
                mma_bf16_ss(tO, a, bb + 2 * vsp, idO, 1u);
                mma_bf16_ss(tO, a + 2 * psp, bb, idO, 1u);
                mma_bf16_ss(tO, a, bb + vsp, idO, 1u);
                mma_bf16_ss(tO, a + psp, bb, idO, 1u);
              }
              mma_bf16_ss(tO, a, bb, idO, st.att.planes == 3 ? 1u : acc);
            }
            mma_commit(at_o);
          }
          __syncwarp();
          if (lane == 0) EC_TRACE(sidx, 3);
          ++att_use;
        }
      }
    }
  } else {
    // ===================== epilogue warps (conv epilogue | reduce + LayerNorm | softmax) =====================
    uint32_t use = 0, att_use = 0;
    const int it = 1;  // (tracing in the shared epilogue text is tied to it == 0: off)
    for (int b = blockIdx.z; b < fp.batch; b += gridDim.z) {
      const int Lb = fp.lens[b];
      if (t0 >= Lb) continue;
      const unsigned int m_tiles = (unsigned int)((Lb + 127) / 128);
      unsigned int* cnt = fp.counter + (size_t)b * n_steps;
      for (int sidx = 0; sidx < n_steps; ++sidx) {
        const EncStep& st = s_steps[sidx];
        bool took_part = false;
        if (st.type == ES_CONV) {
          const EncConvOp* op = ec_my_op(st, y);
          if (!op) continue;
          took_part = true;
          const int NT = op->NT;
          const int nt = (y - op->y0) % op->ntiles, ks = (y - op->y0) / op->ntiles;
          const int n0 = nt * NT;
          for (int i = tid; i < NT; i += EW * 32) s_bias[i] = (op->bias && (n0 + i) < op->Cout) ? op->bias[n0 + i] : 0.f;
          ec_epi_bar();
          mbar_wait(acc_full, use & 1);
          tc_fence_after();
          if (tid == 0) EC_TRACE(sidx, 4);
          {
            const uint32_t tm = tmem;
            EncEpiP p;
            p.yp = op->yp; p.yp_bs = op->yp_bs; p.yp_kc = op->yp_kc; p.yp_rows = fp.rows;
            p.y = op->y ? op->y + (size_t)ks * op->part_stride : nullptr; p.y_bs = op->y_bs; p.y_ld = op->y_ld;
            p.act = op->act; p.Cout = op->Cout;
#include "conv_umma_epilogue.inc"
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
          ++use;
        } else if (st.type == ES_REDUCE) {
          // all EC_YT CTA rows of the time tile take 128 / EC_YT = 8 rows each: one 8-channel group per thread, so every
          // thread has ONE batch of independent L2 loads in flight (4 CTAs x 4 groups per thread took 8-12 us per step)
          took_part = true;
          const EncReduceOp& r = st.red;
          if (tid == 0 && sidx > 0) ec_wait_step(cnt, sidx - 1, ec_expect(s_steps[sidx - 1], m_tiles), b);
          ec_epi_bar();
          if (tid == 0) EC_TRACE(sidx, 4);
          constexpr int RPC = 128 / EC_YT;  // rows per CTA
          static_assert(RPC == 8 && EW * 32 == 256, "reduce mapping: 8 rows x 32 channel groups per CTA");
          const int tx = tid & 7, g = tid >> 3;  // row, 8-channel group (a warp = 4 groups x 8 rows)
          const int t = t0 + y * RPC + tx;
          const bool valid = t < Lb;
          const int ngrp = r.C >> 3;
          const bool ok = valid && g < ngrp;
          const float* xb = r.x + (size_t)b * r.x_bs;
          const float* rb = r.res ? r.res + (size_t)b * r.res_bs : nullptr;
          float v[8];
          float sum = 0.f;
          {
            // plain L2 loads (ld.global.cg: the partials and the residual were written by other CTAs in earlier steps), all
            // independent and issued back to back
            float part[4][8], rs[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const size_t off = ok ? (size_t)(g * 8 + e) * r.ld + t : 0;
#pragma unroll
              for (int pp = 0; pp < 4; ++pp) part[pp][e] = pp < r.parts ? __ldcg(xb + (size_t)pp * r.part_stride + off) : 0.f;
              rs[e] = rb ? __ldcg(rb + off) : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c = g * 8 + e;
              float a = part[0][e];
#pragma unroll
              for (int pp = 1; pp < 4; ++pp)
                if (pp < r.parts) a += part[pp][e];
              if (r.bias) a += r.bias[ok ? c : 0];
              if (rb) a += rs[e];
              if (r.pre_relu) a = fmaxf(a, 0.f);
              a = ok ? a : 0.f;
              v[e] = a;
              sum += a;
            }
          }
          float mean = 0.f, rstd = 1.f;
          if (r.gamma) {
            // row statistics: 4 groups of a row inside the warp (lanes tx, tx+8, tx+16, tx+24), then the 8 warps through smem
            sum += __shfl_xor_sync(0xffffffffu, sum, 8);
            sum += __shfl_xor_sync(0xffffffffu, sum, 16);
            if (lane < 8) s_red[warp * 8 + lane] = sum;
            ec_epi_bar();
            float tot = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) tot += s_red[j * 8 + tx];
            mean = tot / (float)r.C;
            float q = 0.f;
            if (g < ngrp) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float d = v[e] - mean;
                q += d * d;
              }
            }
            q += __shfl_xor_sync(0xffffffffu, q, 8);
            q += __shfl_xor_sync(0xffffffffu, q, 16);
            ec_epi_bar();  // every thread has read the sums
            if (lane < 8) s_red[warp * 8 + lane] = q;
            ec_epi_bar();
            float qt = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) qt += s_red[j * 8 + tx];
            rstd = rsqrtf(qt / (float)r.C + 1e-4f);
          }
          float* yb = r.y ? r.y + (size_t)b * r.y_bs : nullptr;
          unsigned char* pb = r.yp ? r.yp + (size_t)b * r.yp_bs : nullptr;
          const size_t pstride = (size_t)r.yp_kc * fp.rows * 16;
          const uint4 z4 = make_uint4(0, 0, 0, 0);
          if (g < ngrp) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c = g * 8 + e;
              float w = v[e];
              if (r.gamma) w = (w - mean) * rstd * r.gamma[c] + r.beta[c];
              if (r.post_relu) w = fmaxf(w, 0.f);
              o[e] = valid ? w : 0.f;
              if (valid && yb) yb[(size_t)c * r.ld + t] = w;
            }
            if (pb) {
              uint4 q3[3];
              split8<3>(o, q3);
              unsigned char* base = pb + ((size_t)g * fp.rows + UM_PADL + t) * 16;
              *reinterpret_cast<uint4*>(base) = q3[0];
              *reinterpret_cast<uint4*>(base + pstride) = q3[1];
              *reinterpret_cast<uint4*>(base + 2 * pstride) = q3[2];
              // zero rows: t in [-PADL, 0) by the first row block; the 128 rows behind the last tile, 8 per row block
              if (mtile == 0 && y == 0) {
                for (int rr = tx; rr < UM_PADL; rr += RPC) {
                  unsigned char* zb = pb + ((size_t)g * fp.rows + rr) * 16;
                  *reinterpret_cast<uint4*>(zb) = z4;
                  *reinterpret_cast<uint4*>(zb + pstride) = z4;
                  *reinterpret_cast<uint4*>(zb + 2 * pstride) = z4;
                }
              }
              if (t0 + 128 >= Lb) {
                unsigned char* zb = pb + ((size_t)g * fp.rows + UM_PADL + t0 + 128 + y * RPC + tx) * 16;
                *reinterpret_cast<uint4*>(zb) = z4;
                *reinterpret_cast<uint4*>(zb + pstride) = z4;
                *reinterpret_cast<uint4*>(zb + 2 * pstride) = z4;
              }
            }
          }
        } else {  // ES_ATTN
          if (!att_mine(st)) continue;
          took_part = true;
          const EncAttnOp& at = st.att;
          const int h = y - at.y0;
          const int nrel = 2 * at.window + 1;
          unsigned char* sP = smem_raw + EA_NS * EA_XPLANE;
          // ---- softmax + epilogue on all 8 warps: thread = (query row = TMEM lane, column half).  Warps w and w + 4 read the
          // same 32 lanes; `half` picks the 64 key columns (softmax) / 48 output channels (epilogue) a thread owns.  One
          // thread per row (the first version) spent ~6k dependent instructions on a single warp per scheduler: 56 us.
          const int i = (warp & 3) * 32 + lane;  // 0..127
          const int half = warp >> 2;
          const bool row_ok = i < Lb;
          const float inv_sqrt = 1.f / sqrtf((float)EA_DK);
          float* my_relk = s_relk + (half * 128 + i) * 9;
          float* my_band = s_band + i * 9;
          mbar_wait(at_s, att_use & 1);
          tc_fence_after();
          if (tid == 0) EC_TRACE(sidx, 4);
          const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32) << 16);
          {
            uint32_t raw[32];
            tmem_ld32(tS + lane_addr + EA_ROWS, raw);  // columns 128..159: q_i . Ek[r] in the first 2w+1
            tmem_ld_wait();
#pragma unroll
            for (int rr = 0; rr < 9; ++rr) my_relk[rr] = __uint_as_float(raw[rr]);
          }
          constexpr int HC = EA_ROWS / 2;  // key columns per thread
          float srow[HC];
          float mx = -INFINITY;
#pragma unroll
          for (int g = 0; g < HC / 32; ++g) {
            uint32_t raw[32];
            tmem_ld32(tS + lane_addr + half * HC + g * 32, raw);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = half * HC + g * 32 + j;
              float sv = __uint_as_float(raw[j]);
              const int rr = col - i + at.window;
              const bool in_band = (unsigned)rr < (unsigned)nrel;
              if (in_band) sv += my_relk[rr];
              sv *= inv_sqrt;
              sv = col < Lb ? sv : -INFINITY;
              if (in_band) my_band[rr] = sv;  // the band probabilities are rebuilt from these (no dynamic register index)
              srow[g * 32 + j] = sv;
              mx = fmaxf(mx, sv);
            }
          }
          s_mx[half * 128 + i] = mx;
          ec_epi_bar();
          mx = fmaxf(s_mx[i], s_mx[128 + i]);
          float ssum = 0.f;
#pragma unroll
          for (int c2 = 0; c2 < HC; ++c2) {
            const int col = half * HC + c2;
            const float e = col < Lb ? __expf(srow[c2] - mx) : 0.f;
            srow[c2] = e;
            ssum += e;
          }
          s_sum[half * 128 + i] = ssum;
          ec_epi_bar();
          ssum = s_sum[i] + s_sum[128 + i];
          const float inv = row_ok ? 1.f / ssum : 0.f;
          // P' image over the dead K' bytes (the S MMAs have retired: at_s)
#pragma unroll
          for (int q8 = 0; q8 < HC / 8; ++q8) {
            const int kc = half * (HC / 8) + q8;
            float v8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v8[e] = srow[q8 * 8 + e] * inv;
            uint4 q3[3];
            split8<3>(v8, q3);
            unsigned char* o = sP + ((size_t)kc * EA_ROWS + i) * 16;
            *reinterpret_cast<uint4*>(o) = q3[0];
            *reinterpret_cast<uint4*>(o + EA_PPLANE) = q3[1];
            *reinterpret_cast<uint4*>(o + 2 * EA_PPLANE) = q3[2];
          }
          // band probabilities p[i][i + r - w] as 16 extra "keys" (they meet the Ev rows of V'): half 0 writes r = 0..7,
          // half 1 writes r = 8 and the zero padding
          {
            float band[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int rr = half * 8 + e;
              const int col = i + rr - at.window;
              float pv = 0.f;
              if (rr < nrel && col >= 0 && col < Lb) pv = __expf(my_band[rr < 9 ? rr : 0] - mx) * inv;
              band[e] = pv;
            }
            uint4 q3[3];
            split8<3>(band, q3);
            unsigned char* o = sP + ((size_t)(EA_ROWS / 8 + half) * EA_ROWS + i) * 16;
            *reinterpret_cast<uint4*>(o) = q3[0];
            *reinterpret_cast<uint4*>(o + EA_PPLANE) = q3[1];
            *reinterpret_cast<uint4*>(o + 2 * EA_PPLANE) = q3[2];
          }
          fence_proxy_async_smem();
          mbar_arrive(at_p);

          mbar_wait(at_o, att_use & 1);
          tc_fence_after();
          unsigned char* opb = at.out_p + (size_t)b * at.out_bs;
          const size_t ostride = (size_t)at.out_kc * fp.rows * 16;
          // output channels [48*half, 48*half + 48) of this head: three 16-column pieces
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            const int c0 = half * (EA_DK / 2) + g * 16;
            uint32_t raw[16];
            tmem_ld16(tO + lane_addr + c0, raw);
            tmem_ld_wait();
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = row_ok ? __uint_as_float(raw[j]) : 0.f;
#pragma unroll
            for (int q4 = 0; q4 < 2; ++q4) {
              uint4 q3[3];
              split8<3>(o + q4 * 8, q3);
              const int kc = h * EA_KC + c0 / 8 + q4;
              unsigned char* dst = opb + ((size_t)kc * fp.rows + UM_PADL + i) * 16;
              *reinterpret_cast<uint4*>(dst) = q3[0];
              *reinterpret_cast<uint4*>(dst + ostride) = q3[1];
              *reinterpret_cast<uint4*>(dst + 2 * ostride) = q3[2];
              // zero rows around the data: t in [-PADL, 0) by rows i < PADL, t in [128, 256) by all rows
              const uint4 z = make_uint4(0, 0, 0, 0);
              if (i < UM_PADL) {
                unsigned char* zf = opb + ((size_t)kc * fp.rows + i) * 16;
                *reinterpret_cast<uint4*>(zf) = z;
                *reinterpret_cast<uint4*>(zf + ostride) = z;
                *reinterpret_cast<uint4*>(zf + 2 * ostride) = z;
              }
              unsigned char* zt = opb + ((size_t)kc * fp.rows + UM_PADL + EA_ROWS + i) * 16;
              *reinterpret_cast<uint4*>(zt) = z;
              *reinterpret_cast<uint4*>(zt + ostride) = z;
              *reinterpret_cast<uint4*>(zt + 2 * ostride) = z;
            }
          }
          tc_fence_before();
          ++att_use;
        }
        if (!took_part) continue;
        if (tid == 0) EC_TRACE(sidx, 5);
        // publish this CTA's part of the step: the named barrier orders every epilogue thread's global writes before
        // thread 0, whose gpu-scope release is cumulative over them
        ec_epi_bar();
        if (tid == 0) {
          asm volatile("fence.proxy.async;" ::: "memory");
          ec_red_release(cnt + sidx);
          EC_TRACE(sidx, 6);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace b2
