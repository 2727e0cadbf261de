// conv_umma.cuh — Conv1d as an implicit GEMM on the 5th-gen tensor cores (tcgen05, accumulators in TMEM)
// with fp32-class accuracy from a bf16 split:  x = hi + lo (+ lo2),  D = Ahi*Bhi + (Ahi*Blo + Alo*Bhi) (+ ...)
//
//   D[t][co] = sum_k sum_ci  X[ci][t + k*dil - pad_left] * W[co][ci][k]
//   M = time (128 rows per MMA, MT tiles per CTA), N = output channels (NT <= 128 per CTA), K = input channels
//
// Operand images.  Both MMA operands are K-major, SWIZZLE_NONE with SBO = 128 B, i.e. [k-chunk of 8][row][8 bf16]:
// the address is linear in the row, so conv tap k is the SAME activation image with the descriptor start address
// advanced by k*dil rows (16 B each) — no im2col, no per-tap copies.
//   * weights: pre-split bf16 image per (n-tile, ci-chunk, tap) with the hi/lo(/lo2) planes of a k-chunk as adjacent
//     row blocks: one cp.async.bulk (TMA engine) per ring slot, and ONE MMA with N = 2*NT (3*NT) multiplies an
//     activation plane with all weight planes at once (an SS-mode MMA re-reads its A block from shared memory, so
//     fewer, wider MMAs are what narrow layers need);
//   * activations, "plane" mode: the PRODUCING layer's epilogue already wrote the bf16 hi/lo planes of its
//     (pre-activated) output in image order [plane][C/8][PADL + rows][8] to global memory, so the consumer's A
//     tile is 2*4 contiguous bulk copies per 32-channel chunk (from the 128-byte boundary below its first row) — no SM
//     instructions on the operand path;
//   * activations, "gather" mode (inputs produced by non-conv kernels): the epilogue warps read fp32 [C][L], apply the
//     pre-op, split and write the image to smem themselves.
//
// One CTA = EW*32 + 64 (+32) threads:
//   warps 0..EW-1 : gather-mode A producers during the main loop; otherwise the epilogue:
//                   TMEM -> registers (tcgen05.ld 32x32b) -> bias / act / residual / tanh*sigmoid gate / WN res-skip /
//                   affine coupling + InvConvNear^-1 + ActNorm^-1 / pixel shuffle -> fp32 [C][L] and/or the bf16 planes
//                   of the next conv (including the zero rows that implement the reference's padding/masking);
//   warp EW       : MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::f16, one accumulator per magnitude class;
//   warp EW+1     : lane 0: weight bulk copies;
//   warp EW+2     : (NBUF = 2 only) activation bulk copies, one copy per lane (issue_a_chunks); with NBUF = 1 warp 0 issues them
//                   for its one tile before it joins the epilogue.
// Pipelines: A ring (AST stages, per ci-chunk of 32 channels, reused by all K taps), W ring (WST slots of G taps),
// accumulator full/empty barriers per TMEM buffer (NBUF = 2: the MMA warp fills one buffer while the epilogue warps drain
// the other and the rings run ahead into the CTA's next time tile).  All waits are mbarrier based; tcgen05.commit
// releases stages.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace b2 {

constexpr int UM_BKC = 32;               // input channels per A stage
constexpr int UM_KCH = UM_BKC / 8;       // 16-byte k-chunks per stage
constexpr int UM_HALO = 80;              // max (K-1)*dil
constexpr int UM_WSTAGES_MAX = 16;
constexpr int UM_PADL = 40;              // zero rows in front of a plane image (>= any pad_left)
constexpr int UM_TAIL = 128;             // zero rows written behind the last tile (>= UM_HALO)

// rows (per 8-channel group) a plane image must provide for a tensor of max length L
__host__ __device__ inline int um_plane_rows(int L) { return UM_PADL + (L + 511) / 512 * 512 + UM_TAIL + 128; }

template <int MT>
struct UmmaCfg {
  static constexpr int RA = MT * 128 + UM_HALO;  // rows of the A image
  static size_t a_stage_bytes(int nsplit) { return (size_t)nsplit * UM_KCH * RA * 16; }
  static size_t w_stage_bytes(int nsplit, int NT) { return (size_t)nsplit * UM_KCH * NT * 16; }
  static size_t smem_bytes(int nsplit, int NT, int wstages, int astages, int taps_per_stage) {
    return astages * a_stage_bytes(nsplit) + (size_t)wstages * taps_per_stage * w_stage_bytes(nsplit, NT) + 384 +
           4 * (256 + 16 + 2 * 256);
  }
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define UM_TRACE(slot)                                                                           \
  do {                                                                                           \
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.dbg[slot] = gtimer(); \
  } while (0)

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// split 8 floats into NSPLIT bf16 planes, 16 bytes each
// two floats -> packed bf16x2 (one cvt.rn.bf16x2.f32), low half = a
__device__ __forceinline__ uint32_t cvt_bf16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
template <int NSPLIT>
__device__ __forceinline__ void split8(const float* v, uint4* out) {
  uint32_t h[4], l[4], l2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float a = v[2 * e], b = v[2 * e + 1];
    h[e] = cvt_bf16x2(a, b);
    const float ra = a - __uint_as_float(h[e] << 16), rb = b - __uint_as_float(h[e] & 0xffff0000u);
    l[e] = cvt_bf16x2(ra, rb);
    if (NSPLIT == 3) l2[e] = cvt_bf16x2(ra - __uint_as_float(l[e] << 16), rb - __uint_as_float(l[e] & 0xffff0000u));
  }
  out[0] = make_uint4(h[0], h[1], h[2], h[3]);
  out[1] = make_uint4(l[0], l[1], l[2], l[3]);
  if (NSPLIT == 3) out[2] = make_uint4(l2[0], l2[1], l2[2], l2[3]);
}

// A plane image in global memory (written by epilogues, read by bulk copies)
struct PlaneOut {
  unsigned char* base;  // item base
  int kc_total, rows, nsplit;
  float slope;          // value stored = v > 0 ? v : v*slope  (1 = identity)
  // zero-row bookkeeping of the calling thread
  bool pad_front;       // this CTA owns t = [-PADL, 0)
  bool pad_tail;        // this CTA owns the UM_TAIL rows behind its tile
  int m;                // thread's row index inside the CTA tile (0 .. MT*128-1)
  long long tail_row0;  // first tail row (image row index, includes PADL)
  __device__ __forceinline__ unsigned char* at(int plane, int kc, long long row) const {
    return base + (((size_t)plane * kc_total + kc) * rows + row) * 16;
  }
};

// store 8 consecutive channels (one 16-byte chunk per plane) of image row `row` (+ this thread's share of zero rows)
__device__ __forceinline__ void plane_store8(const PlaneOut& po, int kc, long long row, const float* v8) {
  float w[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) w[e] = v8[e] > 0.f ? v8[e] : v8[e] * po.slope;
  const size_t pstride = (size_t)po.kc_total * po.rows * 16;  // bytes between split planes
  unsigned char* base = po.base + ((size_t)kc * po.rows) * 16;
  const uint4 z = make_uint4(0, 0, 0, 0);
  const bool zf = po.pad_front && po.m < UM_PADL, zt = po.pad_tail && po.m < UM_TAIL;
  uint4 q[3];
  if (po.nsplit == 3) {  // straight-line code per split count: no dynamically indexed local arrays
    split8<3>(w, q);
    *reinterpret_cast<uint4*>(base + row * 16) = q[0];
    *reinterpret_cast<uint4*>(base + pstride + row * 16) = q[1];
    *reinterpret_cast<uint4*>(base + 2 * pstride + row * 16) = q[2];
    if (zf) {
      *reinterpret_cast<uint4*>(base + (size_t)po.m * 16) = z;
      *reinterpret_cast<uint4*>(base + pstride + (size_t)po.m * 16) = z;
      *reinterpret_cast<uint4*>(base + 2 * pstride + (size_t)po.m * 16) = z;
    }
    if (zt) {
      *reinterpret_cast<uint4*>(base + (po.tail_row0 + po.m) * 16) = z;
      *reinterpret_cast<uint4*>(base + pstride + (po.tail_row0 + po.m) * 16) = z;
      *reinterpret_cast<uint4*>(base + 2 * pstride + (po.tail_row0 + po.m) * 16) = z;
    }
  } else {
    split8<2>(w, q);
    *reinterpret_cast<uint4*>(base + row * 16) = q[0];
    *reinterpret_cast<uint4*>(base + pstride + row * 16) = q[1];
    if (zf) {
      *reinterpret_cast<uint4*>(base + (size_t)po.m * 16) = z;
      *reinterpret_cast<uint4*>(base + pstride + (size_t)po.m * 16) = z;
    }
    if (zt) {
      *reinterpret_cast<uint4*>(base + (po.tail_row0 + po.m) * 16) = z;
      *reinterpret_cast<uint4*>(base + pstride + (po.tail_row0 + po.m) * 16) = z;
    }
  }
}
// store 4 consecutive channels (8 bytes at byte offset `boff` inside the 16-byte chunk); no zero-row duty
__device__ __forceinline__ void plane_store4(const PlaneOut& po, int kc, int boff, long long row, const float* v4) {
  for (int pl = 0; pl < po.nsplit; ++pl) {
    uint32_t q[2];
#pragma unroll
    for (int e = 0; e < 4; e += 2) {
      float a0 = v4[e] > 0.f ? v4[e] : v4[e] * po.slope, a1 = v4[e + 1] > 0.f ? v4[e + 1] : v4[e + 1] * po.slope;
      for (int s = 0; s < pl; ++s) {  // peel the planes above this one
        a0 -= __bfloat162float(__float2bfloat16_rn(a0));
        a1 -= __bfloat162float(__float2bfloat16_rn(a1));
      }
      q[e >> 1] = pack_bf16x2(__float2bfloat16_rn(a0), __float2bfloat16_rn(a1));
    }
    *reinterpret_cast<uint2*>(po.at(pl, kc, row) + boff) = make_uint2(q[0], q[1]);
  }
}

// Activation tile of one 32-channel chunk = NSPLIT planes x 4 channel groups, each a contiguous run of rows in the producer's
// plane image: NSPLIT*4 bulk copies of 2-8 KB.  A single lane needs ~100 ns per copy (measured in the persistent GlowTTS
// kernels, profiles/r2a_chain_trace.log), which exceeds the MMA time of a k = 3 chunk; so a whole warp issues them: lane
// groups of NSPLIT*4 lanes take consecutive chunks (= consecutive ring stages), one copy per lane.  `as` / `aph` (next
// stage, parity to wait for on its empty barrier) are warp-uniform ring state and advance by the chunks issued.
// Bulk copies whose global source is not 128-byte aligned run ~3x slower (measured: 72 copies of 2.3 KB took 7.0 us from
// row 39 / 38 and 2.5 us from row 40, profiles/r2b_chain_trace.log), so a tile is copied from the 8-row boundary below its
// first row and the MMA descriptors start `a_row_off(row0)` rows into the stage.
__device__ __forceinline__ uint32_t a_row_off(long long row0) { return (uint32_t)(row0 & 7); }
template <int NSPLIT, int AST>
__device__ __forceinline__ void issue_a_chunks(int lane, int& as, uint32_t& aph, int nch, uint64_t* a_empty, uint64_t* a_full,
                                               unsigned char* sA, uint32_t a_stage, uint32_t a_split_bytes, int RA,
                                               const unsigned char* xp, int xp_kc, int xp_rows, long long row0, int rows) {
  using namespace umma;
  constexpr int CP = NSPLIT * UM_KCH;                    // copies per chunk
  constexpr int GR = (32 / CP) < AST ? (32 / CP) : AST;  // chunks per round
  const int q = lane / CP, j = lane - q * CP;
  const uint32_t roff = a_row_off(row0);
  const long long row0a = row0 - roff;
  const uint32_t kc_bytes = (uint32_t)((rows + (int)roff + 7) & ~7) * 16;
  for (int c0 = 0; c0 < nch; c0 += GR) {
    int sq = as + q;
    uint32_t pq = aph;
    if (sq >= AST) { sq -= AST; pq ^= 1; }
    if (q < GR && c0 + q < nch) {
      mbar_wait(&a_empty[sq], pq);
      if (j == 0) mbar_arrive_expect_tx(&a_full[sq], CP * kc_bytes);
      const int sp = j / UM_KCH, kc = j - sp * UM_KCH;
      bulk_g2s(sA + (size_t)sq * a_stage + sp * a_split_bytes + kc * RA * 16,
               xp + (((size_t)sp * xp_kc + (c0 + q) * UM_KCH + kc) * xp_rows + row0a) * 16, kc_bytes, &a_full[sq]);
    }
    as += min(GR, nch - c0);
    if (as >= AST) { as -= AST; aph ^= 1; }
    __syncwarp();
  }
}

// MT: 128-row tiles per CTA; NSPLIT: bf16 planes per operand; EW: epilogue/gather warps (8 or 16);
// WST: weight ring depth.  <1,2,8,3> fits two CTAs per SM (large grids); EW = 16 variants are for small grids
// where a CTA has the SM to itself and the serial latency of one tile is what matters.
// AST: activation ring depth (per 32-channel chunk); G: conv taps per weight stage (one bulk copy, one wait, one
// commit per stage — the single MMA-issuing warp is the serial resource of the main loop, so fewer, fatter stages).
// NBUF: accumulator buffers in TMEM.  A CTA walks the time tiles blockIdx.x, blockIdx.x + gridDim.x, ... of its
// (n-tile, item/branch); with NBUF = 2 the MMA warp fills one buffer while the epilogue warps drain the other, and the
// operand rings run ahead into the next tile, so the tensor pipe does not idle during epilogues ("persistent" launch:
// gridDim.x < number of tiles).  With gridDim.x = number of tiles every CTA has one tile (the classic launch).
template <int MT, int NSPLIT, int EW, int WST, int AST, int G, int NBUF>
__global__ void __launch_bounds__(EW * 32 + 64 + (NBUF == 2 ? 32 : 0), EW == 8 ? 2 : 1)
    conv_umma_kernel(const __grid_constant__ ConvBatch pb) {
  using namespace umma;
  const ConvParams& p = pb.p[blockIdx.z / pb.zb];
  using C = UmmaCfg<MT>;
  constexpr int RA = C::RA;
  constexpr int UM_WSTAGES = WST;
  constexpr int UM_NPROD = EW * 32;
  // NBUF = 2 (several tiles per CTA): the activation bulk copies get a warp of their own so that they run ahead into the
  // next tile; NBUF = 1 launches give every CTA exactly one tile and thread 0 issues them before it joins the epilogue
  constexpr bool A_WARP = NBUF == 2;
  constexpr int UM_NT_THREADS = UM_NPROD + 64 + (A_WARP ? 32 : 0);
  constexpr int W_MMA = EW, W_LOAD = EW + 1, W_ALOAD = EW + 2;
  static_assert(WST <= UM_WSTAGES_MAX, "weight ring too deep");
  constexpr int NACC = NSPLIT;  // one fp32 accumulator per magnitude class of split products: the tensor core
                                // truncates when it adds into the accumulator, so the long hi*hi chain must not
                                // absorb the small cross terms
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int NT = p.um_nt;
  const uint32_t a_split_bytes = UM_KCH * RA * 16;  // one split plane (hi or lo) of an A stage
  const uint32_t a_stage = NSPLIT * a_split_bytes;
  const uint32_t w_split_bytes = UM_KCH * NT * 16;
  const uint32_t w_stage = NSPLIT * w_split_bytes;
  unsigned char* sA = smem_raw;
  const uint32_t w_ring = G * w_stage;  // bytes of one weight ring slot (G taps)
  unsigned char* sW = smem_raw + AST * a_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + UM_WSTAGES * w_ring);
  uint64_t* a_full = bars;        // [AST <= 4]
  uint64_t* a_empty = bars + 4;   // [AST <= 4]
  uint64_t* w_full = bars + 8;    // [UM_WSTAGES]
  uint64_t* w_empty = bars + 24;  // [UM_WSTAGES]
  uint64_t* acc_full = bars + 40;   // [NBUF <= 2]  MMA -> epilogue
  uint64_t* acc_empty = bars + 42;  // [NBUF <= 2]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 44);
  float* s_bias = reinterpret_cast<float*>(bars + 46);  // [NT] bias of this N tile
  static_assert(NBUF == 1 || NBUF == 2, "one or two accumulator buffers");
  static_assert(AST <= 4, "activation ring too deep");
  float* s_inv = s_bias + 256;                          // [16]      coupling: W^-1
  float* s_anb = s_inv + 16;                            // [2*half]  ActNorm bias
  float* s_ans = s_anb + 256;                           // [2*half]  ActNorm exp(-logs)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z % pb.zb;
  const int ntile = blockIdx.y;
  const int n0 = ntile * NT;
  const int Lb = p.lens ? p.lens[b] * p.len_mul : p.Lin;
  const int m_tiles = (Lb + MT * 128 - 1) / (MT * 128);
  const int gstride = pb.gx[blockIdx.z / pb.zb] > 0 ? pb.gx[blockIdx.z / pb.zb] : (int)gridDim.x;  // CTAs sharing these tiles
  if ((int)blockIdx.x >= m_tiles || (int)blockIdx.x >= gstride) return;
  constexpr uint32_t ACC_COLS = NACC * MT;  // x NT: TMEM columns of one accumulator buffer
  const int K = p.K, dil = p.dil;
  const int nch = (p.Cin + UM_BKC - 1) / UM_BKC;
  const bool planes_in = p.xp != nullptr;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(NBUF * NACC * MT * NT)) tmem_cols <<= 1;

  if (tid == 0) UM_TRACE(0);
  if (tid == 0) {
    for (int i = 0; i < AST; ++i) {
      mbar_init(&a_full[i], planes_in ? 1 : UM_NPROD / 32);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < UM_WSTAGES; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], EW);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, tmem_cols);
  // epilogue constants -> smem (global loads in the epilogue would serialise on L2 latency)
  for (int i = tid; i < NT; i += UM_NT_THREADS) s_bias[i] = (n0 + i) < p.Cout ? p.bias[n0 + i] : 0.f;
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
  // Programmatic dependent launch: let the next kernel in the stream start its prologue / weight prefetch now;
  // it blocks in griddepcontrol.wait before touching anything this grid writes.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == W_LOAD) {
    // ===================== weight producer =====================
    if (lane == 0) {
      const unsigned char* gw = reinterpret_cast<const unsigned char*>(p.wu) + (size_t)ntile * nch * K * w_stage;
      int s = 0;
      uint32_t ph = 1;  // parity to wait for on w_empty (fresh barrier: passes)
      for (int mtile = blockIdx.x; mtile < m_tiles; mtile += gstride)
        for (int c = 0; c < nch; ++c)
          for (int k0 = 0; k0 < K; k0 += G) {
            const uint32_t bytes = (uint32_t)min(G, K - k0) * w_stage;  // taps of one chunk are contiguous in the image
            mbar_wait(&w_empty[s], ph);
            mbar_arrive_expect_tx(&w_full[s], bytes);
            bulk_g2s(sW + (size_t)s * w_ring, gw + ((size_t)c * K + k0) * w_stage, bytes, &w_full[s]);
            if (++s == UM_WSTAGES) { s = 0; ph ^= 1; }
          }
    }
  } else if (A_WARP && warp == W_ALOAD) {
    // ===================== A by bulk copy from the producer's plane image (persistent launches) =====================
    if (planes_in) {
      asm volatile("griddepcontrol.wait;" ::: "memory");  // activations are written by the previous grid
      const unsigned char* xp = reinterpret_cast<const unsigned char*>(p.xp) + (size_t)b * p.xp_bs;
      // only the rows the taps can reach: MT*128 + (K-1)*dil of the RA rows a stage has room for
      const int rows_used = MT * 128 + (K - 1) * dil;
      int as = 0;
      uint32_t aph = 1;
      for (int mtile = blockIdx.x; mtile < m_tiles; mtile += gstride) {
        const long long row0 = (long long)mtile * (MT * 128) - p.pad_left + UM_PADL;
        issue_a_chunks<NSPLIT, AST>(lane, as, aph, nch, a_empty, a_full, sA, a_stage, a_split_bytes, RA, xp, p.xp_kc, p.xp_rows, row0,
                                    rows_used);
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer =====================
    {
      // The whole warp runs this loop with uniform control flow so the descriptor arithmetic lives in uniform
      // registers (UTCHMMA takes its operands from the uniform register file); only lane 0 issues.
      // The weight image keeps the split planes of a k-chunk as adjacent row blocks, so one MMA with N = j*NT multiplies
      // an activation plane with j weight planes at once and lands in j adjacent accumulator classes:
      //   A_hi x [B_hi | B_lo (| B_lo2)] -> [c0 | c1 (| c2)],   A_lo x [B_hi (| B_lo)] -> [c1 (| c2)],   A_lo2 x B_hi -> c2.
      // Fewer, wider MMAs: every MMA re-reads its 128x16 A block from shared memory, which (not the tensor pipe) bounds
      // narrow-N MMAs.
      const uint32_t idesc1 = make_idesc_bf16_f32(128, NT), idesc2 = make_idesc_bf16_f32(128, 2 * NT);
      const uint32_t idesc3 = make_idesc_bf16_f32(128, 3 * NT);
      const bool wide3 = NSPLIT == 3 && 3 * NT <= 256;
      // Descriptors differ only in the 14-bit start-address field (bytes >> 4): build them once and add offsets.
      const uint64_t a_desc0 = make_smem_desc(smem_u32(sA), RA * 16, 128);
      const uint64_t b_desc0 = make_smem_desc(smem_u32(sW), NSPLIT * NT * 16, 128);
      const uint32_t a_ks = (2 * RA * 16) >> 4, b_ks = (2 * NSPLIT * NT * 16) >> 4;  // next K=16 step
      const uint32_t a_sp = a_split_bytes >> 4, b_sp = (NT * 16) >> 4;               // next split plane
      const uint32_t a_mt = (128 * 16) >> 4;                                          // next 128-row tile
      int ws = 0, as = 0;
      uint32_t wph = 0, aph = 0;
      int it = 0;
      for (int mtile = blockIdx.x; mtile < m_tiles; mtile += gstride, ++it) {
      const int buf = NBUF == 2 ? (it & 1) : 0;
      const uint32_t use = NBUF == 2 ? (uint32_t)(it >> 1) : (uint32_t)it;
      mbar_wait(&acc_empty[buf], (use & 1) ^ 1);  // the epilogue has drained this buffer (fresh barrier: passes)
      tc_fence_after();
      const uint32_t tm = tmem + (uint32_t)buf * (ACC_COLS * NT);
      // bulk-copied tiles start at the 8-row boundary below their first row (issue_a_chunks)
      const uint32_t roff = planes_in ? a_row_off((long long)mtile * (MT * 128) - p.pad_left + UM_PADL) : 0u;
      if (lane == 0 && it < 8) UM_TRACE(16 + it * 4 + 0);
      for (int c = 0; c < nch; ++c) {
        mbar_wait(&a_full[as], aph);
        tc_fence_after();
        if (c == 0 && lane == 0 && it == 0) UM_TRACE(3);
        const uint64_t a_c = a_desc0 + ((as * a_stage) >> 4) + roff;
        for (int k0 = 0; k0 < K; k0 += G) {
          mbar_wait(&w_full[ws], wph);
          tc_fence_after();
          const int cnt = min(G, K - k0);
          if (elect_one()) {
            for (int kk = 0; kk < cnt; ++kk) {
              const int k = k0 + kk;
              const uint64_t a_k = a_c + (uint32_t)(k * dil);  // conv tap: shift the image by k*dil rows of 16 B
              const uint64_t b_k = b_desc0 + ((ws * w_ring + kk * w_stage) >> 4);
              const uint32_t first = (c == 0 && k == 0) ? 0u : 1u;
#pragma unroll
              for (int ks = 0; ks < UM_BKC / 16; ++ks) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  const uint64_t a0 = a_k + ks * a_ks + mt * a_mt, b0 = b_k + ks * b_ks;
                  const uint32_t acc0 = ks == 0 ? first : 1u;
                  const uint32_t d0 = tm + (uint32_t)(mt * NACC) * NT;  // classes of one M tile are adjacent columns
                  if (NSPLIT == 2) {
                    mma_bf16_ss(d0, a0, b0, idesc2, acc0);              // hi x [hi|lo] -> [c0|c1]
                    mma_bf16_ss(d0 + NT, a0 + a_sp, b0, idesc1, 1u);    // lo x hi      -> c1
                  } else if (wide3) {
                    mma_bf16_ss(d0, a0, b0, idesc3, acc0);                  // hi  x [hi|lo|lo2] -> [c0|c1|c2]
                    mma_bf16_ss(d0 + NT, a0 + a_sp, b0, idesc2, 1u);        // lo  x [hi|lo]     -> [c1|c2]
                    mma_bf16_ss(d0 + 2 * NT, a0 + 2 * a_sp, b0, idesc1, 1u);  // lo2 x hi          -> c2
                  } else {
                    mma_bf16_ss(d0, a0, b0, idesc2, acc0);                       // hi  x [hi|lo] -> [c0|c1]
                    mma_bf16_ss(d0 + 2 * NT, a0, b0 + 2 * b_sp, idesc1, acc0);   // hi  x lo2     -> c2
                    mma_bf16_ss(d0 + NT, a0 + a_sp, b0, idesc2, 1u);             // lo  x [hi|lo] -> [c1|c2]
                    mma_bf16_ss(d0 + 2 * NT, a0 + 2 * a_sp, b0, idesc1, 1u);     // lo2 x hi      -> c2
                  }
                }
              }
            }
            mma_commit(&w_empty[ws]);  // frees the weight slot when these MMAs have completed
          }
          __syncwarp();
          if (++ws == UM_WSTAGES) { ws = 0; wph ^= 1; }
        }
        if (elect_one()) mma_commit(&a_empty[as]);  // frees the activation stage
        __syncwarp();
        if (++as == AST) { as = 0; aph ^= 1; }
      }
      if (elect_one()) mma_commit(&acc_full[buf]);
      __syncwarp();
      if (lane == 0 && it < 8) UM_TRACE(16 + it * 4 + 1);
      if (lane == 0 && it == 0) UM_TRACE(4);
      }
    }
  } else {
    // everything below reads activations written by the previous grid (weights above are static)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!A_WARP && planes_in && warp == 0) {
      // ===================== A by bulk copy from the producer's plane image (one tile per CTA): warp 0, then it joins the epilogue
      const unsigned char* xp = reinterpret_cast<const unsigned char*>(p.xp) + (size_t)b * p.xp_bs;
      // only the rows the taps can reach: MT*128 + (K-1)*dil of the RA rows a stage has room for
      int as_l = 0;
      uint32_t aph_l = 1;
      const long long row0 = (long long)blockIdx.x * (MT * 128) - p.pad_left + UM_PADL;
      issue_a_chunks<NSPLIT, AST>(lane, as_l, aph_l, nch, a_empty, a_full, sA, a_stage, a_split_bytes, RA, xp, p.xp_kc, p.xp_rows, row0,
                                  MT * 128 + (K - 1) * dil);
    }
    int as = 0;        // gather-mode producer state of the activation ring (runs on across tiles)
    uint32_t aph = 1;
    int it = 0;
    for (int mtile = blockIdx.x; mtile < m_tiles; mtile += gstride, ++it) {
    const int t0 = mtile * (MT * 128);
    if (!planes_in) {
      // ===================== A by gather + split (warps 0..7) =====================
      const float* xb = p.x + (size_t)b * p.x_bs;
      const float* xb2 = p.x2 ? p.x2 + (size_t)b * p.x_bs : nullptr;
      const float* xb3 = p.x3 ? p.x3 + (size_t)b * p.x_bs : nullptr;
      const int rows_used = MT * 128 + (K - 1) * dil;
      const bool three = xb2 != nullptr && xb3 != nullptr;  // MRF mean of three inputs (uniform)
      for (int c = 0; c < nch; ++c, as = (as + 1 == AST ? 0 : as + 1), aph ^= (as == 0)) {
        mbar_wait(&a_empty[as], aph);
        unsigned char* dst = sA + (size_t)as * a_stage;
        constexpr int NTASK = (UM_KCH * RA + UM_NPROD - 1) / UM_NPROD;
        constexpr int TB = NTASK < 4 ? NTASK : (MT == 1 ? 4 : 3);  // tasks per batch: TB x 8 loads in flight per thread
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
            float w[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float a = v[i][e];
              if (p.in_div != 1.f) a = a / p.in_div;
              if (p.pre == PRE_LRELU) a = a > 0.f ? a : a * p.pre_slope;
              w[e] = a;
            }
            uint4 q[3];
            split8<NSPLIT>(w, q);
            unsigned char* o = dst + so[i];
#pragma unroll
            for (int sp = 0; sp < NSPLIT; ++sp) *reinterpret_cast<uint4*>(o + sp * a_split_bytes) = q[sp];
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[as]);
        if (tid == 0 && c == 0 && it == 0) UM_TRACE(2);
        if (tid == 0 && c == nch - 1 && it == 0) UM_TRACE(5);
      }
    }

    // ===================== epilogue (warps 0..EW-1) =====================
    const int buf = NBUF == 2 ? (it & 1) : 0;
    const uint32_t use = NBUF == 2 ? (uint32_t)(it >> 1) : (uint32_t)it;
    mbar_wait(&acc_full[buf], use & 1);
    tc_fence_after();
    const uint32_t tm = tmem + (uint32_t)buf * (ACC_COLS * NT);
    if (tid == 0 && it == 0) UM_TRACE(6);
    if (tid == 0 && it < 8) UM_TRACE(16 + it * 4 + 2);
#include "conv_umma_epilogue.inc"
    // this warp is done reading the accumulator buffer: hand it back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&acc_empty[buf]);
    if (tid == 0 && it < 8) UM_TRACE(16 + it * 4 + 3);
    }
  }
  if (tid == 0) UM_TRACE(7);
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, tmem_cols);
  if (tid == 0) UM_TRACE(8);
}

}  // namespace b2
