// attn_umma.cu — relative-position multi-head self-attention (glow_tts/attentions.py:214-264) on tcgen05.
//
//   S = Q K^T            M = queries (128 rows), N = keys (128), K = d_k (96)     -> TMEM
//   s[i][j] = (S[i][j] + [|j-i|<=w] q_i.Ek[j-i+w]) / sqrt(dk);  p = softmax_j over j < len
//   O = P V              M = queries, N = d_k (96), K = keys (128)                 -> TMEM
//   o_i += sum_{|r|<=w} p[i][i+r] Ev[r+w]
//
// All operands are bf16 3-way split images (fp32-equivalent products, six MMAs per K=16 step):
//   * Q, K, V come straight from the plane image the fused q/k/v 1x1 conv epilogue wrote
//     ([plane][576/8][PADL+rows][8]); head h owns 12 consecutive 8-channel groups of each of q, k, v.
//     Q (rows = queries) and K (rows = keys) are K-major operands; V is read as an MN-major B operand
//     (8 channels contiguous, keys along rows) — the same bytes, only the descriptor differs.
//   * each TMEM lane is one query, so thread i owns row i of S: the softmax is thread-local (no shuffles);
//     the thread writes its row of P as a K-major image (rows = queries, k-chunks = keys/8) into the shared
//     memory that held Q and K.
// One CTA per (head, batch item) for sequences up to 128 phonemes; attn_umma2_kernel below covers 129..256 (longer ones use
// rel_attention_kernel).
// Output: the attention result as a plane image for conv_o (+ optional fp32 [H][T]).
#include "kernels.cuh"
#include "umma.cuh"

namespace b2 {

constexpr int AT_ROWS = 128;   // queries == keys tile
constexpr int AT_DK = 96;      // head dim
constexpr int AT_KC = AT_DK / 8;
constexpr int AT_NS = 3;       // split planes
constexpr int AT_PADL = 40;    // == UM_PADL
constexpr int AT_THREADS = 128 + 64;

struct AttnParams {
  const unsigned char* qkv_p;  // plane image of the fused qkv conv output
  long long qkv_bs;
  int qkv_kc, qkv_rows;
  const unsigned char* relk_img;  // [3 planes][12][16][8] bf16: rows r = Ek[r] (r < 2w+1, rest zero)
  const unsigned char* relv_img;  // same for Ev
  int window, H, n_heads;
  const int* lens;
  unsigned char* out_p;        // plane image of the output [H channels], nsplit 3
  long long out_bs;
  int out_kc, out_rows;
  float* out;                  // optional fp32 [H][out_ld]
  long long out_fbs;
  int out_ld;
};

__device__ __forceinline__ uint32_t at_cvt2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ void at_split8(const float* v, uint4* out) {
  uint32_t h[4], l[4], l2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float a = v[2 * e], b = v[2 * e + 1];
    h[e] = at_cvt2(a, b);
    const float ra = a - __uint_as_float(h[e] << 16), rb = b - __uint_as_float(h[e] & 0xffff0000u);
    l[e] = at_cvt2(ra, rb);
    l2[e] = at_cvt2(ra - __uint_as_float(l[e] << 16), rb - __uint_as_float(l[e] & 0xffff0000u));
  }
  out[0] = make_uint4(h[0], h[1], h[2], h[3]);
  out[1] = make_uint4(l[0], l[1], l[2], l[3]);
  out[2] = make_uint4(l2[0], l2[1], l2[2], l2[3]);
}

__device__ __forceinline__ uint32_t make_idesc_bf16_f32_bmn(uint32_t M, uint32_t N) {
  return umma::make_idesc_bf16_f32(M, N) | (1u << 16);  // b_major = MN
}

// The relative-position terms ride on the same MMAs: the 2w+1 rows of Ek are appended to the key operand
// (S' = Q [K; Ek]^T has the relative-key logits in columns 128..136) and the rows of Ev are appended to the
// value operand, with the band probabilities appended to P as two extra key chunks (O = [P, P_band] [V; Ev]).
constexpr int AT_XR = AT_ROWS + 16;  // rows of the extended key / value operands

__global__ void __launch_bounds__(AT_THREADS, 1) attn_umma_kernel(const __grid_constant__ AttnParams p) {
  using namespace umma;
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr uint32_t QPLANE = AT_KC * AT_ROWS * 16;          // [12][128][8]   24 KB
  constexpr uint32_t XPLANE = AT_KC * AT_XR * 16;            // [12][144][8]   27 KB
  constexpr uint32_t PPLANE = (AT_XR / 8) * AT_ROWS * 16;    // [18][128][8]   36 KB
  // phase 1: Q | K'            phase 2 (Q, K' dead): V' | P'
  unsigned char* sQ = smem;
  unsigned char* sK = smem + AT_NS * QPLANE;
  unsigned char* sV = smem;
  unsigned char* sP = smem + AT_NS * XPLANE;
  constexpr uint32_t DATA = AT_NS * XPLANE + AT_NS * PPLANE;  // 193536 >= phase 1 (156672)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DATA);
  uint64_t* qk_full = bars;      // Q, K' landed
  uint64_t* s_full = bars + 1;   // S accumulated (Q, K' dead)
  uint64_t* v_full = bars + 2;   // V' landed
  uint64_t* p_full = bars + 3;   // P' image written (128 arrivals)
  uint64_t* o_full = bars + 4;   // O accumulated
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* s_relk = reinterpret_cast<float*>(bars + 8);  // [128][9] relative-key logits of each query

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int len = p.lens[b];
  if (len <= 0) return;
  const int nrel = 2 * p.window + 1;

  if (tid == 0) {
    mbar_init(qk_full, 1);
    mbar_init(s_full, 1);
    mbar_init(v_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem + AT_XR + 16;  // S': 144 columns (reads touch 160), O: 96 columns
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 5) {
    if (lane == 0) {
      const unsigned char* src = p.qkv_p + (size_t)b * p.qkv_bs;
      const int hk = p.H / 8;  // channel groups per q / k / v block
      auto img = [&](int op, int sp, int kc) {
        const size_t g = (size_t)op * hk + (size_t)h * AT_KC + kc;
        return src + (((size_t)sp * p.qkv_kc + g) * p.qkv_rows + AT_PADL) * 16;
      };
      // phase 1: Q rows [0,128) and K' = [K rows [0,128); Ek rows] (image rows >= len are zero)
      mbar_arrive_expect_tx(qk_full, AT_NS * (QPLANE + XPLANE));
      for (int sp = 0; sp < AT_NS; ++sp)
        for (int kc = 0; kc < AT_KC; ++kc) {
          bulk_g2s(sQ + sp * QPLANE + kc * AT_ROWS * 16, img(0, sp, kc), AT_ROWS * 16, qk_full);
          bulk_g2s(sK + sp * XPLANE + kc * AT_XR * 16, img(1, sp, kc), AT_ROWS * 16, qk_full);
          bulk_g2s(sK + sp * XPLANE + kc * AT_XR * 16 + AT_ROWS * 16, p.relk_img + ((size_t)sp * AT_KC + kc) * 256, 256, qk_full);
        }
      // phase 2: V' = [V; Ev] over the dead Q / K' bytes, as soon as the S MMAs have retired
      mbar_wait(s_full, 0);
      mbar_arrive_expect_tx(v_full, AT_NS * XPLANE);
      for (int sp = 0; sp < AT_NS; ++sp)
        for (int kc = 0; kc < AT_KC; ++kc) {
          bulk_g2s(sV + sp * XPLANE + kc * AT_XR * 16, img(2, sp, kc), AT_ROWS * 16, v_full);
          bulk_g2s(sV + sp * XPLANE + kc * AT_XR * 16 + AT_ROWS * 16, p.relv_img + ((size_t)sp * AT_KC + kc) * 256, 256, v_full);
        }
    }
  } else if (warp == 4) {
    // ---- MMA issuer (uniform control flow, one elected lane issues)
    const uint32_t idS = make_idesc_bf16_f32(128, AT_XR);
    const uint32_t idO = make_idesc_bf16_f32_bmn(128, AT_DK);
    mbar_wait(qk_full, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint64_t qd = make_smem_desc(smem_u32(sQ), AT_ROWS * 16, 128);
      const uint64_t kd = make_smem_desc(smem_u32(sK), AT_XR * 16, 128);
      const uint32_t qks = (2 * AT_ROWS * 16) >> 4, kks = (2 * AT_XR * 16) >> 4, qsp = QPLANE >> 4, ksp = XPLANE >> 4;
#pragma unroll
      for (int k = 0; k < AT_DK / 16; ++k) {
        const uint64_t a = qd + k * qks, bb = kd + k * kks;
        const uint32_t acc = k ? 1u : 0u;
        mma_bf16_ss(tS, a + qsp, bb + ksp, idS, acc);     // lo*lo
        mma_bf16_ss(tS, a, bb + 2 * ksp, idS, 1u);        // hi*lo2
        mma_bf16_ss(tS, a + 2 * qsp, bb, idS, 1u);        // lo2*hi
        mma_bf16_ss(tS, a, bb + ksp, idS, 1u);            // hi*lo
        mma_bf16_ss(tS, a + qsp, bb, idS, 1u);            // lo*hi
        mma_bf16_ss(tS, a, bb, idS, 1u);                  // hi*hi
      }
      mma_commit(s_full);
    }
    __syncwarp();
    mbar_wait(v_full, 0);
    mbar_wait(p_full, 0);
    tc_fence_after();
    if (elect_one()) {
      // A = P' image [18 key chunks][128 rows][8], K-major;  B = V' [12 channel groups][144 keys][8], MN-major:
      // LBO = 128 B (next 8 keys), SBO = 144*16 B (next 8 channels)
      const uint64_t pd = make_smem_desc(smem_u32(sP), AT_ROWS * 16, 128);
      const uint64_t vd = make_smem_desc(smem_u32(sV), 128, AT_XR * 16);
      const uint32_t pks = (2 * AT_ROWS * 16) >> 4, psp = PPLANE >> 4;
      const uint32_t vks = (16 * 16) >> 4, vsp = XPLANE >> 4;
#pragma unroll
      for (int k = 0; k < AT_XR / 16; ++k) {
        const uint64_t a = pd + k * pks, bb = vd + k * vks;
        const uint32_t acc = k ? 1u : 0u;
        mma_bf16_ss(tO, a + psp, bb + vsp, idO, acc);
        mma_bf16_ss(tO, a, bb + 2 * vsp, idO, 1u);
        mma_bf16_ss(tO, a + 2 * psp, bb, idO, 1u);
        mma_bf16_ss(tO, a, bb + vsp, idO, 1u);
        mma_bf16_ss(tO, a + psp, bb, idO, 1u);
        mma_bf16_ss(tO, a, bb, idO, 1u);
      }
      mma_commit(o_full);
    }
    __syncwarp();
  } else {
    // ---- softmax + epilogue: thread = query row = TMEM lane
    const int i = tid;  // 0..127
    const bool row_ok = i < len;
    const float inv_sqrt = 1.f / sqrtf((float)AT_DK);
    float* my_relk = s_relk + i * 9;
    mbar_wait(s_full, 0);
    tc_fence_after();
    const uint32_t lane_addr = ((uint32_t)(warp * 32) << 16);
    {
      uint32_t raw[32];
      tmem_ld32(tS + lane_addr + AT_ROWS, raw);  // columns 128..159: q_i . Ek[r] in the first 2w+1
      tmem_ld_wait();
#pragma unroll
      for (int r = 0; r < 9; ++r) my_relk[r] = __uint_as_float(raw[r]);
    }
    float srow[AT_ROWS];
    float mx = -INFINITY;
#pragma unroll
    for (int g = 0; g < AT_ROWS / 32; ++g) {
      uint32_t raw[32];
      tmem_ld32(tS + lane_addr + g * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = g * 32 + j;
        float s = __uint_as_float(raw[j]);
        const int r = col - i + p.window;
        if (r >= 0 && r < nrel) s += my_relk[r];
        s *= inv_sqrt;
        s = col < len ? s : -INFINITY;
        srow[col] = s;
        mx = fmaxf(mx, s);
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int col = 0; col < AT_ROWS; ++col) {
      const float e = col < len ? __expf(srow[col] - mx) : 0.f;
      srow[col] = e;
      sum += e;
    }
    const float inv = row_ok ? 1.f / sum : 0.f;
    // P' image over the dead K' bytes (the S MMAs have retired: s_full)
    float band[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) band[r] = 0.f;
#pragma unroll
    for (int kc = 0; kc < AT_ROWS / 8; ++kc) {
      float v8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v8[e] = srow[kc * 8 + e] * inv;
      uint4 q[3];
      at_split8(v8, q);
      unsigned char* o = sP + ((size_t)kc * AT_ROWS + i) * 16;
      *reinterpret_cast<uint4*>(o) = q[0];
      *reinterpret_cast<uint4*>(o + PPLANE) = q[1];
      *reinterpret_cast<uint4*>(o + 2 * PPLANE) = q[2];
    }
    // band probabilities p[i][i + r - w] as 16 extra "keys" (they meet the Ev rows of V')
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int col = i + r - p.window;
      float pv = 0.f;
      if (r < nrel && col >= 0 && col < len) {
        float sv = 0.f;
#pragma unroll
        for (int c2 = 0; c2 < AT_ROWS; ++c2) sv = c2 == col ? srow[c2] : sv;  // register select, no local memory
        pv = sv * inv;
      }
      band[r] = pv;
    }
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
      uint4 q[3];
      at_split8(band + kc * 8, q);
      unsigned char* o = sP + ((size_t)(AT_ROWS / 8 + kc) * AT_ROWS + i) * 16;
      *reinterpret_cast<uint4*>(o) = q[0];
      *reinterpret_cast<uint4*>(o + PPLANE) = q[1];
      *reinterpret_cast<uint4*>(o + 2 * PPLANE) = q[2];
    }
    fence_proxy_async_smem();
    mbar_arrive(p_full);

    mbar_wait(o_full, 0);
    tc_fence_after();
    unsigned char* op = p.out_p ? p.out_p + (size_t)b * p.out_bs : nullptr;
    const size_t ostride = (size_t)p.out_kc * p.out_rows * 16;
    float* of = p.out ? p.out + (size_t)b * p.out_fbs + (size_t)(h * AT_DK) * p.out_ld + i : nullptr;
#pragma unroll
    for (int g = 0; g < AT_DK / 32; ++g) {
      uint32_t raw[32];
      tmem_ld32(tO + lane_addr + g * 32, raw);
      tmem_ld_wait();
      float o[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float a = __uint_as_float(raw[j]);
        o[j] = row_ok ? a : 0.f;
        if (of && row_ok) of[(size_t)(g * 32 + j) * p.out_ld] = a;
      }
      if (op) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 q[3];
          at_split8(o + q4 * 8, q);
          const int kc = h * AT_KC + g * 4 + q4;
          unsigned char* dst = op + ((size_t)kc * p.out_rows + AT_PADL + i) * 16;
          *reinterpret_cast<uint4*>(dst) = q[0];
          *reinterpret_cast<uint4*>(dst + ostride) = q[1];
          *reinterpret_cast<uint4*>(dst + 2 * ostride) = q[2];
          // zero rows around the data: t in [-PADL, 0) by rows i < PADL, t in [128, 384) by all rows
          const uint4 z = make_uint4(0, 0, 0, 0);
          if (i < AT_PADL) {
            unsigned char* zf = op + ((size_t)kc * p.out_rows + i) * 16;
            *reinterpret_cast<uint4*>(zf) = z;
            *reinterpret_cast<uint4*>(zf + ostride) = z;
            *reinterpret_cast<uint4*>(zf + 2 * ostride) = z;
          }
#pragma unroll
          for (int rep = 1; rep <= 2; ++rep) {
            unsigned char* zt = op + ((size_t)kc * p.out_rows + AT_PADL + rep * AT_ROWS + i) * 16;
            *reinterpret_cast<uint4*>(zt) = z;
            *reinterpret_cast<uint4*>(zt + ostride) = z;
            *reinterpret_cast<uint4*>(zt + 2 * ostride) = z;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------
// Sequences of 129..256 phonemes: the same attention with two query tiles (one CTA per (query tile, head, item)) and the
// keys / values in blocks of 128 that pass through one shared-memory buffer one after the other:
//   S'[:, 128 kb ..] = Q [K_kb (; Ek)]^T   for kb = 0, 1        -> all of S' (128 x 272 fp32) stays in TMEM
//   softmax over the whole row (2 threads per query row: thread (i, hf) owns the columns of key block hf), in three passes
//   over TMEM (row maximum | sum of exponentials | probabilities), so no thread holds a row in registers
//   O (+)= [P_kb (, P_band)] [V_kb (; Ev)]  for kb = 0, 1        -> 128 x 96 in TMEM
// The relative-position rows Ek / Ev and the band probabilities ride on the LAST block (16 extra operand rows), as above.
// Shared memory: region A = Q, later the P block image; region B = the current K / V block.
// ------------------------------------------------------------------------------------------------
constexpr int AT2_THREADS = 256 + 64;

__global__ void __launch_bounds__(AT2_THREADS, 1) attn_umma2_kernel(const __grid_constant__ AttnParams p) {
  using namespace umma;
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr uint32_t QPLANE = AT_KC * AT_ROWS * 16;          // [12][128][8]   24 KB
  constexpr uint32_t XPLANE = AT_KC * AT_XR * 16;            // [12][144][8]   27 KB
  constexpr uint32_t PPLANE = (AT_XR / 8) * AT_ROWS * 16;    // [18][128][8]   36 KB
  unsigned char* sQ = smem;                    // region A (110592 B): Q (73728 B), then the P block
  unsigned char* sP = smem;
  unsigned char* sX = smem + AT_NS * PPLANE;   // region B (82944 B): K block, then V block
  constexpr uint32_t DATA = AT_NS * PPLANE + AT_NS * XPLANE;  // 193536
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DATA);
  uint64_t* q_full = bars;        // Q landed
  uint64_t* k_full = bars + 1;    // [2] K block landed
  uint64_t* s_done = bars + 3;    // [2] S block accumulated (the K block is dead)
  uint64_t* v_full = bars + 5;    // [2] V block landed
  uint64_t* p_full = bars + 7;    // [2] P block image written (256 arrivals)
  uint64_t* o_done = bars + 9;    // [2] O block accumulated (P and V block dead)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  float* s_relk = reinterpret_cast<float*>(bars + 12);  // [2][128][9] relative-key logits of each query (a copy per thread)
  float* s_band = s_relk + 2 * 128 * 9;                 // [128][9] scaled, masked logits on the band |j - i| <= window
  float* s_mx = s_band + 128 * 9;                       // [2][128]
  float* s_sum = s_mx + 2 * 128;                        // [2][128]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int len = p.lens[b];
  if (qt * AT_ROWS >= len) return;
  const int nkb = len > AT_ROWS ? 2 : 1;  // key blocks; the last one carries the relative-position rows
  const int nrel = 2 * p.window + 1;
  const int n_qt = (len + AT_ROWS - 1) / AT_ROWS;

  if (tid == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&s_done[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&p_full[i], 256);
      mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem + 288;  // S': up to 272 columns, O: 96 columns
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 9) {
    // ---- loader (every lane issues its share of the small per-channel-group copies)
    const unsigned char* src = p.qkv_p + (size_t)b * p.qkv_bs;
    const int hk = p.H / 8;  // channel groups per q / k / v block
    auto img = [&](int op, int sp, int kc, int row) {
      const size_t g = (size_t)op * hk + (size_t)h * AT_KC + kc;
      return src + (((size_t)sp * p.qkv_kc + g) * p.qkv_rows + AT_PADL + row) * 16;
    };
    if (lane == 0) mbar_arrive_expect_tx(q_full, AT_NS * QPLANE);
    __syncwarp();
    for (int i = lane; i < AT_NS * AT_KC; i += 32) {
      const int sp = i / AT_KC, kc = i - sp * AT_KC;
      bulk_g2s(sQ + sp * QPLANE + kc * AT_ROWS * 16, img(0, sp, kc, qt * AT_ROWS), AT_ROWS * 16, q_full);
    }
    for (int ph = 0; ph < 2; ++ph) {      // ph 0: key blocks, ph 1: value blocks
      for (int kb = 0; kb < nkb; ++kb) {
        // the buffer is free when the previous user's MMAs have retired
        if (ph == 0 && kb > 0) mbar_wait(&s_done[kb - 1], 0);
        if (ph == 1) mbar_wait(kb == 0 ? &s_done[nkb - 1] : &o_done[kb - 1], 0);
        const bool last = kb == nkb - 1;
        uint64_t* bar = ph == 0 ? &k_full[kb] : &v_full[kb];
        if (lane == 0) mbar_arrive_expect_tx(bar, AT_NS * AT_KC * (AT_ROWS * 16 + (last ? 256 : 0)));
        __syncwarp();
        const unsigned char* rel = ph == 0 ? p.relk_img : p.relv_img;
        for (int i = lane; i < AT_NS * AT_KC; i += 32) {
          const int sp = i / AT_KC, kc = i - sp * AT_KC;
          bulk_g2s(sX + sp * XPLANE + kc * AT_XR * 16, img(ph == 0 ? 1 : 2, sp, kc, kb * AT_ROWS), AT_ROWS * 16, bar);
          if (last) bulk_g2s(sX + sp * XPLANE + kc * AT_XR * 16 + AT_ROWS * 16, rel + ((size_t)sp * AT_KC + kc) * 256, 256, bar);
        }
      }
    }
  } else if (warp == 8) {
    // ---- MMA issuer
    const uint64_t qd = make_smem_desc(smem_u32(sQ), AT_ROWS * 16, 128);
    const uint64_t kd = make_smem_desc(smem_u32(sX), AT_XR * 16, 128);
    const uint32_t qks = (2 * AT_ROWS * 16) >> 4, kks = (2 * AT_XR * 16) >> 4, qsp = QPLANE >> 4, ksp = XPLANE >> 4;
    mbar_wait(q_full, 0);
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t idS = make_idesc_bf16_f32(128, kb == nkb - 1 ? AT_XR : AT_ROWS);
      mbar_wait(&k_full[kb], 0);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tS + kb * AT_ROWS;
#pragma unroll
        for (int k = 0; k < AT_DK / 16; ++k) {
          const uint64_t a = qd + k * qks, bb = kd + k * kks;
          const uint32_t acc = k ? 1u : 0u;
          mma_bf16_ss(d, a + qsp, bb + ksp, idS, acc);     // lo*lo
          mma_bf16_ss(d, a, bb + 2 * ksp, idS, 1u);        // hi*lo2
          mma_bf16_ss(d, a + 2 * qsp, bb, idS, 1u);        // lo2*hi
          mma_bf16_ss(d, a, bb + ksp, idS, 1u);            // hi*lo
          mma_bf16_ss(d, a + qsp, bb, idS, 1u);            // lo*hi
          mma_bf16_ss(d, a, bb, idS, 1u);                  // hi*hi
        }
        mma_commit(&s_done[kb]);
      }
      __syncwarp();
    }
    const uint32_t idO = make_idesc_bf16_f32_bmn(128, AT_DK);
    const uint64_t pd = make_smem_desc(smem_u32(sP), AT_ROWS * 16, 128);
    const uint64_t vd = make_smem_desc(smem_u32(sX), 128, AT_XR * 16);
    const uint32_t pks = (2 * AT_ROWS * 16) >> 4, psp = PPLANE >> 4;
    const uint32_t vks = (16 * 16) >> 4, vsp = XPLANE >> 4;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(&v_full[kb], 0);
      mbar_wait(&p_full[kb], 0);
      tc_fence_after();
      if (elect_one()) {
        const int ksteps = (kb == nkb - 1 ? AT_XR : AT_ROWS) / 16;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t a = pd + k * pks, bb = vd + k * vks;
          const uint32_t acc = (kb | k) ? 1u : 0u;
          mma_bf16_ss(tO, a + psp, bb + vsp, idO, acc);
          mma_bf16_ss(tO, a, bb + 2 * vsp, idO, 1u);
          mma_bf16_ss(tO, a + 2 * psp, bb, idO, 1u);
          mma_bf16_ss(tO, a, bb + vsp, idO, 1u);
          mma_bf16_ss(tO, a + psp, bb, idO, 1u);
          mma_bf16_ss(tO, a, bb, idO, 1u);
        }
        mma_commit(&o_done[kb]);
      }
      __syncwarp();
    }
  } else {
    // ---- softmax + epilogue: thread = (query row = TMEM lane, half); half hf owns the columns of key block hf
    const int i = (warp & 3) * 32 + lane;  // row of this tile
    const int hf = warp >> 2;
    const int ig = qt * AT_ROWS + i;       // query position
    const bool row_ok = ig < len;
    const float inv_sqrt = 1.f / sqrtf((float)AT_DK);
    float* my_relk = s_relk + (hf * 128 + i) * 9;
    float* my_band = s_band + i * 9;
    const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32) << 16);
    auto epi_bar = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    mbar_wait(&s_done[nkb - 1], 0);  // (the blocks complete in order)
    tc_fence_after();
    {
      uint32_t raw[16];
      tmem_ld16(tS + lane_addr + nkb * AT_ROWS, raw);  // q_i . Ek[r] in the first 2w+1 columns behind the keys
      tmem_ld_wait();
#pragma unroll
      for (int r = 0; r < 9; ++r) my_relk[r] = __uint_as_float(raw[r]);
    }
    // logit of column c of key block kb from its raw accumulator value
    auto logit = [&](float raw, int kb, int c, bool store_band) {
      const int jg = kb * AT_ROWS + c;  // key position
      const int r = jg - ig + p.window;
      const bool in_band = (unsigned)r < (unsigned)nrel;
      float sv = raw;
      if (in_band) sv += my_relk[r];
      sv *= inv_sqrt;
      sv = jg < len ? sv : -INFINITY;
      if (in_band && store_band) my_band[r] = sv;
      return sv;
    };
    // pass 1: row maximum (each half over its own key block)
    float mx = -INFINITY;
    if (hf < nkb) {
#pragma unroll
      for (int g = 0; g < AT_ROWS / 32; ++g) {
        uint32_t raw[32];
        tmem_ld32(tS + lane_addr + hf * AT_ROWS + g * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, logit(__uint_as_float(raw[j]), hf, g * 32 + j, true));
      }
    }
    s_mx[hf * 128 + i] = mx;
    epi_bar();
    mx = fmaxf(s_mx[i], s_mx[128 + i]);
    // pass 2: sum of exponentials
    float ssum = 0.f;
    if (hf < nkb) {
#pragma unroll
      for (int g = 0; g < AT_ROWS / 32; ++g) {
        uint32_t raw[32];
        tmem_ld32(tS + lane_addr + hf * AT_ROWS + g * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float sv = logit(__uint_as_float(raw[j]), hf, g * 32 + j, false);
          ssum += hf * AT_ROWS + g * 32 + j < len ? __expf(sv - mx) : 0.f;
        }
      }
    }
    s_sum[hf * 128 + i] = ssum;
    epi_bar();
    ssum = s_sum[i] + s_sum[128 + i];
    const float inv = row_ok ? 1.f / ssum : 0.f;
    // pass 3: the P image of each key block (both halves write 8 of its 16 key chunks), as soon as region A is free
    mbar_wait(q_full, 0);  // (acquire of the loader's writes is not needed; Q is dead once the last S block is done)
    for (int kb = 0; kb < nkb; ++kb) {
      if (kb > 0) {
        mbar_wait(&o_done[kb - 1], 0);  // the previous P block has been read
        tc_fence_after();
      }
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t raw[32];
        tmem_ld32(tS + lane_addr + kb * AT_ROWS + hf * 64 + g * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int q8 = 0; q8 < 4; ++q8) {
          float v8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = hf * 64 + g * 32 + q8 * 8 + e;
            const float sv = logit(__uint_as_float(raw[q8 * 8 + e]), kb, c, false);
            v8[e] = kb * AT_ROWS + c < len ? __expf(sv - mx) * inv : 0.f;
          }
          uint4 q3[3];
          at_split8(v8, q3);
          unsigned char* o = sP + ((size_t)(hf * 8 + g * 4 + q8) * AT_ROWS + i) * 16;
          *reinterpret_cast<uint4*>(o) = q3[0];
          *reinterpret_cast<uint4*>(o + PPLANE) = q3[1];
          *reinterpret_cast<uint4*>(o + 2 * PPLANE) = q3[2];
        }
      }
      if (kb == nkb - 1) {
        // band probabilities p[i][i + r - w] as 16 extra "keys" (they meet the Ev rows): half 0 writes r = 0..7, half 1 the rest
        float band[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int r = hf * 8 + e;
          const int jg = ig + r - p.window;
          float pv = 0.f;
          if (r < nrel && jg >= 0 && jg < len) pv = __expf(my_band[r < 9 ? r : 0] - mx) * inv;
          band[e] = pv;
        }
        uint4 q3[3];
        at_split8(band, q3);
        unsigned char* o = sP + ((size_t)(AT_ROWS / 8 + hf) * AT_ROWS + i) * 16;
        *reinterpret_cast<uint4*>(o) = q3[0];
        *reinterpret_cast<uint4*>(o + PPLANE) = q3[1];
        *reinterpret_cast<uint4*>(o + 2 * PPLANE) = q3[2];
      }
      fence_proxy_async_smem();
      mbar_arrive(&p_full[kb]);
    }
    // epilogue: 48 of the head's 96 output channels per thread
    mbar_wait(&o_done[nkb - 1], 0);
    tc_fence_after();
    unsigned char* op = p.out_p ? p.out_p + (size_t)b * p.out_bs : nullptr;
    const size_t ostride = (size_t)p.out_kc * p.out_rows * 16;
    float* of = p.out ? p.out + (size_t)b * p.out_fbs + (size_t)(h * AT_DK) * p.out_ld + ig : nullptr;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int c0 = hf * (AT_DK / 2) + g * 16;
      uint32_t raw[16];
      tmem_ld16(tO + lane_addr + c0, raw);
      tmem_ld_wait();
      float o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float a = __uint_as_float(raw[j]);
        o[j] = row_ok ? a : 0.f;
        if (of && row_ok) of[(size_t)(c0 + j) * p.out_ld] = a;
      }
      if (op) {
#pragma unroll
        for (int q4 = 0; q4 < 2; ++q4) {
          uint4 q3[3];
          at_split8(o + q4 * 8, q3);
          const int kc = h * AT_KC + c0 / 8 + q4;
          unsigned char* dst = op + ((size_t)kc * p.out_rows + AT_PADL + ig) * 16;
          *reinterpret_cast<uint4*>(dst) = q3[0];
          *reinterpret_cast<uint4*>(dst + ostride) = q3[1];
          *reinterpret_cast<uint4*>(dst + 2 * ostride) = q3[2];
          // zero rows around the data: t in [-PADL, 0) by the first query tile, the 256 rows behind the last tile by the last one
          const uint4 z = make_uint4(0, 0, 0, 0);
          if (qt == 0 && i < AT_PADL) {
            unsigned char* zf = op + ((size_t)kc * p.out_rows + i) * 16;
            *reinterpret_cast<uint4*>(zf) = z;
            *reinterpret_cast<uint4*>(zf + ostride) = z;
            *reinterpret_cast<uint4*>(zf + 2 * ostride) = z;
          }
          if (qt == n_qt - 1) {
#pragma unroll
            for (int rep = 0; rep < 2; ++rep) {
              unsigned char* zt = op + ((size_t)kc * p.out_rows + AT_PADL + (n_qt + rep) * AT_ROWS + i) * 16;
              *reinterpret_cast<uint4*>(zt) = z;
              *reinterpret_cast<uint4*>(zt + ostride) = z;
              *reinterpret_cast<uint4*>(zt + 2 * ostride) = z;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

void attn_umma_init_attributes() {
  B2_CUDA(cudaFuncSetAttribute(attn_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  B2_CUDA(cudaFuncSetAttribute(attn_umma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
}

constexpr int AT_TMAX = 2 * AT_ROWS;  // attn_umma2_kernel: two query tiles, two key blocks

bool attn_umma_supported(int H, int n_heads, int window, int Tmax) {
  return H % n_heads == 0 && H / n_heads == AT_DK && window <= 4 && Tmax <= AT_TMAX && H % 32 == 0;
}

void launch_attn_umma(const void* qkv_p, long long qkv_bs, int qkv_kc, int qkv_rows, const void* relk_img,
                      const void* relv_img, int window, int H, int n_heads, const int* lens,
                      int batch, void* out_p, long long out_bs, int out_kc, int out_rows, float* out, long long out_fbs,
                      int out_ld, cudaStream_t s, int Tmax) {
  AttnParams p;
  p.qkv_p = (const unsigned char*)qkv_p; p.qkv_bs = qkv_bs; p.qkv_kc = qkv_kc; p.qkv_rows = qkv_rows;
  p.relk_img = (const unsigned char*)relk_img; p.relv_img = (const unsigned char*)relv_img; p.window = window; p.H = H; p.n_heads = n_heads; p.lens = lens;
  p.out_p = (unsigned char*)out_p; p.out_bs = out_bs; p.out_kc = out_kc; p.out_rows = out_rows;
  p.out = out; p.out_fbs = out_fbs; p.out_ld = out_ld;
  B2_CHECK(window <= 4, "attention: relative window too wide");
  B2_CHECK(Tmax >= 1 && Tmax <= AT_TMAX, "attention: sequence too long for the tcgen05 kernels");
  const bool two = Tmax > AT_ROWS;
  const int n_qt = two ? 2 : 1;
  B2_CHECK(out_rows >= AT_PADL + (n_qt + 2) * AT_ROWS, "attention: output plane image too short");
  B2_CHECK(qkv_rows >= AT_PADL + n_qt * AT_ROWS, "attention: qkv plane image too short");
  const size_t data = (size_t)AT_NS * AT_KC * AT_XR * 16 + (size_t)AT_NS * (AT_XR / 8) * AT_ROWS * 16;
  const size_t smem = two ? data + 128 + sizeof(float) * (3 * 128 * 9 + 4 * 128) : data + 64 + AT_ROWS * 9 * sizeof(float);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = two ? dim3(n_qt, n_heads, batch) : dim3(n_heads, batch);
  cfg.blockDim = dim3(two ? AT2_THREADS : AT_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (two) B2_CUDA(cudaLaunchKernelEx(&cfg, attn_umma2_kernel, p));
  else B2_CUDA(cudaLaunchKernelEx(&cfg, attn_umma_kernel, p));
}

}  // namespace b2
