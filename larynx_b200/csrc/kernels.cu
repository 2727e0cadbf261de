// kernels.cu — every non-conv kernel of the hot path + the conv launcher.
#include <string.h>

#include "conv1d.cuh"
#include <mutex>

#include "kernels.cuh"

namespace b2 {

// ------------------------------------------------------------------------------------------------
// conv launcher
// ------------------------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN>
static void launch_cfg(const ConvParams& p, int batch, int Lmax, cudaStream_t s) {
  using C = ConvCfg<BM, BN, TM, TN>;
  dim3 grid((Lmax + BN - 1) / BN, (p.Cout + BM - 1) / BM, batch * p.n_phase);
  conv1d_kernel<BM, BN, TM, TN><<<grid, C::NT, C::smem_bytes(p.K), s>>>(p);
}

template <int BM, int BN, int TM, int TN>
static void set_attr() {
  B2_CUDA(cudaFuncSetAttribute(conv1d_kernel<BM, BN, TM, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)ConvCfg<BM, BN, TM, TN>::smem_bytes(16)));
}

void conv1d_init_attributes() {
  set_attr<64, 128, 8, 8>();
  set_attr<64, 64, 8, 4>();
  set_attr<32, 32, 4, 4>();
}

int launch_conv1d(const ConvParams& p, int batch, int Lmax, cudaStream_t stream) {
  B2_CHECK(p.Cout % 4 == 0, "conv1d: Cout must be a multiple of 4");
  B2_CHECK((p.K - 1) * p.dil <= CONV_HALO_MAX, "conv1d: receptive field too wide");
  B2_CHECK(p.K <= 16, "conv1d: K too large");
  B2_CHECK(p.n_phase == 1 || p.K == 2, "conv1d: transposed conv needs 2 polyphase taps");
  if (Lmax <= 0 || batch <= 0) return 0;
  auto ctas = [&](int bm, int bn) {
    return (long long)((Lmax + bn - 1) / bn) * ((p.Cout + bm - 1) / bm) * batch * p.n_phase;
  };
  const bool paired = p.epi == EPI_COUPLING;  // needs TM % 4 == 0 (always true here)
  (void)paired;
  if (ctas(64, 128) >= 120 && p.Cout >= 64) launch_cfg<64, 128, 8, 8>(p, batch, Lmax, stream);
  else if (ctas(64, 64) >= 120 && p.Cout >= 64) launch_cfg<64, 64, 8, 4>(p, batch, Lmax, stream);
  else launch_cfg<32, 32, 4, 4>(p, batch, Lmax, stream);
  B2_CUDA(cudaGetLastError());
  return 1;
}

// ------------------------------------------------------------------------------------------------
// embedding: x[b][c][t] = emb[ids[b][t]][c] * sqrt(H)           (glow_tts/models.py:119-120)
// ------------------------------------------------------------------------------------------------
__global__ void embed_kernel(const long long* __restrict__ ids, const int* __restrict__ lens, int max_tx,
                             const float* __restrict__ emb, int n_vocab, int H, float scale, float* __restrict__ x,
                             long long x_bs, int x_ld, int* __restrict__ bad_id) {
  const int b = blockIdx.z;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (t >= lens[b]) return;
  long long id = ids[(size_t)b * max_tx + t];
  if (id < 0 || id >= n_vocab) {  // nn.Embedding raises IndexError (glow_tts/models.py:119): flag it, the host throws
    if (c == 0) atomicExch(bad_id, 1);
    id = 0;
  }
  x[(size_t)b * x_bs + (size_t)c * x_ld + t] = emb[(size_t)id * H + c] * scale;
}

void launch_embed(const long long* ids, const int* lens, int batch, int max_tx, const float* emb, int n_vocab, int H,
                  float* x, long long x_bs, int x_ld, int* bad_id, cudaStream_t s) {
  dim3 grid((max_tx + 63) / 64, H, batch);
  embed_kernel<<<grid, 64, 0, s>>>(ids, lens, max_tx, emb, n_vocab, H, sqrtf((float)H), x, x_bs, x_ld, bad_id);
  B2_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// channel LayerNorm  (glow_tts/layers.py:19-28): per (b,t) over channels, biased variance, eps 1e-4
//   v = x (+ res); v = relu(v) if pre_relu; y = (v-mean)*rsqrt(var+eps)*gamma+beta; relu if post_relu
// block = 32 time steps x 8 channel groups; each thread keeps its C/8 values in registers.
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAXC = 256;
constexpr int LN_PADL = 40;  // == UM_PADL (conv_umma.cuh)

__device__ __forceinline__ unsigned int ln_cvt_bf16x2(float a, float b) {
  unsigned int r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// Each thread owns whole 8-channel groups (g = ty + 8*i) of one time step, so that besides the fp32 result it can
// emit the bf16 split planes [plane][C/8][PADL + rows][8] that the tensor-core convs load by bulk copy.
// Grid x covers `cover` >= Lmax time steps: blocks past the valid length only write the zero rows of the image.
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* __restrict__ y,
                                                        long long bs, int ld, int C, const int* __restrict__ lens,
                                                        int pre_relu, int post_relu, float eps, unsigned char* planes,
                                                        long long pl_bs, int pl_kc, int pl_rows, int pl_nsplit,
                                                        int cover) {
  __shared__ float red[8][33];
  const int b = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int t = blockIdx.x * 32 + tx;
  const int len = lens[b];
  const bool valid = t < len;
  const float* xb = x + (size_t)b * bs;
  const float* rb = res ? res + (size_t)b * bs : nullptr;
  const int ngrp = C >> 3;
  float v[LN_MAXC / 64][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXC / 64; ++i) {
    const int g = ty + i * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = g * 8 + e;
      const bool ok = valid && g < ngrp;  // branch-free: all loads of the column are in flight together
      const size_t off = ok ? (size_t)c * ld + t : 0;
      float a = xb[off];
      if (rb) a += rb[off];
      if (pre_relu) a = fmaxf(a, 0.f);
      a = ok ? a : 0.f;
      v[i][e] = a;
      s += a;
    }
  }
  red[ty][tx] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) tot += red[j][tx];
  const float mean = tot / (float)C;
  __syncthreads();
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXC / 64; ++i) {
    if (ty + i * 8 < ngrp) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[i][e] - mean;
        q += d * d;
      }
    }
  }
  red[ty][tx] = q;
  __syncthreads();
  float qt = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) qt += red[j][tx];
  const float rstd = rsqrtf(qt / (float)C + eps);
  float* yb = y + (size_t)b * bs;
  unsigned char* pb = planes ? planes + (size_t)b * pl_bs : nullptr;
  const size_t pstride = (size_t)pl_kc * pl_rows * 16;
#pragma unroll
  for (int i = 0; i < LN_MAXC / 64; ++i) {
    const int g = ty + i * 8;
    if (g >= ngrp) continue;
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = g * 8 + e;
      float w = (v[i][e] - mean) * rstd * gamma[c] + beta[c];
      if (post_relu) w = fmaxf(w, 0.f);
      o[e] = valid ? w : 0.f;
      if (valid) yb[(size_t)c * ld + t] = w;
    }
    if (pb && t < cover) {
      unsigned char* base = pb + ((size_t)g * pl_rows + LN_PADL + t) * 16;
      unsigned int h[4], l[4], l2[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a0 = o[2 * e], a1 = o[2 * e + 1];
        h[e] = ln_cvt_bf16x2(a0, a1);
        const float r0 = a0 - __uint_as_float(h[e] << 16), r1 = a1 - __uint_as_float(h[e] & 0xffff0000u);
        l[e] = ln_cvt_bf16x2(r0, r1);
        l2[e] = ln_cvt_bf16x2(r0 - __uint_as_float(l[e] << 16), r1 - __uint_as_float(l[e] & 0xffff0000u));
      }
      *reinterpret_cast<uint4*>(base) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(base + pstride) = make_uint4(l[0], l[1], l[2], l[3]);
      if (pl_nsplit == 3) *reinterpret_cast<uint4*>(base + 2 * pstride) = make_uint4(l2[0], l2[1], l2[2], l2[3]);
      if (blockIdx.x == 0) {  // zero rows in front of the image: t in [-PADL, 0)
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int r = tx; r < LN_PADL; r += 32)
          for (int pl = 0; pl < pl_nsplit; ++pl) *reinterpret_cast<uint4*>(pb + pl * pstride + ((size_t)g * pl_rows + r) * 16) = z;
      }
    }
  }
}

void launch_layernorm(const float* x, const float* res, const float* gamma, const float* beta, float* y, long long bs,
                      int ld, int C, const int* lens, int batch, int Lmax, int pre_relu, int post_relu, cudaStream_t s,
                      void* planes, long long pl_bs, int pl_kc, int pl_rows, int pl_nsplit) {
  B2_CHECK(C <= LN_MAXC && C % 8 == 0, "layernorm: channel count must be a multiple of 8, at most 256");
  int cover = Lmax;
  if (planes) {
    cover = (Lmax + 255) / 256 * 256 + 128;  // rows the consumers may read behind the data must be zero
    B2_CHECK(LN_PADL + cover <= pl_rows, "layernorm: plane image too short");
  }
  dim3 grid((cover + 31) / 32, batch);
  layernorm_kernel<<<grid, 256, 0, s>>>(x, res, gamma, beta, y, bs, ld, C, lens, pre_relu, post_relu, 1e-4f,
                                        (unsigned char*)planes, pl_bs, pl_kc, pl_rows, pl_nsplit, cover);
  B2_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// relative-position multi-head self-attention (glow_tts/attentions.py:214-264), closed form:
//   s[i][j] = (q_i.k_j + [|j-i|<=w] q_i.Ek[j-i+w]) / sqrt(dk);  p = softmax_j(s) over j < len
//   o_i = sum_j p[i][j] v_j + sum_{|r|<=w} p[i][i+r] Ev[r+w]
// (keys j >= len get -1e4 in the reference, i.e. exp underflows to exactly 0 in fp32 for valid rows)
// qkv: [b][3*H][ld]: rows [0,H) = q, [H,2H) = k, [2H,3H) = v; head h owns channels h*dk..(h+1)*dk.
// grid (ceil(T/QT), heads, batch); block 128.  Scores for QT queries x T keys live in smem.
// ------------------------------------------------------------------------------------------------
constexpr int ATT_QT = 8;
constexpr int ATT_JT = 32;
constexpr int ATT_DK_MAX = 96;
constexpr int ATT_NT = 128;

__global__ void __launch_bounds__(ATT_NT) rel_attention_kernel(const float* __restrict__ qkv, long long bs, int ld,
                                                                int H, int dk, const float* __restrict__ emb_k,
                                                                const float* __restrict__ emb_v, int window,
                                                                const int* __restrict__ lens, float* __restrict__ out,
                                                                long long out_bs, int out_ld, int Tpad) {
  extern __shared__ float sm[];
  float* qs = sm;                         // [dk][QT]
  float* kv = qs + ATT_DK_MAX * ATT_QT;   // [dk][JT+1]
  float* sc = kv + ATT_DK_MAX * (ATT_JT + 1);  // [QT][Tpad]
  const int b = blockIdx.z, h = blockIdx.y;
  const int i0 = blockIdx.x * ATT_QT;
  const int len = lens[b];
  if (i0 >= len) return;
  const int tid = threadIdx.x;
  const float* qb = qkv + (size_t)b * bs + (size_t)(h * dk) * ld;
  const float* kb = qb + (size_t)H * ld;
  const float* vb = kb + (size_t)H * ld;
  const float inv_sqrt = 1.f / sqrtf((float)dk);
  const int nrel = 2 * window + 1;

  for (int e = tid; e < dk * ATT_QT; e += ATT_NT) {
    const int d = e / ATT_QT, i = e % ATT_QT;
    qs[d * ATT_QT + i] = (i0 + i) < len ? qb[(size_t)d * ld + i0 + i] : 0.f;
  }
  __syncthreads();

  // ---- scores: tile over keys
  for (int j0 = 0; j0 < len; j0 += ATT_JT) {
    for (int e = tid; e < dk * ATT_JT; e += ATT_NT) {
      const int d = e / ATT_JT, j = e % ATT_JT;
      kv[d * (ATT_JT + 1) + j] = (j0 + j) < len ? kb[(size_t)d * ld + j0 + j] : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < ATT_QT * ATT_JT; e += ATT_NT) {
      const int i = e / ATT_JT, j = e % ATT_JT;
      if (j0 + j < len) {
        float a = 0.f;
        for (int d = 0; d < dk; ++d) a = fmaf(qs[d * ATT_QT + i], kv[d * (ATT_JT + 1) + j], a);
        sc[i * Tpad + j0 + j] = a * inv_sqrt;
      }
    }
    __syncthreads();
  }
  // ---- relative key logits on the band
  for (int e = tid; e < ATT_QT * nrel; e += ATT_NT) {
    const int i = e / nrel, r = e % nrel;
    const int j = i0 + i + r - window;
    if ((i0 + i) < len && j >= 0 && j < len) {
      float a = 0.f;
      for (int d = 0; d < dk; ++d) a = fmaf(qs[d * ATT_QT + i], emb_k[r * dk + d], a);
      sc[i * Tpad + j] += a * inv_sqrt;
    }
  }
  __syncthreads();
  // ---- softmax: one warp per query row
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = warp; i < ATT_QT; i += ATT_NT / 32) {
      if (i0 + i >= len) continue;
      float* row = sc + i * Tpad;
      float mx = -INFINITY;
      for (int j = lane; j < len; j += 32) mx = fmaxf(mx, row[j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
      for (int j = lane; j < len; j += 32) {
        const float ev = expf(row[j] - mx);
        row[j] = ev;
        sum += ev;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.f / sum;
      for (int j = lane; j < len; j += 32) row[j] *= inv;
    }
  }
  __syncthreads();
  // ---- output: thread owns (d, i) pairs; accumulate over key tiles
  constexpr int PER = (ATT_DK_MAX * ATT_QT + ATT_NT - 1) / ATT_NT;
  float oacc[PER];
#pragma unroll
  for (int u = 0; u < PER; ++u) oacc[u] = 0.f;
  for (int j0 = 0; j0 < len; j0 += ATT_JT) {
    for (int e = tid; e < dk * ATT_JT; e += ATT_NT) {
      const int d = e / ATT_JT, j = e % ATT_JT;
      kv[d * (ATT_JT + 1) + j] = (j0 + j) < len ? vb[(size_t)d * ld + j0 + j] : 0.f;
    }
    __syncthreads();
    const int jn = min(ATT_JT, len - j0);
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int e = tid + u * ATT_NT;
      if (e < dk * ATT_QT) {
        const int d = e / ATT_QT, i = e % ATT_QT;
        float a = oacc[u];
        for (int j = 0; j < jn; ++j) a = fmaf(sc[i * Tpad + j0 + j], kv[d * (ATT_JT + 1) + j], a);
        oacc[u] = a;
      }
    }
    __syncthreads();
  }
  float* ob = out + (size_t)b * out_bs + (size_t)(h * dk) * out_ld;
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int e = tid + u * ATT_NT;
    if (e < dk * ATT_QT) {
      const int d = e / ATT_QT, i = e % ATT_QT;
      const int qi = i0 + i;
      if (qi < len) {
        float a = oacc[u];
        for (int r = 0; r < nrel; ++r) {
          const int j = qi + r - window;
          if (j >= 0 && j < len) a = fmaf(sc[i * Tpad + j], emb_v[r * dk + d], a);
        }
        ob[(size_t)d * out_ld + qi] = a;
      }
    }
  }
}

void attention_init_attributes() {
  B2_CUDA(cudaFuncSetAttribute(rel_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}

void launch_rel_attention(const float* qkv, long long bs, int ld, int H, int n_heads, const float* emb_k,
                          const float* emb_v, int window, const int* lens, int batch, int Tmax, float* out,
                          long long out_bs, int out_ld, cudaStream_t s) {
  const int dk = H / n_heads;
  B2_CHECK(dk <= ATT_DK_MAX, "attention: head dim too large");
  const int Tpad = ((Tmax + 31) / 32) * 32 + 1;
  const size_t smem = sizeof(float) * (ATT_DK_MAX * ATT_QT + ATT_DK_MAX * (ATT_JT + 1) + (size_t)ATT_QT * Tpad);
  B2_CHECK(smem <= 200 * 1024, "attention: sequence too long for the shared-memory score tile");
  dim3 grid((Tmax + ATT_QT - 1) / ATT_QT, n_heads, batch);
  rel_attention_kernel<<<grid, ATT_NT, smem, s>>>(qkv, bs, ld, H, dk, emb_k, emb_v, window, lens, out, out_bs,
                                                   out_ld, Tpad);
  B2_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// duration head tail + lengths + alignment   (glow_tts/models.py:49,323-331,356-363; utils.py:99-115)
//   logw[t] = bias + sum_c w[c]*h[c][t]   (1x1 conv 256->1, masked)
//   w = exp(logw)*length_scale; w_ceil = ceil(w); cum = inclusive scan; y_len = max(sum,1)//n_sqz*n_sqz
//   idx[f] = phoneme whose span [cum[i-1], cum[i]) contains frame f  (f < y_len), else -1
// one block per batch item.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) duration_generic_kernel(const float* __restrict__ h, long long h_bs, int h_ld, int C,
                                                       const float* __restrict__ pw, const float* __restrict__ pb,
                                                       const int* __restrict__ lens, int max_tx, float length_scale,
                                                       int n_sqz, float* __restrict__ logw, int* __restrict__ w_ceil,
                                                       int* __restrict__ cum, int* __restrict__ y_len) {
  __shared__ int part[256];
  const int b = blockIdx.x;
  const int len = lens[b];
  const int tid = threadIdx.x;
  const float* hb = h + (size_t)b * h_bs;
  // each thread owns a contiguous run of phonemes so the scan is a two-level sum
  const int per = (max_tx + 255) / 256;
  const int tbeg = tid * per;
  int local = 0;
  for (int u = 0; u < per; ++u) {
    const int t = tbeg + u;
    if (t >= max_tx) break;
    int wc = 0;
    float lw = 0.f;
    if (t < len) {
      // the fma chain keeps the channel order (bit-identical results); the loads of 32 channels are issued together
      // (one dependent L2 round trip per channel made this 1-CTA kernel 53 us long)
      float a = 0.f;
      int c0 = 0;
      for (; c0 + 32 <= C; c0 += 32) {
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __ldcg(hb + (size_t)(c0 + j) * h_ld + t);
#pragma unroll
        for (int j = 0; j < 32; ++j) a = fmaf(pw[c0 + j], x[j], a);
      }
      for (int c = c0; c < C; ++c) a = fmaf(pw[c], hb[(size_t)c * h_ld + t], a);
      lw = a + pb[0];
      const float w = expf(lw) * length_scale;
      wc = (int)ceilf(w);
    }
    logw[(size_t)b * max_tx + t] = lw;
    w_ceil[(size_t)b * max_tx + t] = wc;
    local += wc;
  }
  part[tid] = local;
  __syncthreads();
  // exclusive prefix over threads (256 entries: serial per thread is fine, tiny)
  int before = 0;
  for (int j = 0; j < tid; ++j) before += part[j];
  int run = before;
  for (int u = 0; u < per; ++u) {
    const int t = tbeg + u;
    if (t >= max_tx) break;
    run += w_ceil[(size_t)b * max_tx + t];
    cum[(size_t)b * max_tx + t] = run;
  }
  if (tid == 255) {
    int total = run;
    if (total < 1) total = 1;
    y_len[b] = (total / n_sqz) * n_sqz;
  }
}

// C % 4 == 0, C <= 256 (every reference voice: filter_channels_dp = 256).  Four threads per phoneme load a quarter of the
// channels each — all loads of the block are in flight together, one L2 round trip instead of one per 32 channels — and
// then hand the fma chain from quarter to quarter, so the sum keeps the plain channel order of the generic kernel
// (bit-identical logw).  The scan is a warp scan per 32 phonemes + warp totals.
constexpr int DUR_PH = 128;  // phonemes per pass
__global__ void __launch_bounds__(DUR_PH * 4) duration_kernel(const float* __restrict__ h, long long h_bs, int h_ld, int C,
                                                              const float* __restrict__ pw, const float* __restrict__ pb,
                                                              const int* __restrict__ lens, int max_tx, float length_scale,
                                                              int n_sqz, float* __restrict__ logw, int* __restrict__ w_ceil,
                                                              int* __restrict__ cum, int* __restrict__ y_len) {
  __shared__ int s_wc[DUR_PH];
  __shared__ int s_wtot[DUR_PH / 32];
  __shared__ int s_carry;
  const int b = blockIdx.x, len = lens[b];
  const int tid = threadIdx.x, lane = tid & 31, tl = tid >> 2, q = tid & 3;
  const float* hb = h + (size_t)b * h_bs;
  const int cq = C >> 2;
  if (tid == 0) s_carry = 0;
  for (int t0 = 0; t0 < max_tx; t0 += DUR_PH) {
    const int t = t0 + tl;
    const bool on = t < max_tx && t < len;
    float x[64];
    if (on) {
#pragma unroll
      for (int j = 0; j < 64; ++j)
        if (j < cq) x[j] = __ldcg(hb + (size_t)(q * cq + j) * h_ld + t);
    }
    float a = 0.f;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      if (on && q == qq) {
#pragma unroll
        for (int j = 0; j < 64; ++j)
          if (j < cq) a = fmaf(pw[qq * cq + j], x[j], a);
      }
      a = __shfl_sync(0xffffffffu, a, (lane & ~3) | qq);  // the chain value after quarter qq, to the whole quad
    }
    if (q == 0) {
      int wc = 0;
      float lw = 0.f;
      if (on) {
        lw = a + pb[0];
        wc = (int)ceilf(expf(lw) * length_scale);
      }
      if (t < max_tx) {
        logw[(size_t)b * max_tx + t] = lw;
        w_ceil[(size_t)b * max_tx + t] = wc;
      }
      s_wc[tl] = wc;
    }
    __syncthreads();
    int incl = 0;
    if (tid < DUR_PH) {
      incl = s_wc[tid];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
      }
      if (lane == 31) s_wtot[tid >> 5] = incl;
    }
    __syncthreads();
    if (tid < DUR_PH) {
      int off = s_carry;
      for (int w = 0; w < (tid >> 5); ++w) off += s_wtot[w];
      incl += off;
      if (t0 + tid < max_tx) cum[(size_t)b * max_tx + t0 + tid] = incl;
    }
    __syncthreads();
    if (tid == DUR_PH - 1) s_carry = incl;
  }
  __syncthreads();
  if (tid == 0) {
    int total = s_carry;
    if (total < 1) total = 1;
    y_len[b] = (total / n_sqz) * n_sqz;
  }
}

void launch_duration(const float* h, long long h_bs, int h_ld, int C, const float* pw, const float* pb,
                     const int* lens, int batch, int max_tx, float length_scale, int n_sqz, float* logw, int* w_ceil,
                     int* cum, int* y_len, cudaStream_t s) {
  if (C % 4 == 0 && C <= 256)
    duration_kernel<<<batch, DUR_PH * 4, 0, s>>>(h, h_bs, h_ld, C, pw, pb, lens, max_tx, length_scale, n_sqz, logw, w_ceil, cum,
                                                 y_len);
  else
    duration_generic_kernel<<<batch, 256, 0, s>>>(h, h_bs, h_ld, C, pw, pb, lens, max_tx, length_scale, n_sqz, logw, w_ceil,
                                                  cum, y_len);
  B2_CUDA(cudaGetLastError());
}

// frame -> phoneme index by binary search over cum (first i with cum[i] > f)
__global__ void align_index_kernel(const int* __restrict__ cum, const int* __restrict__ lens,
                                   const int* __restrict__ y_len, int max_tx, int* __restrict__ idx, int idx_ld,
                                   int* __restrict__ sq_len, int n_sqz) {
  const int b = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f == 0) sq_len[b] = y_len[b] / n_sqz;
  if (f >= idx_ld) return;
  int r = -1;
  if (f < y_len[b]) {
    const int* cb = cum + (size_t)b * max_tx;
    int lo = 0, hi = lens[b] - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cb[mid] > f) hi = mid;
      else lo = mid + 1;
    }
    r = lo;
  }
  idx[(size_t)b * idx_ld + f] = r;
}

void launch_align_index(const int* cum, const int* lens, const int* y_len, int batch, int max_tx, int* idx, int idx_ld,
                        int* sq_len, int n_sqz, cudaStream_t s) {
  if (idx_ld <= 0) idx_ld = 1;
  dim3 grid((idx_ld + 127) / 128, batch);
  align_index_kernel<<<grid, 128, 0, s>>>(cum, lens, y_len, max_tx, idx, idx_ld, sq_len, n_sqz);
  B2_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller (production noise when no eps is injected)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                             uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned long long ctr) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x9E3779B9u, c3 = 0xBB67AE85u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const float u1 = ((float)(c0 >> 8) + 0.5f) * (1.f / 16777216.f);
  const float u2 = ((float)(c1 >> 8) + 0.5f) * (1.f / 16777216.f);
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

// ------------------------------------------------------------------------------------------------
// expand + noise + squeeze   (glow_tts/models.py:336-348; utils.py:135-147)
//   z[c][f] = x_m[c][idx[f]] + eps[c][f]*noise_scale   (f < y_len), written squeezed:
//   zs[(f % n_sqz)*mel + c][f / n_sqz]
// ------------------------------------------------------------------------------------------------
__global__ void expand_squeeze_kernel(const float* __restrict__ x_m, long long xm_bs, int xm_ld,
                                      const int* __restrict__ idx, int idx_ld, const int* __restrict__ y_len,
                                      const float* __restrict__ eps, long long eps_bs, int eps_ld, float noise_scale,
                                      unsigned long long seed, int mel, int n_sqz, float* __restrict__ zs,
                                      long long zs_bs, int zs_ld) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= y_len[b]) return;
  const int i = idx[(size_t)b * idx_ld + f];
  float e;
  if (eps) e = eps[(size_t)b * eps_bs + (size_t)c * eps_ld + f];
  else e = philox_normal(seed, ((unsigned long long)b << 40) ^ ((unsigned long long)c << 24) ^ (unsigned long long)f);
  const float v = x_m[(size_t)b * xm_bs + (size_t)c * xm_ld + i] + e * noise_scale;
  zs[(size_t)b * zs_bs + (size_t)((f % n_sqz) * mel + c) * zs_ld + f / n_sqz] = v;
}

void launch_expand_squeeze(const float* x_m, long long xm_bs, int xm_ld, const int* idx, int idx_ld, const int* y_len,
                           const float* eps, long long eps_bs, int eps_ld, float noise_scale, unsigned long long seed,
                           int mel, int n_sqz, float* zs, long long zs_bs, int zs_ld, int batch, int Tymax,
                           cudaStream_t s) {
  if (Tymax <= 0) return;
  dim3 grid((Tymax + 127) / 128, mel, batch);
  expand_squeeze_kernel<<<grid, 128, 0, s>>>(x_m, xm_bs, xm_ld, idx, idx_ld, y_len, eps, eps_bs, eps_ld, noise_scale,
                                             seed, mel, n_sqz, zs, zs_bs, zs_ld);
  B2_CUDA(cudaGetLastError());
}

// the production noise stream on its own: eps[b][c][f] exactly as expand_squeeze_kernel draws it (moment tests)
__global__ void noise_fill_kernel(unsigned long long seed, int mel, int n, float* __restrict__ out, long long out_bs, int out_ld) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  out[(size_t)b * out_bs + (size_t)c * out_ld + f] =
      philox_normal(seed, ((unsigned long long)b << 40) ^ ((unsigned long long)c << 24) ^ (unsigned long long)f);
}
void launch_noise_fill(unsigned long long seed, int batch, int mel, int n, float* out, long long out_bs, int out_ld, cudaStream_t s) {
  if (n <= 0 || batch <= 0) return;
  dim3 grid((n + 127) / 128, mel, batch);
  noise_fill_kernel<<<grid, 128, 0, s>>>(seed, mel, n, out, out_bs, out_ld);
  B2_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// unsqueeze (+ optional mel glue)   (glow_tts/utils.py:150-160; larynx/__init__.py:242-249)
//   mel[c][f] = zs[(f % n_sqz)*mel + c][f / n_sqz] for f < y_len else 0
//   glued = log(max(10^((denorm(mel))/gain), 1e-5))   (larynx/audio.py:83-108)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float glue_value(float m, const b200tts_glue g) {
  if (g.signal_norm) {
    if (g.symmetric_norm) {
      if (g.clip_norm) m = fminf(fmaxf(m, -g.max_norm), g.max_norm);
      m = ((m + g.max_norm) * -g.min_level_db / (2.f * g.max_norm)) + g.min_level_db;
    } else {
      if (g.clip_norm) m = fminf(fmaxf(m, 0.f), g.max_norm);
      m = (m * -g.min_level_db / g.max_norm) + g.min_level_db;
    }
    m += g.ref_level_db;
  }
  if (g.convert_db_to_amp) m = powf(10.f, m / g.spec_gain);
  if (g.do_dynamic_range_compression) m = logf(fmaxf(m, 1e-5f));
  return m;
}

__global__ void unsqueeze_kernel(const float* __restrict__ zs, long long zs_bs, int zs_ld,
                                 const int* __restrict__ y_len, int mel, int n_sqz, float* __restrict__ out,
                                 long long out_bs, int out_ld, float* __restrict__ glued, long long gl_bs, int gl_ld,
                                 const b200tts_glue g, int width) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= width) return;
  float v = 0.f;
  const bool valid = f < y_len[b];
  if (valid) v = zs[(size_t)b * zs_bs + (size_t)((f % n_sqz) * mel + c) * zs_ld + f / n_sqz];
  if (out && f < out_ld) out[(size_t)b * out_bs + (size_t)c * out_ld + f] = v;
  if (glued && valid && f < gl_ld) glued[(size_t)b * gl_bs + (size_t)c * gl_ld + f] = glue_value(v, g);
}

void launch_unsqueeze(const float* zs, long long zs_bs, int zs_ld, const int* y_len, int mel, int n_sqz, float* out,
                      long long out_bs, int out_ld, float* glued, long long gl_bs, int gl_ld, const b200tts_glue* g,
                      int batch, int width, cudaStream_t s) {
  if (width <= 0) return;
  b200tts_glue gg;
  if (g) gg = *g;
  else memset(&gg, 0, sizeof(gg));
  dim3 grid((width + 127) / 128, mel, batch);
  unsqueeze_kernel<<<grid, 128, 0, s>>>(zs, zs_bs, zs_ld, y_len, mel, n_sqz, out, out_bs, out_ld, glued, gl_bs, gl_ld,
                                        gg, width);
  B2_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// conv_post: lrelu(0.01) -> Conv1d(C -> 1, k7) -> tanh, + per item |max| (hifi_gan/models.py:198-200)
// input is the MRF mean of up to three buffers.  One thread per output sample.
// ------------------------------------------------------------------------------------------------
constexpr int POST_BT = 256;
constexpr int POST_KMAX = 7;
__global__ void __launch_bounds__(POST_BT) conv_post_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                             const float* __restrict__ x3, long long bs, int ld,
                                                             float in_div, float slope, const float* __restrict__ w,
                                                             const float* __restrict__ bias, int C, int K,
                                                             const int* __restrict__ frames, int len_mul,
                                                             float* __restrict__ pcm, int pcm_ld,
                                                             unsigned int* __restrict__ peak_bits) {
  extern __shared__ float sm[];
  float* xs = sm;                              // [C][POST_BT + K - 1]
  float* wsm = sm + C * (POST_BT + POST_KMAX - 1);  // [C][K]
  const int b = blockIdx.y;
  const int L = frames[b] * len_mul;
  const int t0 = blockIdx.x * POST_BT;
  if (t0 >= L) return;
  const int W = POST_BT + K - 1, padl = (K - 1) / 2;
  const float* a1 = x1 + (size_t)b * bs;
  const float* a2 = x2 ? x2 + (size_t)b * bs : nullptr;
  const float* a3 = x3 ? x3 + (size_t)b * bs : nullptr;
  // one warp per channel row, lanes along time: coalesced, no index division, and the (up to 27) loads of a row are
  // independent so that they are all in flight together
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < C; c += POST_BT / 32) {
    const size_t row = (size_t)c * ld;
    float v[(POST_BT + POST_KMAX - 1 + 31) / 32];
#pragma unroll
    for (int i = 0; i < (POST_BT + POST_KMAX - 1 + 31) / 32; ++i) {
      const int pcol = lane + 32 * i, t = t0 + pcol - padl;
      float a = 0.f;
      if (pcol < W && t >= 0 && t < L) {
        a = a1[row + t];
        if (a2) a += a2[row + t];
        if (a3) a += a3[row + t];
      }
      v[i] = a;
    }
#pragma unroll
    for (int i = 0; i < (POST_BT + POST_KMAX - 1 + 31) / 32; ++i) {
      const int pcol = lane + 32 * i;
      float a = v[i];
      if (in_div != 1.f) a = a / in_div;
      a = a > 0.f ? a : a * slope;
      if (pcol < W) xs[c * W + pcol] = a;
    }
  }
  for (int e = threadIdx.x; e < C * K; e += POST_BT) wsm[e] = w[e];
  __syncthreads();
  const int t = t0 + threadIdx.x;
  float av = 0.f;
  if (t < L) {
    float a = 0.f;
    if (K == POST_KMAX) {  // every reference vocoder: same accumulation order, taps unrolled
      for (int c = 0; c < C; ++c) {
        const float* xr = xs + c * W + threadIdx.x;
        const float* wr = wsm + c * POST_KMAX;
#pragma unroll
        for (int k = 0; k < POST_KMAX; ++k) a = fmaf(wr[k], xr[k], a);
      }
    } else {
      for (int c = 0; c < C; ++c)
        for (int k = 0; k < K; ++k) a = fmaf(wsm[c * K + k], xs[c * W + threadIdx.x + k], a);
    }
    const float o = tanhf(a + bias[0]);
    pcm[(size_t)b * pcm_ld + t] = o;
    av = fabsf(o);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) av = fmaxf(av, __shfl_xor_sync(0xffffffffu, av, o));
  if ((threadIdx.x & 31) == 0) atomicMax(peak_bits + b, __float_as_uint(av));
}

void launch_conv_post(const float* x1, const float* x2, const float* x3, long long bs, int ld, float in_div,
                      float slope, const float* w, const float* bias, int C, int K, const int* frames, int len_mul,
                      int batch, int Lmax, float* pcm, int pcm_ld, unsigned int* peak_bits, cudaStream_t s) {
  B2_CHECK(K <= POST_KMAX, "conv_post: kernel too wide");
  if (Lmax <= 0) return;
  const size_t smem = sizeof(float) * ((size_t)C * (POST_BT + POST_KMAX - 1) + (size_t)C * K);
  B2_CHECK(smem <= 48 * 1024, "conv_post: too many channels");
  dim3 grid((Lmax + POST_BT - 1) / POST_BT, batch);
  conv_post_kernel<<<grid, POST_BT, smem, s>>>(x1, x2, x3, bs, ld, in_div, slope, w, bias, C, K, frames, len_mul, pcm,
                                               pcm_ld, peak_bits);
  B2_CUDA(cudaGetLastError());
}

// audio_float_to_int16 (larynx/audio.py:118-125): a * (32767 / max(0.01, max|a|)), clip, truncate
__global__ void pcm_to_i16_kernel(const float* __restrict__ pcm, int pcm_ld, const unsigned int* __restrict__ peak_bits,
                                  const int* __restrict__ frames, int len_mul, short* __restrict__ out, int out_ld) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= out_ld) return;
  short r = 0;
  if (t < frames[b] * len_mul) {
    const float peak = __uint_as_float(peak_bits[b]);
    const float scale = 32767.0f / fmaxf(0.01f, peak);
    float v = pcm[(size_t)b * pcm_ld + t] * scale;
    v = fminf(fmaxf(v, -32767.0f), 32767.0f);
    r = (short)(int)v;  // truncation toward zero like ndarray.astype("int16")
  }
  out[(size_t)b * out_ld + t] = r;
}

void launch_pcm_to_i16(const float* pcm, int pcm_ld, const unsigned int* peak_bits, const int* frames, int len_mul,
                       int batch, short* out, int out_ld, cudaStream_t s) {
  if (out_ld <= 0) return;
  dim3 grid((out_ld + 255) / 256, batch);
  pcm_to_i16_kernel<<<grid, 256, 0, s>>>(pcm, pcm_ld, peak_bits, frames, len_mul, out, out_ld);
  B2_CUDA(cudaGetLastError());
}

// y[b][c][t] = 0 for t in [lens[b]*len_mul, width)  — keeps padded tails of caller-visible outputs at 0
__global__ void zero_tail_kernel(float* __restrict__ y, long long bs, int ld, int C, const int* __restrict__ lens,
                                 int len_mul, int width) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < width && t >= lens[b] * len_mul) y[(size_t)b * bs + (size_t)c * ld + t] = 0.f;
}
void launch_zero_tail(float* y, long long bs, int ld, int C, const int* lens, int len_mul, int batch, int width,
                      cudaStream_t s) {
  if (width <= 0) return;
  dim3 grid((width + 255) / 256, C, batch);
  zero_tail_kernel<<<grid, 256, 0, s>>>(y, bs, ld, C, lens, len_mul, width);
  B2_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// Spectral-subtraction denoiser (larynx/hifi_gan.py:171-179) on the device.
//   transform (larynx/audio.py:291-306):  frames i = 0, 256, ... < N - 1024; X_i = rfft(hanning(1024) * x[i : i + 1024])
//   denoise:                               |X| <- max(|X| - strength * bias, 0), phase kept
//   inverse (larynx/audio.py:251-288):    out[i : i + 1024] += hanning(1024) * irfft(X_i); no window-power normalisation
// One CTA per frame: 1024-point complex FFT (radix-4 Stockham, 5 passes, one butterfly per thread per pass) in shared
// memory, the per-bin gain max(1 - strength * bias / |X|, 0) applied to bin k and its mirror, and the inverse transform
// computed as Re(FFT(conj X)) / 1024.  The windowed frames go to a scratch buffer; a second kernel overlap-adds the
// (up to four) frames that cover each output sample in the reference's order and tracks the peak for the int16
// conversion — no atomics on the samples, so results are reproducible run to run.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int stft_n_frames(int L) { return L > DN_FFT ? (L - DN_FFT + DN_HOP - 1) / DN_HOP : 0; }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// forward DFT of a[0..1024); b is scratch; returns the buffer holding the result.  tw[t] = exp(-2 pi i t / 1024).
__device__ float2* fft1024(float2* a, float2* b, const float2* __restrict__ tw) {
  const int j = threadIdx.x;
#pragma unroll
  for (int Ns = 1; Ns < DN_FFT; Ns *= 4) {
    const int k = j & (Ns - 1);
    const int ts = k * (DN_FFT / 4 / Ns);
    const float2 v0 = a[j];
    const float2 v1 = cmul(a[j + 256], tw[ts]);
    const float2 v2 = cmul(a[j + 512], tw[2 * ts]);
    const float2 v3 = cmul(a[j + 768], tw[3 * ts]);
    const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
    const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
    const int d = ((j - k) << 2) + k;
    b[d] = make_float2(s02.x + s13.x, s02.y + s13.y);
    b[d + Ns] = make_float2(d02.x + d13.y, d02.y - d13.x);      // d02 - i d13
    b[d + 2 * Ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
    b[d + 3 * Ns] = make_float2(d02.x - d13.y, d02.y + d13.x);  // d02 + i d13
    __syncthreads();
    float2* t = a;
    a = b;
    b = t;
  }
  return a;
}

// grid (max frames, batch), 256 threads.  mag_out != nullptr: write |X| of frame 0 of item 0 (the bias spectrum,
// larynx/hifi_gan.py:199-203) instead of denoising.
__global__ void __launch_bounds__(256) stft_denoise_kernel(const float* __restrict__ pcm, int pcm_ld,
                                                           const int* __restrict__ frames, int len_mul,
                                                           const float* __restrict__ hann, const float2* __restrict__ tw,
                                                           const float* __restrict__ bias, float strength,
                                                           float* __restrict__ fr, long long fr_bs,
                                                           float* __restrict__ mag_out) {
  __shared__ float2 A[DN_FFT], B[DN_FFT];
  const int b = blockIdx.y, i = blockIdx.x, tid = threadIdx.x;
  if (i >= stft_n_frames(frames[b] * len_mul)) return;
  const float* x = pcm + (size_t)b * pcm_ld + (size_t)i * DN_HOP;
#pragma unroll
  for (int e = tid; e < DN_FFT; e += 256) A[e] = make_float2(hann[e] * x[e], 0.f);
  __syncthreads();
  float2* X = fft1024(A, B, tw);
  float2* Y = X == A ? B : A;
  for (int k = tid; k <= DN_FFT / 2; k += 256) {
    const float2 v = X[k];
    const float m = sqrtf(v.x * v.x + v.y * v.y);
    if (mag_out) {
      mag_out[k] = m;
      continue;
    }
    const float g = m > 0.f ? fmaxf(m - strength * bias[k], 0.f) / m : 0.f;
    const float re = v.x * g, im = v.y * g;
    if (k == 0 || k == DN_FFT / 2) {
      Y[k] = make_float2(re, 0.f);  // irfft ignores the imaginary part of DC and Nyquist
    } else {
      Y[k] = make_float2(re, -im);  // conj(X): the inverse transform is Re(FFT(conj X)) / N
      Y[DN_FFT - k] = make_float2(re, im);
    }
  }
  if (mag_out) return;
  __syncthreads();
  const float2* Z = fft1024(Y, X, tw);
  float* o = fr + (size_t)b * fr_bs + (size_t)i * DN_FFT;
#pragma unroll
  for (int e = tid; e < DN_FFT; e += 256) o[e] = hann[e] * (Z[e].x * (1.0f / DN_FFT));
}

// overlap-add (larynx/audio.py:262-267): out[n] = sum over frames i covering n, i ascending; peak for audio_float_to_int16
__global__ void __launch_bounds__(256) overlap_add_kernel(const float* __restrict__ fr, long long fr_bs,
                                                          const int* __restrict__ frames, int len_mul,
                                                          float* __restrict__ pcm, int pcm_ld,
                                                          unsigned int* __restrict__ peak_bits) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * 256 + threadIdx.x;
  const int nt = stft_n_frames(frames[b] * len_mul);
  const int olen = nt * DN_HOP + DN_FFT;
  float av = 0.f;
  if (n < olen && n < pcm_ld) {
    const int blk = n / DN_HOP;
    const float* f = fr + (size_t)b * fr_bs;
    float v = 0.f;
#pragma unroll
    for (int j = DN_FFT / DN_HOP - 1; j >= 0; --j) {
      const int i = blk - j;
      if (i >= 0 && i < nt) v += f[(size_t)i * DN_FFT + (n - i * DN_HOP)];
    }
    pcm[(size_t)b * pcm_ld + n] = v;
    av = fabsf(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) av = fmaxf(av, __shfl_xor_sync(0xffffffffu, av, o));
  if ((threadIdx.x & 31) == 0 && av > 0.f) atomicMax(peak_bits + b, __float_as_uint(av));
}

void launch_stft_denoise(const float* pcm, int pcm_ld, const int* frames, int len_mul, int batch, int max_frames,
                         const float* hann, const float* tw, const float* bias, float strength, float* fr,
                         long long fr_bs, float* mag_out, cudaStream_t s) {
  if (max_frames <= 0 || batch <= 0) return;
  dim3 grid(max_frames, batch);
  stft_denoise_kernel<<<grid, 256, 0, s>>>(pcm, pcm_ld, frames, len_mul, hann, (const float2*)tw, bias, strength, fr, fr_bs,
                                           mag_out);
  B2_CUDA(cudaGetLastError());
}

void launch_overlap_add(const float* fr, long long fr_bs, const int* frames, int len_mul, int batch, int max_len,
                        float* pcm, int pcm_ld, unsigned int* peak_bits, cudaStream_t s) {
  if (max_len <= 0 || batch <= 0) return;
  dim3 grid((max_len + 255) / 256, batch);
  overlap_add_kernel<<<grid, 256, 0, s>>>(fr, fr_bs, frames, len_mul, pcm, pcm_ld, peak_bits);
  B2_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// device-wide ordering of the kernels whose CTAs wait for each other (see kernels.cuh)
// ------------------------------------------------------------------------------------------------
namespace {
struct SpinGate {
  std::mutex m;
  cudaEvent_t ev[64] = {};
  bool has[64] = {};
};
SpinGate& spin_gate() {
  static SpinGate g;
  return g;
}
}  // namespace

void spin_gate_begin(cudaStream_t s) {
  int dev = 0;
  B2_CUDA(cudaGetDevice(&dev));
  B2_CHECK(dev >= 0 && dev < 64, "device index out of range");
  SpinGate& g = spin_gate();
  g.m.lock();  // released by spin_gate_end: wait + launch + record are one step with respect to the other host threads
  if (g.has[dev]) {
    const cudaError_t e = cudaStreamWaitEvent(s, g.ev[dev], 0);
    if (e != cudaSuccess) {
      g.m.unlock();
      B2_CUDA(e);
    }
  }
}

void spin_gate_end(cudaStream_t s) {
  int dev = 0;
  cudaGetDevice(&dev);
  SpinGate& g = spin_gate();
  cudaError_t e = cudaSuccess;
  if (!g.ev[dev]) e = cudaEventCreateWithFlags(&g.ev[dev], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(g.ev[dev], s);
  if (e == cudaSuccess) g.has[dev] = true;
  g.m.unlock();
  B2_CUDA(e);
}

}  // namespace b2
