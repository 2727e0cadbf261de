// conv1d.cuh — fused fp32 Conv1d / polyphase ConvTranspose1d as a SIMT implicit GEMM.
//
//   y[b][co][t] = epilogue( bias[co] + sum_ci sum_k W[ci][k][co] * pre(x[b][ci][t + k*dil - pad_left]) )
//
// Covers every convolution on the path (reference call sites):
//   glow_tts/layers.py:73-80 (prenet), attentions.py:205-211 (q/k/v/o 1x1), attentions.py:375-383 (FFN),
//   models.py:39-49 (duration predictor), models.py:133 (proj_m), attentions.py:125-127 (start/end),
//   layers.py:146-162 (WN in_layer + res_skip), hifi_gan/models.py:187 (conv_pre),
//   hifi_gan/models.py:91-98 (ResBlock1 convs), hifi_gan/models.py:190 (ConvTranspose1d, as u phase convs).
// Fusions: zero padding by per-item valid length (== every "* x_mask" of the reference), LeakyReLU on
// load, MRF mean of three inputs on load, bias, ReLU/tanh, residual add, tanh*sigmoid gate
// (glow_tts/utils.py:31-38), WN residual/skip split, affine coupling + InvConvNear^-1 + ActNorm^-1.
//
// Tiling: CTA = BM out-channels x BN time steps, thread = TM x TN micro tile with the TN time steps
// interleaved (t = tx + LX*i) so shared-memory reads of the shifted taps are conflict free for any
// dilation.  K loop runs over Cin in chunks of CK with a 2-stage shared-memory ring (cp.async for
// the weight tile, register-staged loads with the pre-op applied for the activation tile).
#pragma once
#include "common.cuh"

namespace b2 {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

constexpr int CONV_CK = 8;        // input channels per smem stage
constexpr int CONV_HALO_MAX = 80; // (K-1)*dil must be <= this (vctk_small: k7, dil 12 -> 72)

template <int BM, int BN, int TM, int TN>
struct ConvCfg {
  static constexpr int LX = BN / TN;             // threads along time
  static constexpr int LY = BM / TM;             // threads along channels
  static constexpr int NT = LX * LY;
  static constexpr int NW = NT / 32;
  static constexpr int XW = BN + CONV_HALO_MAX;  // smem row width
  static constexpr int PF_ROWS = (CONV_CK + NW - 1) / NW;
  static constexpr int PF_COLS = (XW + 31) / 32;
  static size_t smem_bytes(int K) { return sizeof(float) * 2 * (size_t)(CONV_CK * XW + CONV_CK * K * BM); }
};

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(ConvCfg<BM, BN, TM, TN>::NT) conv1d_kernel(const ConvParams p) {
  using C = ConvCfg<BM, BN, TM, TN>;
  constexpr int LX = C::LX, NT = C::NT, NW = C::NW, XW = C::XW, CK = CONV_CK;
  static_assert(TM % 4 == 0, "TM must be a multiple of 4");
  static_assert(NT % 32 == 0, "whole warps");
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                      // [2][CK][XW]
  float* ws = smem + 2 * CK * XW;        // [2][CK][K][BM]

  const int tid = threadIdx.x;
  const int tx = tid % LX, ty = tid / LX;
  const int warp = tid >> 5, lane = tid & 31;
  const int K = p.K, dil = p.dil;
  const int b = blockIdx.z / p.n_phase;
  const int phase = blockIdx.z % p.n_phase;
  const int t0 = blockIdx.x * BN;
  const int co0 = blockIdx.y * BM;
  const int Lb = p.lens ? p.lens[b] * p.len_mul : p.Lin;  // valid input (= output) length
  if (t0 >= Lb) return;

  int pad_left = p.pad_left;
  const float* wbase = p.w;
  if (p.n_phase > 1) {
    // polyphase transposed conv: output n = t*u + r reads x[t + q - 1], x[t + q], q = (r + pad) / u
    const int q = (phase + p.tpad) / p.n_phase;
    pad_left = 1 - q;
    wbase += (size_t)phase * p.Cin * K * p.Cout;
  }
  const int xw_used = BN + (K - 1) * dil;
  const float* xb = p.x + (size_t)b * p.x_bs;
  const float* xb2 = p.x2 ? p.x2 + (size_t)b * p.x_bs : nullptr;
  const float* xb3 = p.x3 ? p.x3 + (size_t)b * p.x_bs : nullptr;
  const int wrow = K * BM;  // floats per input channel in the smem weight tile

  float pf[C::PF_ROWS][C::PF_COLS];

  auto load_x_regs = [&](int ci0) {
#pragma unroll
    for (int r = 0; r < C::PF_ROWS; ++r) {
      const int ci = warp + r * NW;
#pragma unroll
      for (int c = 0; c < C::PF_COLS; ++c) {
        const int pcol = lane + c * 32;
        float v = 0.f;
        const int tin = t0 + pcol - pad_left;
        if (ci < CK && pcol < xw_used && tin >= 0 && tin < Lb && (ci0 + ci) < p.Cin) {
          const size_t off = (size_t)(ci0 + ci) * p.x_ld + tin;
          v = xb[off];
          if (xb2) v += xb2[off];
          if (xb3) v += xb3[off];
          if (p.in_div != 1.f) v = v / p.in_div;
          if (p.pre == PRE_LRELU) v = v > 0.f ? v : v * p.pre_slope;
        }
        pf[r][c] = v;
      }
    }
  };
  auto store_x_regs = [&](int stage) {
    float* dst = xs + stage * CK * XW;
#pragma unroll
    for (int r = 0; r < C::PF_ROWS; ++r) {
      const int ci = warp + r * NW;
#pragma unroll
      for (int c = 0; c < C::PF_COLS; ++c) {
        const int pcol = lane + c * 32;
        if (ci < CK && pcol < XW) dst[ci * XW + pcol] = pf[r][c];
      }
    }
  };
  auto load_w_async = [&](int stage, int ci0) {
    float* dst = ws + stage * CK * wrow;
    const int n4 = CK * K * (BM / 4);
    for (int i = tid; i < n4; i += NT) {
      const int m4 = i % (BM / 4);
      const int rk = i / (BM / 4);  // ci*K + k
      const int ci = rk / K;
      const int co = co0 + m4 * 4;
      const bool ok = (ci0 + ci) < p.Cin && co < p.Cout;
      const float* src = ok ? wbase + ((size_t)(ci0 + ci) * K + (rk - ci * K)) * p.Cout + co : wbase;
      cp_async16(dst + rk * BM + m4 * 4, src, ok ? 16 : 0);
    }
    cp_async_commit();
  };

  float acc[TM][TN];
#pragma unroll
  for (int m = 0; m < TM; ++m)
#pragma unroll
    for (int i = 0; i < TN; ++i) acc[m][i] = 0.f;

  const int nchunks = (p.Cin + CK - 1) / CK;
  load_x_regs(0);
  load_w_async(0, 0);
  store_x_regs(0);
  cp_async_wait_all();
  __syncthreads();

  for (int ch = 0; ch < nchunks; ++ch) {
    const int st = ch & 1;
    const bool more = (ch + 1) < nchunks;
    if (more) {
      load_x_regs((ch + 1) * CK);
      load_w_async(st ^ 1, (ch + 1) * CK);
    }
    const float* xsst = xs + st * CK * XW + tx;
    const float* wsst = ws + st * CK * wrow + ty * TM;
#pragma unroll 2
    for (int ci = 0; ci < CK; ++ci) {
      const float* xr = xsst + ci * XW;
      const float* wr = wsst + ci * wrow;
      for (int k = 0; k < K; ++k) {
        float xv[TN];
#pragma unroll
        for (int i = 0; i < TN; ++i) xv[i] = xr[k * dil + i * LX];
        float wv[TM];
#pragma unroll
        for (int m4 = 0; m4 < TM / 4; ++m4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wr + k * BM + m4 * 4);
          wv[m4 * 4 + 0] = w4.x;
          wv[m4 * 4 + 1] = w4.y;
          wv[m4 * 4 + 2] = w4.z;
          wv[m4 * 4 + 3] = w4.w;
        }
#pragma unroll
        for (int m = 0; m < TM; ++m)
#pragma unroll
          for (int i = 0; i < TN; ++i) acc[m][i] = fmaf(wv[m], xv[i], acc[m][i]);
      }
    }
    if (more) {
      store_x_regs(st ^ 1);
      cp_async_wait_all();
    }
    __syncthreads();
  }

  // ------------------------------------------------------------------ epilogue
  const int cbase = co0 + ty * TM;
  if (p.epi == EPI_PLAIN) {
    float* yb = p.y + (size_t)b * p.y_bs;
    const float* rb = p.res ? p.res + (size_t)b * p.res_bs : nullptr;
#pragma unroll
    for (int m = 0; m < TM; ++m) {
      const int co = cbase + m;
      if (co >= p.Cout) continue;
      const float bv = p.bias ? p.bias[co] : 0.f;
#pragma unroll
      for (int i = 0; i < TN; ++i) {
        const int t = t0 + tx + i * LX;
        if (t >= Lb) continue;
        float v = acc[m][i] + bv;
        if (p.act == ACT_RELU) v = fmaxf(v, 0.f);
        else if (p.act == ACT_TANH) v = tanhf(v);
        const int to = t * p.n_phase + phase;
        if (rb) v += rb[(size_t)co * p.res_ld + to];
        yb[(size_t)co * p.y_ld + to] = v;
      }
    }
  } else if (p.epi == EPI_GATE) {
    // packed channel 2c = tanh half (reference channel c), 2c+1 = sigmoid half (channel c + Cout/2)
    float* yb = p.y + (size_t)b * p.y_bs;
#pragma unroll
    for (int m = 0; m < TM; m += 2) {
      const int co = cbase + m;
      if (co >= p.Cout) continue;
      const float ba = p.bias[co], bb = p.bias[co + 1];
#pragma unroll
      for (int i = 0; i < TN; ++i) {
        const int t = t0 + tx + i * LX;
        if (t >= Lb) continue;
        const float a = acc[m][i] + ba, g = acc[m + 1][i] + bb;
        yb[(size_t)(co >> 1) * p.y_ld + t] = tanhf(a) * sigmoidf_(g);
      }
    }
  } else if (p.epi == EPI_RESSKIP) {
    // co < split: y[co] = y[co] + v (hidden state, in place); co >= split: aux[co-split] (+)= v
    float* yb = p.y + (size_t)b * p.y_bs;
    float* ab = p.aux + (size_t)b * p.aux_bs;
#pragma unroll
    for (int m = 0; m < TM; ++m) {
      const int co = cbase + m;
      if (co >= p.Cout) continue;
      const float bv = p.bias[co];
#pragma unroll
      for (int i = 0; i < TN; ++i) {
        const int t = t0 + tx + i * LX;
        if (t >= Lb) continue;
        const float v = acc[m][i] + bv;
        if (co < p.split) {
          float* d = yb + (size_t)co * p.y_ld + t;
          *d = *d + v;
        } else {
          float* d = ab + (size_t)(co - p.split) * p.aux_ld + t;
          *d = p.first ? v : (*d + v);
        }
      }
    }
  } else {  // EPI_COUPLING: packed channel 2i = m_i, 2i+1 = logs_i
    // z1_i = (x1_i - m_i) * exp(-logs_i)                     (attentions.py:139)
    // then InvConvNear^-1 on groups (x0[2q], x0[2q+1], z1[2q], z1[2q+1]) (layers.py:246-271)
    // then ActNorm^-1: (z - bias) * exp(-logs)                (layers.py:195)
    float* sb = p.y + (size_t)b * p.y_bs;  // flow state [2*half][y_ld], updated in place
    const int half = p.half;
    float wi[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) wi[j] = p.inv_w[j];
#pragma unroll
    for (int m = 0; m < TM; m += 4) {
      const int co = cbase + m;
      if (co >= p.Cout) continue;
      const int i0 = co >> 1;  // even: coupling channel pair (i0, i0+1) == (2q, 2q+1)
      const float bm0 = p.bias[co], bl0 = p.bias[co + 1], bm1 = p.bias[co + 2], bl1 = p.bias[co + 3];
      const int ch[4] = {i0, i0 + 1, half + i0, half + i0 + 1};
      float ab_[4], as_[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        ab_[o] = p.an_bias[ch[o]];
        as_[o] = p.an_scale[ch[o]];
      }
#pragma unroll
      for (int i = 0; i < TN; ++i) {
        const int t = t0 + tx + i * LX;
        if (t >= Lb) continue;
        float vin[4];
        vin[0] = sb[(size_t)ch[0] * p.y_ld + t];
        vin[1] = sb[(size_t)ch[1] * p.y_ld + t];
        const float x1a = sb[(size_t)ch[2] * p.y_ld + t];
        const float x1b = sb[(size_t)ch[3] * p.y_ld + t];
        vin[2] = (x1a - (acc[m][i] + bm0)) * expf(-(acc[m + 1][i] + bl0));
        vin[3] = (x1b - (acc[m + 2][i] + bm1)) * expf(-(acc[m + 3][i] + bl1));
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          float z = wi[o * 4 + 0] * vin[0];
          z = fmaf(wi[o * 4 + 1], vin[1], z);
          z = fmaf(wi[o * 4 + 2], vin[2], z);
          z = fmaf(wi[o * 4 + 3], vin[3], z);
          sb[(size_t)ch[o] * p.y_ld + t] = (z - ab_[o]) * as_[o];
        }
      }
    }
  }
}

// Host-side launcher.  Picks a tile configuration from the problem size.
int launch_conv1d(const ConvParams& p, int batch, int Lmax, cudaStream_t stream);
void conv1d_init_attributes();

}  // namespace b2
