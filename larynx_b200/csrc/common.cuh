// common.cuh — shared declarations for libb200tts (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <stdexcept>
#include <string>

namespace b2 {

struct Error : public std::runtime_error {
  explicit Error(const std::string& m) : std::runtime_error(m) {}
};

#define B2_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      char _buf[512];                                                                              \
      snprintf(_buf, sizeof(_buf), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,   \
               __LINE__, cudaGetErrorString(_e));                                                  \
      throw b2::Error(_buf);                                                                       \
    }                                                                                              \
  } while (0)

#define B2_CHECK(cond, msg)                                                                        \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      char _buf[512];                                                                              \
      snprintf(_buf, sizeof(_buf), "%s (%s) at %s:%d", msg, #cond, __FILE__, __LINE__);            \
      throw b2::Error(_buf);                                                                       \
    }                                                                                              \
  } while (0)

// ---- fused conv1d parameters (see conv1d.cuh) ------------------------------------------------
enum PreMode { PRE_NONE = 0, PRE_LRELU = 1 };
enum EpiMode { EPI_PLAIN = 0, EPI_GATE = 1, EPI_RESSKIP = 2, EPI_COUPLING = 3 };
enum ActMode { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

struct ConvParams {
  // input: x (+ x2 + x3 when non-null), all [batch][Cin][x_ld]
  const float* x = nullptr;
  const float* x2 = nullptr;
  const float* x3 = nullptr;
  long long x_bs = 0;
  int x_ld = 0;
  float in_div = 1.f;  // summed input is divided by this (MRF mean: 3)
  int pre = PRE_NONE;
  float pre_slope = 0.f;
  // weights packed [phase][Cin][K][Cout] (Cout contiguous), bias [Cout]
  const float* w = nullptr;
  const float* bias = nullptr;
  // output [batch][*][y_ld]
  float* y = nullptr;
  long long y_bs = 0;
  int y_ld = 0;
  // EPI_PLAIN: optional residual added after activation
  const float* res = nullptr;
  long long res_bs = 0;
  int res_ld = 0;
  // EPI_RESSKIP: skip accumulator
  float* aux = nullptr;
  long long aux_bs = 0;
  int aux_ld = 0;
  // valid input length of item b = lens ? lens[b] * len_mul : Lin  (zero padded outside)
  const int* lens = nullptr;
  int len_mul = 1;
  int Lin = 0;
  int Cin = 0, Cout = 0, K = 1, dil = 1, pad_left = 0;
  int n_phase = 1;  // >1: transposed conv with stride n_phase, K must be 2 (polyphase taps)
  int tpad = 0;     // transposed conv padding (k-u)/2
  int epi = EPI_PLAIN, act = ACT_NONE;
  int split = 0, first = 0;  // EPI_RESSKIP
  // EPI_COUPLING: inverse 4x4 invertible conv + inverse ActNorm fused behind the affine coupling
  const float* inv_w = nullptr;     // [4][4] W^-1
  const float* an_bias = nullptr;   // [2*half]
  const float* an_scale = nullptr;  // [2*half] = exp(-logs)
  int half = 0;                     // channels per coupling half (80)
  float* tap = nullptr;             // optional copy of the updated flow state (same layout as y; debug taps)
  // tensor-core path (conv_umma.cuh): pre-split bf16 weight image, N tile, split count, pixel-shuffle factor
  const void* wu = nullptr;
  int um_nt = 0, um_nsplit = 2, pix = 1;
  unsigned long long* dbg = nullptr;  // optional per-phase %globaltimer trace of CTA (0,0,0)
  // bf16 plane images (conv_umma.cuh): xp = input image written by the producing layer (enables bulk-copy A),
  // yp / yp2 = images this launch writes for its consumers (yp2: the WN skip accumulator of EPI_RESSKIP)
  const void* xp = nullptr;
  long long xp_bs = 0;  // bytes between batch items
  int xp_kc = 0, xp_rows = 0;
  void* yp = nullptr;
  void* yp2 = nullptr;
  long long yp_bs = 0;
  int yp_kc = 0, yp_rows = 0, yp_nsplit = 2;
  float yp_slope = 1.f;  // planes hold v > 0 ? v : v*slope (the consumer's LeakyReLU); 1 = identity
  // residual read from a 2-plane bf16 image instead of fp32 `res` (EPI_PLAIN): res = unlrelu(hi + lo), i.e. the image
  // of the tensor this block started from; may alias yp (each thread reads its elements before it rewrites them)
  const void* rp = nullptr;
  long long rp_bs = 0;
  int rp_kc = 0, rp_rows = 0;
  float rp_inv_slope = 1.f;  // 1 / slope of the LeakyReLU the image was stored with
};

// up to three independent convs of identical output geometry in ONE launch (the three MRF ResBlock branches of a
// HiFi-GAN stage: same C and L, different kernel size / dilation / weights / buffers); blockIdx.z = branch*zb + item
struct ConvBatch {
  int n = 1;
  int zb = 1;  // batch items per branch
  int gx[3] = {0, 0, 0};  // persistent launches: CTAs (along x) that share the time tiles of branch i; 0 = gridDim.x
  ConvParams p[3];
};

// convs per launch of the per-stage dataflow kernel (stage_umma.cuh): 3 MRF branches x 3 dilations x (c1, c2)
constexpr int SG_MAX_CONVS = 18;

}  // namespace b2
