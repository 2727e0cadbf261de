// enc_chain_params.cuh — host-visible step table of the persistent text-encoder kernel (enc_chain.cuh).
#pragma once

namespace b2 {

constexpr int EC_MAX_STEPS = 64;
constexpr int EC_YT = 16;  // CTA rows per time tile (ffn1 has 12 N tiles; the duration conv and proj_m share a step: 8 + 5)

enum EncStepType { ES_CONV = 0, ES_REDUCE = 1, ES_ATTN = 2 };

// One Conv1d as tcgen05 implicit GEMM over 3-way split bf16 operand images.  CTA row y in [y0, y0 + ntiles*ksplit) computes
// output-channel tile (y - y0) % ntiles over the input-channel chunks of K-split (y - y0) / ntiles.
struct EncConvOp {
  const void* w;         // weight image (pack.umma_image, nsplit 3)
  const float* bias;     // added in the epilogue; nullptr: raw sums (K-split partials, the following reduce step adds it)
  int Cin, Cout, K, pad, NT, ntiles, ksplit, y0, act;
  int planes;            // activation planes multiplied: 3 = fp32-grade (hi, lo, lo2), 1 = bf16 activations (`half` mode)
  const unsigned char* xp;  // input operand image
  long long xp_bs;
  int xp_kc;
  unsigned char* yp;     // output operand image (3 planes) or nullptr
  long long yp_bs;
  int yp_kc;
  float* y;              // fp32 output [batch][Cout][y_ld] (+ ks * part_stride for K-split partials) or nullptr
  long long y_bs, part_stride;
  int y_ld;
};

// Sum of K-split partials (+ bias, + residual) followed by the channel LayerNorm (glow_tts/layers.py:19-28), or just the
// sum when gamma == nullptr.  Rows are split 128 / EC_YT = 8 per CTA over all CTA rows of each time tile.
struct EncReduceOp {
  const float* x;        // [parts][batch][C][ld]
  long long x_bs, part_stride;
  int ld, parts, C, pre_relu, post_relu;
  const float* bias;     // [C] or nullptr
  const float* res;      // residual [batch][C][ld] or nullptr (may alias y: every element is read and written by one thread)
  long long res_bs;
  const float *gamma, *beta;
  float* y;              // fp32 out [batch][C][ld] or nullptr
  long long y_bs;
  unsigned char* yp;     // operand image out (3 planes) or nullptr
  long long yp_bs;
  int yp_kc;
};

// relative-position multi-head attention (glow_tts/attentions.py:214-264) for sequences of up to 128 phonemes; CTA rows
// y0 .. y0 + n_heads - 1 of time tile 0 take one head each
struct EncAttnOp {
  const unsigned char* qkv_p;
  long long qkv_bs;
  int qkv_kc;
  const unsigned char *relk_img, *relv_img;
  int window, H, n_heads, y0;
  int planes;            // 3: all six split products of Q K'^T and P' V'; 1: bf16 operands (hi x hi only)
  unsigned char* out_p;
  long long out_bs;
  int out_kc;
};

struct EncStep {
  int type, n_conv;
  union {
    EncConvOp conv[2];
    EncReduceOp red;
    EncAttnOp att;
  };
};

struct EncChainParams {
  const EncStep* steps;   // device array [n_steps]
  int n_steps, batch, rows;  // rows: rows per 8-channel group of every operand image
  const int* lens;        // [batch] phoneme counts
  unsigned int* counter;  // [batch][n_steps], zeroed before the launch
  unsigned long long* trace;  // debug stamps or nullptr
};

}  // namespace b2
