// flow_chain_params.cuh — host-visible parameter structs of the persistent flow-decoder kernel (flow_chain.cuh).
#pragma once

namespace b2 {

constexpr int FC_MAX_WN = 4;                 // WN layers per coupling block
constexpr int FC_STEPS_MAX = 2 + 2 * FC_MAX_WN;

// one coupling block: weights in kernel layout (device pointers into the model blob)
struct FlowBlockDev {
  const void* w[FC_STEPS_MAX];   // start, in0, rs0, in1, rs1, ..., end: bf16 hi/lo operand images (pack.umma_image)
  const float* b[FC_STEPS_MAX];  // biases (packed channel order)
  const float* inv;              // [4][4] W^-1 of InvConvNear
  const float* an_bias;          // [2*half] ActNorm bias
  const float* an_scale;         // [2*half] exp(-logs)
  int tap_index;                 // slot of this block in FlowChainParams::taps (the model's block index)
  int pad_;
};

struct FlowChainParams {
  const FlowBlockDev* blocks;  // [n_blocks] in EXECUTION order (the decoder runs its flows reversed)
  int n_blocks, n_wn;
  int batch;
  int half, Hd, K, dil_rate;   // 80, 192, 5, 1
  int nt_end;                  // N tile of the `end` conv
  const int* lens;             // [batch] squeezed lengths
  float* S;                    // fp32 flow state [batch][2*half][S_ld], updated in place
  long long S_bs;
  int S_ld;
  float* taps;                 // optional [n_blocks][batch][2*half][S_ld]: state after block `tap_index`
  unsigned char *S_p, *H_p, *A_p, *K_p;  // operand images: flow state, WN hidden state, gate output, skip sum
  long long Sp_bs, Hp_bs;      // bytes per batch item (S_p | the three Hd-channel images)
  int Sp_kc, Hp_kc, rows;
  unsigned int* counter;       // [batch][n_blocks * (2 + 2 n_wn)], zeroed before the launch
  unsigned long long* trace;   // debug: %globaltimer stamps [layer][6 CTA rows][8 events] of the CTAs (0, y, 0); or nullptr
};

}  // namespace b2
