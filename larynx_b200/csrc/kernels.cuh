// kernels.cuh — launch wrappers of the non-conv kernels (definitions in kernels.cu).
#pragma once
#include "../../include/b200tts.h"
#include "common.cuh"

namespace b2 {

int launch_conv1d(const ConvParams& p, int batch, int Lmax, cudaStream_t stream);
void conv1d_init_attributes();
int launch_conv_umma(const ConvParams& p, int batch, int Lmax, cudaStream_t stream);
int launch_conv_umma_batch(const ConvParams* ps, int n, int batch, int Lmax, cudaStream_t stream);
void conv_umma_init_attributes();
// one ResBlock1 step (c1 -> LeakyReLU -> c2 + residual) per launch, T kept in shared memory (rbpair_umma.cu)
void rbpair_umma_init_attributes();
bool rbpair_umma_supported(const ConvParams* c1, const ConvParams* c2, int n);
int launch_rbpair_umma(const ConvParams* c1, const ConvParams* c2, int n, int batch, int Lmax, cudaStream_t stream);
// all ResBlock convs of a HiFi-GAN stage in one persistent launch with tile-level dataflow (stage_umma.cu)
void stage_umma_init_attributes();
void stage_set_striped(bool on);  // item order of the stage kernels (stage_umma.cu); process-wide, set once at start-up
bool stage_umma_supported(const ConvParams* convs, int n_levels, int n_cols);
size_t stage_umma_table_bytes();
int launch_stage_umma(const ConvParams* convs, int n_levels, int n_cols, int batch, const int* lens_h, void* tab_h, void* tab_d,
                      unsigned int* flags_d, size_t flags_cap, cudaStream_t s, bool trace = false);
// ... and with a whole ResBlock step (c1 -> c2, T in shared memory) per item for the 32-channel stage
bool stage_pair_supported(const ConvParams* c1, const ConvParams* c2, int n_levels, int n_cols);
int launch_stage_pair(const ConvParams* c1, const ConvParams* c2, int n_levels, int n_cols, int batch, const int* lens_h, void* tab_h,
                      void* tab_d, unsigned int* flags_d, size_t flags_cap, cudaStream_t s);
// host logic of the stage kernels' work split, exposed for tests (b200tts_debug_stage_partition)
struct StageStrip;
struct StageSeg;
int stage_partition(const StageStrip* strips, int n_strips, int nctas, int* seg_off, StageSeg* segs, int max_segs);
int stage_partition_debug(const int* strip_tiles, const long long* strip_cost, int n_strips, int nctas, int* cta_of_tile);
void attention_init_attributes();
void attn_umma_init_attributes();
bool attn_umma_supported(int H, int n_heads, int window, int Tmax);
void launch_attn_umma(const void* qkv_p, long long qkv_bs, int qkv_kc, int qkv_rows, const void* relk_img,
                      const void* relv_img, int window, int H, int n_heads, const int* lens,
                      int batch, void* out_p, long long out_bs, int out_kc, int out_rows, float* out, long long out_fbs,
                      int out_ld, cudaStream_t s, int Tmax);

void launch_embed(const long long* ids, const int* lens, int batch, int max_tx, const float* emb, int n_vocab, int H,
                  float* x, long long x_bs, int x_ld, int* bad_id, cudaStream_t s);
void launch_noise_fill(unsigned long long seed, int batch, int mel, int n, float* out, long long out_bs, int out_ld, cudaStream_t s);
void launch_layernorm(const float* x, const float* res, const float* gamma, const float* beta, float* y, long long bs,
                      int ld, int C, const int* lens, int batch, int Lmax, int pre_relu, int post_relu, cudaStream_t s,
                      void* planes = nullptr, long long pl_bs = 0, int pl_kc = 0, int pl_rows = 0, int pl_nsplit = 3);
void launch_rel_attention(const float* qkv, long long bs, int ld, int H, int n_heads, const float* emb_k,
                          const float* emb_v, int window, const int* lens, int batch, int Tmax, float* out,
                          long long out_bs, int out_ld, cudaStream_t s);
void launch_duration(const float* h, long long h_bs, int h_ld, int C, const float* pw, const float* pb,
                     const int* lens, int batch, int max_tx, float length_scale, int n_sqz, float* logw, int* w_ceil,
                     int* cum, int* y_len, cudaStream_t s);
void launch_align_index(const int* cum, const int* lens, const int* y_len, int batch, int max_tx, int* idx, int idx_ld,
                        int* sq_len, int n_sqz, cudaStream_t s);
void launch_expand_squeeze(const float* x_m, long long xm_bs, int xm_ld, const int* idx, int idx_ld, const int* y_len,
                           const float* eps, long long eps_bs, int eps_ld, float noise_scale, unsigned long long seed,
                           int mel, int n_sqz, float* zs, long long zs_bs, int zs_ld, int batch, int Tymax,
                           cudaStream_t s);
void launch_unsqueeze(const float* zs, long long zs_bs, int zs_ld, const int* y_len, int mel, int n_sqz, float* out,
                      long long out_bs, int out_ld, float* glued, long long gl_bs, int gl_ld, const b200tts_glue* g,
                      int batch, int width, cudaStream_t s);
void launch_conv_post(const float* x1, const float* x2, const float* x3, long long bs, int ld, float in_div,
                      float slope, const float* w, const float* bias, int C, int K, const int* frames, int len_mul,
                      int batch, int Lmax, float* pcm, int pcm_ld, unsigned int* peak_bits, cudaStream_t s);
void launch_pcm_to_i16(const float* pcm, int pcm_ld, const unsigned int* peak_bits, const int* frames, int len_mul,
                       int batch, short* out, int out_ld, cudaStream_t s);
void launch_zero_tail(float* y, long long bs, int ld, int C, const int* lens, int len_mul, int batch, int width,
                      cudaStream_t s);

// spectral-subtraction denoiser (larynx/hifi_gan.py:171-179; STFT constants larynx/audio.py:232-249)
constexpr int DN_FFT = 1024, DN_HOP = 256, DN_BINS = DN_FFT / 2 + 1;
// tw: 1024 (cos, sin) pairs of exp(-2 pi i t / 1024); hann: numpy.hanning(1024); fr: scratch [batch][fr_bs] floats,
// fr_bs >= max_frames * 1024.  mag_out != nullptr: only write |X| (513 values) of frame 0 of item 0.
void launch_stft_denoise(const float* pcm, int pcm_ld, const int* frames, int len_mul, int batch, int max_frames,
                         const float* hann, const float* tw, const float* bias, float strength, float* fr,
                         long long fr_bs, float* mag_out, cudaStream_t s);
void launch_overlap_add(const float* fr, long long fr_bs, const int* frames, int len_mul, int batch, int max_len,
                        float* pcm, int pcm_ld, unsigned int* peak_bits, cudaStream_t s);

// bf16 hi/lo operand image of leaky_relu(((x1 + x2) + x3) / div, slope) for the tensor-core convs (conv_umma.cu)
void launch_mean3_image(const float* x1, const float* x2, const float* x3, long long bs, int ld, int C, float div, float slope,
                        const int* lens, int len_mul, int batch, void* img, long long img_bs, int kc_total, int rows,
                        cudaStream_t s);

// The persistent kernels whose CTAs wait for each other (enc_chain, flow_chain, stage_umma) need all their CTAs resident.
// A cooperative launch checks that against an EMPTY device; two such grids from different host threads' streams can each get
// part of the SMs and then wait for the rest forever (seen with two BatchScheduler workers: bench.py config 5).  The gate
// orders them device-wide without blocking the host: a launch between begin/end waits (cudaStreamWaitEvent) for the previous
// gated launch of any stream on this device.  Kernels that never wait for other CTAs may still overlap them.
void spin_gate_begin(cudaStream_t s);
void spin_gate_end(cudaStream_t s);

// the whole flow decoder in one persistent cooperative launch (flow_chain.cu)
struct FlowChainParams;
void flow_chain_init_attributes();
bool flow_chain_geometry_ok(int half, int Hd, int K, int dil_rate, int n_wn, int nt_hidden, int nt_end);
bool flow_chain_fits(int Lmax);
int launch_flow_chain(const FlowChainParams& fp, int Lmax, cudaStream_t s);

// the text encoder + duration convs in one persistent cooperative launch (enc_chain.cu)
struct EncStep;
struct EncChainParams;
void enc_chain_init_attributes();
bool enc_chain_fits(int max_tx);
int launch_enc_chain(const EncStep* steps, const EncChainParams& fp, int max_tx, cudaStream_t s);
void launch_embed_image(const long long* ids, const int* lens, int batch, int max_tx, const float* emb, int n_vocab, int H, float* x,
                        long long x_bs, int x_ld, void* img, long long img_bs, int kc_total, int rows, int* bad_id, cudaStream_t s);

}  // namespace b2
