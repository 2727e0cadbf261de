/*
 * b200tts.h — C ABI of libb200tts.so: the B200-native replacement for the two native
 * calls on Larynx's hot path.
 *
 * In the reference the only FFI into native code on this path is onnxruntime:
 *   larynx/glow_tts.py:98-100   onnxruntime.InferenceSession(generator.onnx)        -> b200tts_glow_create
 *   larynx/glow_tts.py:161-168  session.run(None, {input, input_lengths, scales})   -> b200tts_glow_encode + b200tts_glow_decode
 *   larynx/hifi_gan.py:103-105  onnxruntime.InferenceSession(generator.onnx)        -> b200tts_hifigan_create
 *   larynx/hifi_gan.py:150      session.run(None, {"mel": mels})                    -> b200tts_hifigan_infer
 *   larynx/hifi_gan.py:168-169  audio_float_to_int16 (larynx/audio.py:118-125)      -> pcm_i16 output of b200tts_hifigan_infer
 *   larynx/__init__.py:231-257  phonemes_to_mels -> numpy mel glue -> mels_to_audio -> b200tts_tts_finish (fused, no host round trip)
 * (the PyTorch backend, larynx/glow_tts.py:142-149 / larynx/hifi_gan.py:141-144, computes the same
 * functions; it is the parity oracle, see oracle/port.py.)
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; b200tts_last_error() returns a
 *    thread-local message for the last failure on the calling thread.
 *  - data pointers (`ids`, `mel`, `pcm`, `eps`, ...) may be HOST or DEVICE pointers; the library
 *    detects which (cudaPointerGetAttributes) and stages host buffers through pinned memory.
 *  - model objects (b200tts_glow / b200tts_hifigan) are immutable after creation and may be shared
 *    by any number of host threads.  All mutable state (stream, workspaces, pinned staging) lives
 *    in a b200tts_call; use one b200tts_call per host thread (the reference runs _sentence_task on
 *    a ThreadPoolExecutor against one shared model object, larynx/__init__.py:146-157).
 *  - layouts are the reference's: activations are [batch][channels][time], time contiguous.
 *  - there is NO CPU fallback: without a CUDA device every entry point fails.
 */
#ifndef B200TTS_H
#define B200TTS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200TTS_VERSION 1

typedef struct b200tts_glow b200tts_glow;       /* GlowTTS weights on one device   */
typedef struct b200tts_hifigan b200tts_hifigan; /* HiFi-GAN weights on one device  */
typedef struct b200tts_call b200tts_call;       /* per-thread stream + workspaces  */

/* Mel glue applied between the two passes by larynx/__init__.py:242-249 using the voice's
 * AudioSettings (larynx/audio.py:25-112).  All fields mirror AudioSettings attributes. */
typedef struct b200tts_glue {
  int32_t signal_norm;                  /* denormalize (audio.py:83-104)                    */
  int32_t symmetric_norm;
  int32_t clip_norm;
  int32_t convert_db_to_amp;            /* db_to_amp (audio.py:58-59)                       */
  int32_t do_dynamic_range_compression; /* dynamic_range_compression (audio.py:106-108)     */
  float max_norm;
  float min_level_db;
  float ref_level_db;
  float spec_gain;
} b200tts_glue;

const char* b200tts_last_error(void);
int b200tts_version(void);
int b200tts_device_count(int* count);

/* ---- per-thread call context -------------------------------------------------------------- */
int b200tts_call_create(int device, b200tts_call** out);
int b200tts_call_destroy(b200tts_call* call);
/* device time (ms) of the last encode/decode/infer/finish issued on this call, measured with
 * CUDA events on the call's stream */
int b200tts_call_last_ms(b200tts_call* call, float* ms);
/* number of kernels this library launched on this call since creation */
int b200tts_call_launches(b200tts_call* call, int64_t* n);
/* the cudaStream_t owned by the call (for external event timing / profiling) */
int b200tts_call_stream(b200tts_call* call, void** stream);

/* ---- GlowTTS (glow_tts/models.py:308-354 FlowGenerator.forward, inference branch) -------- */
/* `blob` is the packed weight image written by larynx_b200.pack (host or device memory). */
int b200tts_glow_create(const void* blob, size_t nbytes, int device, b200tts_glow** out);
int b200tts_glow_destroy(b200tts_glow* m);
int b200tts_glow_info(const b200tts_glow* m, int32_t* num_symbols, int32_t* mel_channels, int32_t* n_sqz);

/* Phase A: text encoder + duration predictor + ceil/cumsum/length (models.py:118-140,323-331).
 * ids [batch][max_tx] int64 (rows padded with anything past x_len[b]); y_len_out[b] receives the
 * number of mel frames of item b (already truncated to a multiple of n_sqz, models.py:356-363).
 * This call synchronises the stream (the output length is data dependent). */
int b200tts_glow_encode(const b200tts_glow* m, b200tts_call* call, const int64_t* ids, const int32_t* x_len,
                        int batch, int max_tx, float length_scale, int32_t* y_len_out);

/* Phase B: monotonic expand + noise + 12-block flow decoder reverse (models.py:335-349,191-209).
 * eps: optional injected N(0,1) noise [batch][mel][eps_ld] (eps_ld >= max y_len); when NULL a
 * Philox stream seeded by `seed` is used.  mel_out [batch][mel][mel_ld], mel_ld >= max y_len;
 * columns >= y_len[b] are zero (the reference's z_mask). */
int b200tts_glow_decode(const b200tts_glow* m, b200tts_call* call, float noise_scale, const float* eps, int eps_ld,
                        uint64_t seed, float* mel_out, int mel_ld);

/* Reduced-precision encoder (BASELINE config 3; the reference's counterpart is `half=True`, which casts the whole
 * PyTorch model to fp16: larynx/glow_tts.py:90-91).  on != 0: every following b200tts_glow_encode on this call multiplies
 * the prenet, q/k/v/o, attention and FFN GEMMs on bf16-rounded activations (one operand plane instead of three; fp32
 * accumulation, fp32 LayerNorm / softmax, weights at full split precision) when the batch takes the single-launch
 * encoder (max_tx <= 128); proj_m, the duration predictor, the flow decoder and the vocoder are unchanged.  Alignments
 * can differ from the fp32 result in this mode — tools/half_report.py measures how often.  Default 0. */
int b200tts_call_set_encoder_bf16(b200tts_call* call, int on);

/* Debug/parity taps of the last encode on this call (any pointer may be NULL):
 * logw [batch][max_tx] f32, w_ceil [batch][max_tx] i32, idx [batch][idx_ld] i32 (phoneme index
 * each frame copies, -1 beyond y_len) — "alignment indices bit-exact" is checked on these. */
int b200tts_glow_alignment(const b200tts_glow* m, b200tts_call* call, float* logw, int32_t* w_ceil, int32_t* idx,
                           int idx_ld);
/* Copy flow-decoder state after block `blk` was inverted (blk = n_blocks-1 .. 0; requires that
 * b200tts_call_set_debug(call,1) was set before decode). out [batch][2*mel][ld] */
int b200tts_call_set_debug(b200tts_call* call, int on);
int b200tts_glow_flow_tap(const b200tts_glow* m, b200tts_call* call, int blk, float* out, int ld);
/* The production noise stream on its own (no reference counterpart: glow_tts/models.py:348 draws torch.randn_like):
 * out [batch][mel][ld] (host or device), element (b, c, f) = the N(0,1) draw that decode with eps == NULL and this seed
 * adds to frame f of channel c of item b.  Exists so that the generator's moments can be tested. */
int b200tts_debug_noise(b200tts_call* call, uint64_t seed, int batch, int mel, int n, float* out, int ld);

/* Host logic of the per-stage HiFi-GAN kernels on its own (no device needed, no reference counterpart): how the time tiles of
 * `n_strips` strips (strip i: strip_tiles[i] tiles of modelled cost strip_cost[i] each) are cut into contiguous, equal-cost
 * ranges for `nctas` CTAs.  cta_of_tile [sum of strip_tiles], in strip order; returns 0 and writes the number of ranges. */
int b200tts_debug_stage_partition(const int32_t* strip_tiles, const int64_t* strip_cost, int n_strips, int nctas,
                                  int32_t* cta_of_tile, int32_t* n_ranges);

/* ---- HiFi-GAN (hifi_gan/models.py:186-202 Generator.forward) ------------------------------ */
int b200tts_hifigan_create(const void* blob, size_t nbytes, int device, b200tts_hifigan** out);
int b200tts_hifigan_destroy(b200tts_hifigan* m);
int b200tts_hifigan_info(const b200tts_hifigan* m, int32_t* mel_channels, int32_t* hop);

/* mel [batch][80][mel_ld] (already glued, i.e. what the reference feeds the vocoder);
 * frames[b] = valid frames of item b (ragged batch: every conv zero-pads at each item's own end,
 * so each item equals a per-utterance call — the reference Generator has no mask).
 * pcm_f32 (nullable) [batch][pcm_ld] float in (-1,1); pcm_i16 (nullable) [batch][pcm_ld] after
 * audio_float_to_int16 (per item peak normalisation). pcm_ld >= hop*max(frames). */
int b200tts_hifigan_infer(const b200tts_hifigan* m, b200tts_call* call, const float* mel, const int32_t* frames,
                          int batch, int mel_ld, float* pcm_f32, int16_t* pcm_i16, int pcm_ld);

/* ---- spectral-subtraction denoiser (larynx/hifi_gan.py:152-203, STFT of larynx/audio.py:232-306) ---------- */
/* Strength applied by every following b200tts_hifigan_infer / b200tts_tts_finish on this call (settings key
 * "denoiser_strength", larynx/hifi_gan.py:152-156; the CLI/server default is 0.005); <= 0 switches it off (default).
 * On the device: 1024-point Hann STFT with hop 256 over frames 0, 256, ... < N-1024, |X| <- max(|X| - strength*bias, 0)
 * with the phase kept, windowed inverse FFT, overlap-add without window normalisation — the reference's arithmetic.
 * pcm_f32 then holds the denoised waveform and pcm_i16 its peak-normalised int16 conversion.  Items of at most 4 mel
 * frames (N <= 1024) have no STFT frame: the reference returns 1024 zeros for them, this library zero-fills the item. */
int b200tts_call_set_denoiser(b200tts_call* call, float strength);
/* The bias spectrum (larynx/hifi_gan.py:181-203): |STFT frame 0| of the vocoder's response to zeros[1][mel][88],
 * computed on the device when the model is created.  out receives 513 floats (n = capacity of out). */
int b200tts_hifigan_bias_spectrum(const b200tts_hifigan* m, float* out, int n);

/* HiFiGanVocoder.denoise (larynx/hifi_gan.py:171-179) as a call of its own.  pcm_in [batch][in_ld] float waveforms with
 * n_samples[b] valid samples; bias (nullable): 513 floats used instead of the model's bias spectrum.  Item b of the outputs
 * [batch][out_ld] holds n_t*256 + 1024 samples, n_t = number of frame starts 0, 256, ... < n_samples[b] - 1024 (larynx/audio.py:
 * 232-267; equal to n_samples[b] when that is a multiple of 256 above 1024), zeros beyond.  pcm_f32 / pcm_i16 nullable;
 * pcm_i16 is audio_float_to_int16 of the denoised item.  strength 0 still runs STFT -> ISTFT (the Hann window is applied twice
 * and never normalised, exactly like the reference). */
int b200tts_hifigan_denoise(const b200tts_hifigan* m, b200tts_call* call, float strength, const float* bias,
                            const float* pcm_in, const int32_t* n_samples, int batch, int in_ld, float* pcm_f32,
                            int16_t* pcm_i16, int out_ld);

/* ---- fused ids -> PCM (larynx/__init__.py:229-257 without leaving the device) ------------- */
/* Must follow b200tts_glow_encode on the same call.  Runs decode, the mel glue, the vocoder and
 * the int16 conversion back to back on the call's stream.  mel_out (nullable) receives the
 * un-glued mel like b200tts_glow_decode. */
int b200tts_tts_finish(const b200tts_glow* g, const b200tts_hifigan* v, b200tts_call* call, const b200tts_glue* glue,
                       float noise_scale, const float* eps, int eps_ld, uint64_t seed, float* mel_out, int mel_ld,
                       float* pcm_f32, int16_t* pcm_i16, int pcm_ld);

#ifdef __cplusplus
}
#endif
#endif /* B200TTS_H */
