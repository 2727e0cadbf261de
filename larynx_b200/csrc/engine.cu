// engine.cu — model objects, per-call contexts and the C ABI of libb200tts.so (include/b200tts.h).
//
// Host orchestration of the hot path:
//   glow encode  : glow_tts/models.py:118-140 (TextEncoder) + :323-331 (durations)
//   glow decode  : glow_tts/models.py:335-349 (expand + noise) + :191-209 (FlowSpecDecoder reverse)
//   hifigan infer: hifi_gan/models.py:186-202 (Generator.forward) + larynx/audio.py:118-125
//   tts finish   : larynx/__init__.py:229-257 (_sentence_task) with the numpy mel glue on device
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "enc_chain_params.cuh"
#include "flow_chain_params.cuh"
#include "kernels.cuh"

using namespace b2;

// ------------------------------------------------------------------------------------------------
// packed weight blob (written by larynx_b200/pack.py)
// ------------------------------------------------------------------------------------------------
#pragma pack(push, 1)
struct BlobHeader {
  char magic[8];  // "B2TTSBLB"
  uint32_t version;
  uint32_t kind;  // 1 = glow, 2 = hifigan
  uint32_t n_tensors;
  uint32_t n_cfg;
  int32_t cfg[64];
  uint64_t total_bytes;
};
struct BlobEntry {
  char name[64];
  uint64_t offset;  // bytes from blob start, 256-byte aligned
  uint64_t numel;
  uint32_t dtype;  // 0 = f32
  uint32_t reserved;
};
#pragma pack(pop)

struct Weights {
  int device = 0;
  void* dev = nullptr;
  size_t nbytes = 0;
  std::map<std::string, std::pair<const float*, size_t>> t;
  int32_t cfg[64];
  int n_cfg = 0;

  void load(const void* blob, size_t nbytes_, int device_, uint32_t kind) {
    B2_CHECK(nbytes_ >= sizeof(BlobHeader), "blob too small");
    device = device_;
    B2_CUDA(cudaSetDevice(device));
    // the header/table may live in device memory: stage the front of the blob on the host
    BlobHeader h;
    B2_CUDA(cudaMemcpy(&h, blob, sizeof(h), cudaMemcpyDefault));
    B2_CHECK(memcmp(h.magic, "B2TTSBLB", 8) == 0, "bad blob magic");
    B2_CHECK(h.version == 1, "unsupported blob version");
    B2_CHECK(h.kind == kind, "blob is for a different model kind");
    B2_CHECK(h.total_bytes == nbytes_, "blob size mismatch");
    B2_CHECK(h.n_tensors < 100000, "implausible tensor count");
    std::vector<BlobEntry> table(h.n_tensors);
    B2_CHECK(sizeof(BlobHeader) + sizeof(BlobEntry) * (size_t)h.n_tensors <= nbytes_, "blob table truncated");
    B2_CUDA(cudaMemcpy(table.data(), (const char*)blob + sizeof(BlobHeader), sizeof(BlobEntry) * h.n_tensors,
                       cudaMemcpyDefault));
    nbytes = nbytes_;
    B2_CUDA(cudaMalloc(&dev, nbytes));
    B2_CUDA(cudaMemcpy(dev, blob, nbytes, cudaMemcpyDefault));
    memcpy(cfg, h.cfg, sizeof(cfg));
    n_cfg = h.n_cfg;
    for (auto& e : table) {
      B2_CHECK(e.offset % 16 == 0 && e.offset + e.numel * 4 <= nbytes, "blob entry out of range");
      char nm[65];
      memcpy(nm, e.name, 64);
      nm[64] = 0;
      t[nm] = {reinterpret_cast<const float*>((const char*)dev + e.offset), (size_t)e.numel};
    }
  }
  bool has(const std::string& name) const { return t.find(name) != t.end(); }
  const float* get_any(const std::string& name, size_t* numel) const {
    auto it = t.find(name);
    if (it == t.end()) throw Error("blob is missing tensor '" + name + "'");
    *numel = it->second.second;
    return it->second.first;
  }
  const float* get(const std::string& name, size_t numel) const {
    auto it = t.find(name);
    if (it == t.end()) throw Error("blob is missing tensor '" + name + "'");
    if (it->second.second != numel) throw Error("tensor '" + name + "' has the wrong size");
    return it->second.first;
  }
  ~Weights() {
    if (dev) cudaFree(dev);
  }
};

// ------------------------------------------------------------------------------------------------
// models
// ------------------------------------------------------------------------------------------------
struct ConvW {
  const float* w = nullptr;   // fp32 SIMT layout [Cin][K][Cout] (may be null for tensor-core-only layers)
  const float* b = nullptr;
  int cin = 0, cout = 0, k = 1;
  const void* wu = nullptr;   // bf16 split image for conv_umma_kernel
  int nt = 0, nsplit = 2;
};

struct GlowFlow {
  ConvW start, end;
  ConvW in[8], rs[8];
  const float *inv = nullptr, *an_bias = nullptr, *an_scale = nullptr;
};
struct GlowEnc {
  ConvW qkv, o, ffn1, ffn2;
  const float *rel_k, *rel_v, *ln1g, *ln1b, *ln2g, *ln2b;
  const void *relk_img = nullptr, *relv_img = nullptr;  // bf16 3-plane operand images of emb_rel_k / emb_rel_v
};

struct b200tts_glow {
  Weights W;
  int V, H, F, Fd, K, n_heads, n_layers, window, mel, n_blocks, kd, dil_rate, n_block_layers, n_sqz, n_split, Hd;
  const float* emb;
  ConvW pre_conv[3], pre_proj, proj_m, dp1, dp2;
  const float *pre_lng[3], *pre_lnb[3], *dp_ln1g, *dp_ln1b, *dp_ln2g, *dp_ln2b, *dp_pw, *dp_pb;
  std::vector<GlowEnc> enc;
  std::vector<GlowFlow> flow;
  // persistent flow-decoder kernel (flow_chain.cuh): the blocks' weights in execution order, nullptr if the geometry
  // is not the one the kernel is built for
  FlowBlockDev* flow_dev = nullptr;
  int flow_nt_end = 0;
  ~b200tts_glow() { cudaFree(flow_dev); }
};

struct HifiRB {
  ConvW c1[3], c2[3];
  int k = 3, ndil = 3, dil[3] = {1, 3, 5};
};
struct b200tts_hifigan {
  Weights W;
  int resblock, n_ups, C0, n_kernels, mel, hop;
  int up_rate[8], up_k[8];
  ConvW conv_pre, conv_post, ups[8], ups_ps[8];
  std::vector<HifiRB> rb;  // [n_ups * n_kernels]
  // denoiser tables (device): numpy.hanning(1024), exp(-2 pi i t / 1024), bias spectrum [513]   larynx/hifi_gan.py:181-203
  float *dn_hann = nullptr, *dn_tw = nullptr, *dn_bias = nullptr;
  ~b200tts_hifigan() {
    cudaFree(dn_hann);
    cudaFree(dn_tw);
    cudaFree(dn_bias);
  }
};

// host copy of a small fp32 tensor that lives in the device blob
static std::vector<float> fetch(const float* dev, size_t n) {
  std::vector<float> h(n);
  B2_CUDA(cudaMemcpy(h.data(), dev, n * sizeof(float), cudaMemcpyDeviceToHost));
  return h;
}

static void bind_umma(const Weights& W, const std::string& name, ConvW& c) {
  if (!W.has(name + ".u")) return;
  std::vector<float> meta = fetch(W.get(name + ".um", 8), 8);
  const int nt = (int)meta[0], nsplit = (int)meta[1], nch = (int)meta[2], k = (int)meta[3], ntiles = (int)meta[4];
  B2_CHECK((int)meta[5] == c.cout && (int)meta[6] == c.cin && k == c.k, "tensor-core weight image does not match the layer");
  size_t numel = 0;
  const float* img = W.get_any(name + ".u", &numel);
  const size_t bytes = (size_t)ntiles * nch * k * nsplit * 4 * nt * 16;
  B2_CHECK(numel * 4 >= bytes, "tensor-core weight image truncated");
  c.wu = img;
  c.nt = nt;
  c.nsplit = nsplit;
}

static ConvW conv_w(const Weights& W, const std::string& name, int cin, int cout, int k, int mult = 1) {
  ConvW c;
  c.cin = cin;
  c.cout = cout;
  c.k = k;
  // fp32 [Cin][K][Cout] copy for the SIMT cross-check path: only in blobs packed for it (pack.want_simt_weights)
  c.w = W.has(name + ".w") ? W.get(name + ".w", (size_t)cin * cout * k * mult) : nullptr;
  c.b = W.get(name + ".b", cout);
  if (mult == 1) bind_umma(W, name, c);
  return c;
}

// ------------------------------------------------------------------------------------------------
// per-call context
// ------------------------------------------------------------------------------------------------
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  void reserve(size_t bytes) {
    off = 0;
    if (bytes <= cap) return;
    if (base) B2_CUDA(cudaFree(base));
    base = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + (1 << 20);
    B2_CUDA(cudaMalloc((void**)&base, want));
    cap = want;
  }
  template <typename T>
  T* take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
    B2_CHECK(off + bytes <= cap, "workspace arena overflow");
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
  static size_t pad(size_t n, size_t elt) { return (n * elt + 255) & ~(size_t)255; }
  ~Arena() {
    if (base) cudaFree(base);
  }
};

struct b200tts_call {
  int device = 0;
  cudaStream_t stream = nullptr, aux[2] = {nullptr, nullptr};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, fork = nullptr, join[2] = {nullptr, nullptr};
  long long launches = 0;
  float last_ms = 0.f;
  int debug = 0;
  float denoise = 0.f;  // denoiser strength applied after the vocoder (0 = off)           larynx/hifi_gan.py:152-166
  int enc_bf16 = 0;     // 1: text-encoder GEMMs on bf16 activations (b200tts_call_set_encoder_bf16)
  Arena enc, dec, voc, stg;
  // state carried from encode to decode
  const b200tts_glow* enc_model = nullptr;
  int batch = 0, max_tx = 0, tx_ld = 0, ty_max = 0;
  std::vector<int> y_len_h;
  int *xlen_d = nullptr, *ylen_d = nullptr, *sqlen_d = nullptr, *wceil_d = nullptr, *cum_d = nullptr;
  float *logw_d = nullptr, *xm_d = nullptr;
  // step table of the chained encoder (enc_chain.cuh): pinned host staging + device copy, allocated on first use
  EncStep *enc_steps_h = nullptr, *enc_steps_d = nullptr;
  // tables + tile flags of the per-stage dataflow kernel (stage_umma.cuh), one set per HiFi-GAN stage, allocated on first use
  void *stage_tab_h[8] = {}, *stage_tab_d[8] = {};
  unsigned int* stage_flags[8] = {};
  size_t stage_flags_cap[8] = {};
  // decode state
  int *idx_d = nullptr, idx_ld = 0;
  float* taps_d = nullptr;
  int tq_ld = 0;
  float* state_d = nullptr;
};

static thread_local std::string g_err;
static int fail(const std::exception& e) {
  g_err = e.what();
  return 1;
}
#define B2_API_BEGIN try {
#define B2_API_END \
  }                \
  catch (const std::exception& e) { return fail(e); } \
  return 0;

static bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct Timer {
  b200tts_call* c;
  explicit Timer(b200tts_call* c_) : c(c_) { B2_CUDA(cudaEventRecord(c->ev0, c->stream)); }
  void stop_and_sync() {
    B2_CUDA(cudaEventRecord(c->ev1, c->stream));
    B2_CUDA(cudaEventSynchronize(c->ev1));
    B2_CUDA(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  }
};

// bf16 plane image of an activation tensor (conv_umma.cuh): [nsplit][kc][rows][8] per batch item
struct Planes {
  void* ptr = nullptr;
  long long bs = 0;  // bytes per batch item
  int kc = 0, rows = 0, nsplit = 2;
};
// whole 512-row tiles + tail (the fused ResBlock kernel's last tile window reaches up to ~1100 rows past the end: rbpair_umma.cu)
static int plane_rows(int Lin_max, int u) { return 40 + (Lin_max + 511) / 512 * 512 * u + 1280; }
static size_t plane_item_bytes(int channels, int rows, int nsplit) { return (size_t)nsplit * ((channels + 31) / 32 * 4) * rows * 16; }
static Planes take_planes(Arena& a, int batch, int channels, int rows, int nsplit) {
  Planes pl;
  pl.kc = (channels + 31) / 32 * 4;
  pl.rows = rows;
  pl.nsplit = nsplit;
  pl.bs = (long long)plane_item_bytes(channels, rows, nsplit);
  pl.ptr = a.take<unsigned char>((size_t)batch * pl.bs);
  return pl;
}
static void set_in(ConvParams& p, const Planes& pl) {
  p.xp = pl.ptr; p.xp_bs = pl.bs; p.xp_kc = pl.kc; p.xp_rows = pl.rows;
}
static void set_out(ConvParams& p, const Planes& pl, float slope, const Planes* second = nullptr) {
  p.yp = pl.ptr; p.yp_bs = pl.bs; p.yp_kc = pl.kc; p.yp_rows = pl.rows; p.yp_nsplit = pl.nsplit; p.yp_slope = slope;
  if (second) p.yp2 = second->ptr;
}

// ONE debug switch: B200TTS_DEBUG is a comma-separated list of flags (read once per process, thread-safe since C++11 —
// calls arrive from many host threads).  Nothing here is needed in production; the defaults are the product path.
//   simt          fp32 SIMT convs + SIMT attention everywhere (the in-library cross-check of the tensor-core path)
//   attn_simt     SIMT attention only
//   no_flow_chain flow decoder as one launch per conv instead of the persistent kernel (flow_chain.cuh)
//   no_enc_chain  text encoder as one launch per layer instead of the persistent kernel (enc_chain.cuh)
//   no_rbpair     HiFi-GAN ResBlock convs of the narrow stages as two launches instead of the fused kernel (rbpair_umma.cuh)
//   stage_trace   per-item timeline of one CTA of every stage_umma launch on stderr (single-threaded use)
//   no_stage_pair the 32-channel HiFi-GAN stage as one item per conv instead of one per ResBlock step (stage_pair_kernel)
//   no_stage_stripe  stage kernels: every CTA gets a contiguous tile range of ONE branch per level instead of the striped order
//   no_stage      HiFi-GAN ResBlock convs as one launch per conv position instead of one dataflow launch per stage (stage_umma.cuh)
//   chain_trace   per-layer phase stamps of the persistent kernels on stderr
//   trace         per-conv %globaltimer phase trace on stderr (single-threaded use; implies no_flow_chain)
static bool dbg_flag(const char* name) {
  static const std::string flags = [] {
    const char* e = getenv("B200TTS_DEBUG");
    return std::string(",") + (e ? e : "") + ",";
  }();
  return flags.find(std::string(",") + name + ",") != std::string::npos;
}
static bool use_umma() {
  static const bool on = !dbg_flag("simt");
  return on;
}
static bool conv_trace() {
  static const bool on = dbg_flag("trace");
  return on;
}
static bool flow_chain_enabled() {
  static const bool on = !dbg_flag("no_flow_chain") && !dbg_flag("trace") && !dbg_flag("simt");
  return on;
}
static bool enc_chain_enabled() {
  static const bool on = !dbg_flag("no_enc_chain") && !dbg_flag("trace") && !dbg_flag("simt") && !dbg_flag("attn_simt");
  return on;
}

static void conv(b200tts_call* c, ConvParams& p, const ConvW& w, int batch, int Lmax, cudaStream_t s) {
  p.w = w.w;
  p.bias = w.b;
  p.Cin = w.cin;
  p.Cout = w.cout;
  p.K = w.k;
  if (w.wu && use_umma() && p.n_phase == 1) {
    p.wu = w.wu;
    p.um_nt = w.nt;
    p.um_nsplit = w.nsplit;
    if (conv_trace()) {  // debug only (single-threaded use): per-phase timestamps of CTA (0,0,0), one line per conv
      static unsigned long long* dbg = nullptr;
      static unsigned long long last_end = 0;
      if (!dbg) B2_CUDA(cudaMallocManaged(&dbg, 64 * sizeof(unsigned long long)));
      B2_CUDA(cudaDeviceSynchronize());
      memset(dbg, 0, 64 * sizeof(unsigned long long));
      p.dbg = dbg;
      c->launches += launch_conv_umma(p, batch, Lmax, s);
      B2_CUDA(cudaDeviceSynchronize());
      fprintf(stderr, "[conv] Cin %4d Cout %4d K %2d L %6d nt %3d | gap %6.1f prologue %5.1f A0 %5.1f mma0 %5.1f issued %6.1f Alast %6.1f acc %6.1f epi %5.1f end %5.1f | total %6.1f us | epi: ld %4.1f math %4.1f st %4.1f\n",
              p.Cin, p.Cout, p.K, Lmax, p.um_nt, last_end ? (dbg[0] - last_end) * 1e-3 : 0.0, (dbg[1] - dbg[0]) * 1e-3, (dbg[2] - dbg[1]) * 1e-3,
              (dbg[3] - dbg[1]) * 1e-3, (dbg[4] - dbg[1]) * 1e-3, (dbg[5] - dbg[1]) * 1e-3, (dbg[6] - dbg[1]) * 1e-3, (dbg[7] - dbg[6]) * 1e-3,
              (dbg[8] - dbg[7]) * 1e-3, (dbg[8] - dbg[0]) * 1e-3, dbg[9] ? (dbg[9] - dbg[6]) * 1e-3 : 0.0,
              dbg[10] ? (dbg[10] - dbg[9]) * 1e-3 : 0.0, dbg[11] ? (dbg[11] - dbg[10]) * 1e-3 : 0.0);
      if (dbg[16]) {  // several tiles per CTA: per tile  mainloop start..end | epilogue start..end  (us since CTA start)
        fprintf(stderr, "        tiles:");
        for (int i = 0; i < 8 && dbg[16 + 4 * i]; ++i)
          fprintf(stderr, "  [mma %.1f-%.1f epi %.1f-%.1f]", (dbg[16 + 4 * i] - dbg[0]) * 1e-3, (dbg[17 + 4 * i] - dbg[0]) * 1e-3,
                  (dbg[18 + 4 * i] - dbg[0]) * 1e-3, (dbg[19 + 4 * i] - dbg[0]) * 1e-3);
        fprintf(stderr, "\n");
      }
      last_end = dbg[8];
      return;
    }
    c->launches += launch_conv_umma(p, batch, Lmax, s);
  } else {
    B2_CHECK(w.w != nullptr, "layer has no fp32 weights for the SIMT path");
    B2_CHECK(p.y != nullptr || p.epi == EPI_COUPLING, "SIMT path needs an fp32 output buffer");
    c->launches += launch_conv1d(p, batch, Lmax, s);
  }
}

// ------------------------------------------------------------------------------------------------
// GlowTTS
// ------------------------------------------------------------------------------------------------
static void glow_bind(b200tts_glow* m) {
  const Weights& W = m->W;
  const int32_t* c = W.cfg;
  B2_CHECK(W.n_cfg >= 16, "glow blob config too short");
  m->V = c[0]; m->H = c[1]; m->F = c[2]; m->Fd = c[3]; m->K = c[4]; m->n_heads = c[5]; m->n_layers = c[6];
  m->window = c[7]; m->mel = c[8]; m->n_blocks = c[9]; m->kd = c[10]; m->dil_rate = c[11];
  m->n_block_layers = c[12]; m->n_sqz = c[13]; m->n_split = c[14]; m->Hd = c[15];
  B2_CHECK(m->n_split == 4 && m->n_sqz == 2, "only n_split=4, n_sqz=2 voices are supported");
  B2_CHECK(m->n_block_layers <= 8, "too many WN layers");
  B2_CHECK(m->H % 8 == 0 && m->Hd % 8 == 0 && m->mel % 8 == 0 && m->F % 8 == 0 && m->Fd % 8 == 0, "channel counts must be multiples of 8");
  const int H = m->H, mel = m->mel, C = mel * m->n_sqz, Hd = m->Hd;
  m->emb = W.get("emb", (size_t)m->V * H);
  for (int i = 0; i < 3; ++i) {
    std::string s = "pre.conv" + std::to_string(i);
    m->pre_conv[i] = conv_w(W, s, H, H, 5);
    m->pre_lng[i] = W.get("pre.ln" + std::to_string(i) + ".g", H);
    m->pre_lnb[i] = W.get("pre.ln" + std::to_string(i) + ".b", H);
  }
  m->pre_proj = conv_w(W, "pre.proj", H, H, 1);
  m->enc.resize(m->n_layers);
  const int dk = H / m->n_heads, nrel = 2 * m->window + 1;
  for (int i = 0; i < m->n_layers; ++i) {
    std::string s = "enc" + std::to_string(i);
    GlowEnc& e = m->enc[i];
    e.qkv = conv_w(W, s + ".qkv", H, 3 * H, 1);
    e.o = conv_w(W, s + ".o", H, H, 1);
    e.ffn1 = conv_w(W, s + ".ffn1", H, m->F, m->K);
    e.ffn2 = conv_w(W, s + ".ffn2", m->F, H, m->K);
    e.rel_k = W.get(s + ".rel_k", (size_t)nrel * dk);
    e.rel_v = W.get(s + ".rel_v", (size_t)nrel * dk);
    if (W.has(s + ".rel_k_img") && dk == 96) {
      e.relk_img = W.get(s + ".rel_k_img", 3 * 12 * 16 * 8 / 2);
      e.relv_img = W.get(s + ".rel_v_img", 3 * 12 * 16 * 8 / 2);
    }
    e.ln1g = W.get(s + ".ln1.g", H); e.ln1b = W.get(s + ".ln1.b", H);
    e.ln2g = W.get(s + ".ln2.g", H); e.ln2b = W.get(s + ".ln2.b", H);
  }
  m->proj_m = conv_w(W, "proj_m", H, mel, 1);
  m->dp1 = conv_w(W, "dp.conv1", H, m->Fd, m->K);
  m->dp2 = conv_w(W, "dp.conv2", m->Fd, m->Fd, m->K);
  m->dp_ln1g = W.get("dp.ln1.g", m->Fd); m->dp_ln1b = W.get("dp.ln1.b", m->Fd);
  m->dp_ln2g = W.get("dp.ln2.g", m->Fd); m->dp_ln2b = W.get("dp.ln2.b", m->Fd);
  m->dp_pw = W.get("dp.proj.w", m->Fd);
  m->dp_pb = W.get("dp.proj.b", 1);
  m->flow.resize(m->n_blocks);
  for (int b = 0; b < m->n_blocks; ++b) {
    std::string s = "flow" + std::to_string(b);
    GlowFlow& f = m->flow[b];
    f.start = conv_w(W, s + ".start", C / 2, Hd, 1);
    f.end = conv_w(W, s + ".end", Hd, C, 1);
    for (int l = 0; l < m->n_block_layers; ++l) {
      f.in[l] = conv_w(W, s + ".in" + std::to_string(l), Hd, 2 * Hd, m->kd);
      f.rs[l] = conv_w(W, s + ".rs" + std::to_string(l), Hd, l < m->n_block_layers - 1 ? 2 * Hd : Hd, 1);
    }
    f.inv = W.get(s + ".inv", 16);
    f.an_bias = W.get(s + ".an_bias", C);
    f.an_scale = W.get(s + ".an_scale", C);
  }
  // flow blocks for the persistent decoder kernel, in execution order (FlowSpecDecoder runs reversed(flows), models.py:198)
  bool chain_ok = m->n_blocks > 0 && C % 32 == 0 && Hd % 32 == 0;
  for (const GlowFlow& f : m->flow) {
    chain_ok = chain_ok && f.start.wu && f.end.wu && f.start.nsplit == 2 && f.end.nsplit == 2 && f.end.nt == m->flow[0].end.nt;
    for (int l = 0; l < m->n_block_layers; ++l)
      chain_ok = chain_ok && f.in[l].wu && f.rs[l].wu && f.in[l].nsplit == 2 && f.rs[l].nsplit == 2 &&
                 f.in[l].nt == f.start.nt && f.rs[l].nt == f.start.nt;
  }
  if (chain_ok) chain_ok = flow_chain_geometry_ok(C / 2, Hd, m->kd, m->dil_rate, m->n_block_layers, m->flow[0].start.nt, m->flow[0].end.nt);
  if (chain_ok) {
    std::vector<FlowBlockDev> h(m->n_blocks);
    for (int e = 0; e < m->n_blocks; ++e) {
      const int blk = m->n_blocks - 1 - e;
      const GlowFlow& f = m->flow[blk];
      FlowBlockDev& d = h[e];
      memset(&d, 0, sizeof(d));
      int st = 0;
      d.w[st] = f.start.wu; d.b[st++] = f.start.b;
      for (int l = 0; l < m->n_block_layers; ++l) {
        d.w[st] = f.in[l].wu; d.b[st++] = f.in[l].b;
        d.w[st] = f.rs[l].wu; d.b[st++] = f.rs[l].b;
      }
      d.w[st] = f.end.wu; d.b[st++] = f.end.b;
      d.inv = f.inv; d.an_bias = f.an_bias; d.an_scale = f.an_scale;
      d.tap_index = blk;
    }
    B2_CUDA(cudaMalloc((void**)&m->flow_dev, sizeof(FlowBlockDev) * h.size()));
    B2_CUDA(cudaMemcpy(m->flow_dev, h.data(), sizeof(FlowBlockDev) * h.size(), cudaMemcpyHostToDevice));
    m->flow_nt_end = m->flow[0].end.nt;
  }
}

static const int* lens_for_ln(const b200tts_call* c) { return c->xlen_d; }

static void glow_encode_impl(const b200tts_glow* m, b200tts_call* c, const int64_t* ids, const int32_t* x_len,
                             int batch, int max_tx, float length_scale, int32_t* y_len_out) {
  B2_CHECK(batch > 0 && max_tx > 0, "empty batch");
  B2_CHECK(m->W.device == c->device, "model and call live on different devices");
  B2_CUDA(cudaSetDevice(c->device));
  // the state of an earlier encode dies here: a failed encode must not leave decode / tts_finish a half-written one
  c->enc_model = nullptr;
  c->idx_d = nullptr;
  cudaStream_t s = c->stream;
  const int H = m->H, F = m->F, Fd = m->Fd, mel = m->mel;
  const int ld = round_up(max_tx, 32);
  const long long B = batch;
  // ---- workspace
  size_t need = 0;
  need += Arena::pad(B * max_tx, 8) + 4 * Arena::pad(B, 4) + 2 * Arena::pad(B * max_tx, 4) + Arena::pad(B * max_tx, 4);
  const size_t act = (size_t)B * ld;
  need += Arena::pad(act * H, 4) * 4 + Arena::pad(act * 3 * H, 4) + Arena::pad(act * F, 4) +
          Arena::pad(act * Fd, 4) * 2 + Arena::pad(act * mel, 4);
  // bf16 operand images (3-way split: the encoder feeds the bit-exact alignment) written by LN / conv epilogues
  const bool um = use_umma() && H % 32 == 0 && F % 32 == 0 && Fd % 32 == 0;
  const int prow = plane_rows(max_tx, 1);
  if (um) need += 4 * Arena::pad(B * plane_item_bytes(H, prow, 3), 1) + Arena::pad(B * plane_item_bytes(F, prow, 3), 1) +
                  Arena::pad(B * plane_item_bytes(Fd, prow, 3), 1) + Arena::pad(B * plane_item_bytes(3 * H, prow, 3), 1);
  // tcgen05 attention needs the q/k/v image and handles up to 128 phonemes per item; longer batches use the SIMT kernel
  bool chain = false;
  const bool um_att = um && attn_umma_supported(H, m->n_heads, m->window, max_tx) && !dbg_flag("attn_simt");
  // the persistent encoder kernel (enc_chain.cuh): every conv on 3-way split images, <= 128 phonemes per item
  const int n_enc_steps = 8 + 7 * m->n_layers + 4;
  {
    bool ok = um_att && enc_chain_enabled() && enc_chain_fits(max_tx) && n_enc_steps <= EC_MAX_STEPS && m->K <= 5;
    auto good = [&](const ConvW& w) { return w.wu != nullptr && w.nsplit == 3 && w.nt <= 64; };
    for (int i = 0; i < 3; ++i) ok = ok && good(m->pre_conv[i]);
    ok = ok && good(m->pre_proj) && good(m->proj_m) && good(m->dp1) && good(m->dp2);
    ok = ok && (H / 32) % 3 == 0 && (F / 32) % 4 == 0 && (Fd / 32) % 2 == 0 && H <= 256 && Fd <= 256;
    ok = ok && (m->dp1.cout + m->dp1.nt - 1) / m->dp1.nt * 2 + (m->proj_m.cout + m->proj_m.nt - 1) / m->proj_m.nt <= EC_YT;
    for (const GlowEnc& e : m->enc) ok = ok && good(e.qkv) && good(e.o) && good(e.ffn1) && good(e.ffn2) && e.relk_img && e.relv_img;
    chain = ok;
  }
  const int part_c = std::max(std::max(H, Fd), 8);
  if (chain) need += Arena::pad((size_t)4 * B * part_c * ld, 4) + Arena::pad((size_t)B * n_enc_steps, 4);
  c->enc.reserve(need + 4096);
  long long* ids_d = c->enc.take<long long>(B * max_tx);
  c->xlen_d = c->enc.take<int>(B);
  c->ylen_d = c->enc.take<int>(B);
  c->sqlen_d = c->enc.take<int>(B);
  int* bad_id_d = c->enc.take<int>(1);
  c->wceil_d = c->enc.take<int>(B * max_tx);
  c->cum_d = c->enc.take<int>(B * max_tx);
  c->logw_d = c->enc.take<float>(B * max_tx);
  float* X = c->enc.take<float>(act * H);
  float* A = c->enc.take<float>(act * H);
  float* Bf = c->enc.take<float>(act * H);
  float* ATT = c->enc.take<float>(act * H);
  float* QKV = c->enc.take<float>(act * 3 * H);
  float* FF = c->enc.take<float>(act * F);
  float* D1 = c->enc.take<float>(act * Fd);
  float* D2 = c->enc.take<float>(act * Fd);
  c->xm_d = c->enc.take<float>(act * mel);
  Planes X_p, P_p[2], FF_p, D1_p, QKV_p, ATT_p;
  if (um) {
    QKV_p = take_planes(c->enc, batch, 3 * H, prow, 3);
    ATT_p = take_planes(c->enc, batch, H, prow, 3);
    X_p = take_planes(c->enc, batch, H, prow, 3);
    P_p[0] = take_planes(c->enc, batch, H, prow, 3);
    P_p[1] = take_planes(c->enc, batch, H, prow, 3);
    FF_p = take_planes(c->enc, batch, F, prow, 3);
    D1_p = take_planes(c->enc, batch, Fd, prow, 3);
  }
  auto ln = [&](const float* x, const float* res, const float* g, const float* bta, float* y, int C, int pre_relu, int post_relu,
                const Planes* pl) {
    launch_layernorm(x, res, g, bta, y, (long long)C * ld, ld, C, lens_for_ln(c), batch, max_tx, pre_relu, post_relu, s,
                     pl ? pl->ptr : nullptr, pl ? pl->bs : 0, pl ? pl->kc : 0, pl ? pl->rows : 0, pl ? pl->nsplit : 3);
    c->launches++;
  };

  Timer tm(c);
  B2_CUDA(cudaMemcpyAsync(ids_d, ids, sizeof(long long) * B * max_tx, cudaMemcpyDefault, s));
  B2_CUDA(cudaMemcpyAsync(c->xlen_d, x_len, sizeof(int) * B, cudaMemcpyDefault, s));
  const int* lens = c->xlen_d;

  auto mk = [&](const float* x, int cin_rows, float* y, int cout_rows) {
    ConvParams p;
    p.x = x; p.x_bs = (long long)cin_rows * ld; p.x_ld = ld;
    p.y = y; p.y_bs = (long long)cout_rows * ld; p.y_ld = ld;
    p.lens = lens; p.len_mul = 1; p.Lin = max_tx;
    return p;
  };
  B2_CUDA(cudaMemsetAsync(bad_id_d, 0, sizeof(int), s));
  if (chain) {
    // ---- the whole encoder as one persistent launch: embedding image -> step table -> enc_chain_kernel -> duration head
    float* PART = c->enc.take<float>((size_t)4 * B * part_c * ld);
    unsigned int* counters = c->enc.take<unsigned int>((size_t)B * n_enc_steps);
    launch_embed_image(ids_d, lens, batch, max_tx, m->emb, m->V, H, X, (long long)H * ld, ld, X_p.ptr, X_p.bs, X_p.kc, X_p.rows,
                       bad_id_d, s);
    c->launches++;
    if (!c->enc_steps_h) {
      B2_CUDA(cudaMallocHost((void**)&c->enc_steps_h, sizeof(EncStep) * EC_MAX_STEPS));
      B2_CUDA(cudaMalloc((void**)&c->enc_steps_d, sizeof(EncStep) * EC_MAX_STEPS));
    }
    EncStep* st = c->enc_steps_h;
    memset(st, 0, sizeof(EncStep) * EC_MAX_STEPS);
    int ns = 0;
    const long long part_stride = (long long)B * part_c * ld;  // floats between K-split partial buffers
    auto img = [](const Planes& pl) { return (unsigned char*)pl.ptr; };
    int planes = c->enc_bf16 ? 1 : 3;  // activation planes of the prenet / encoder-layer GEMMs; the duration head keeps 3
    // conv -> raw fp32 partials (K-split `ksplit`) at CTA rows y0..
    auto conv_partial = [&](EncConvOp& o, const ConvW& w, const Planes& in, int pad, int ksplit, int y0) {
      o.w = w.wu; o.bias = nullptr; o.Cin = w.cin; o.Cout = w.cout; o.K = w.k; o.pad = pad; o.NT = w.nt;
      o.ntiles = (w.cout + w.nt - 1) / w.nt; o.ksplit = ksplit; o.y0 = y0; o.act = ACT_NONE; o.planes = planes;
      o.xp = img(in); o.xp_bs = in.bs; o.xp_kc = in.kc;
      o.y = PART; o.y_bs = (long long)part_c * ld; o.part_stride = part_stride; o.y_ld = ld;
    };
    // conv -> operand image (+ bias, activation)
    auto conv_image = [&](EncConvOp& o, const ConvW& w, const Planes& in, int pad, int act, const Planes& out, int y0) {
      o.w = w.wu; o.bias = w.b; o.Cin = w.cin; o.Cout = w.cout; o.K = w.k; o.pad = pad; o.NT = w.nt;
      o.ntiles = (w.cout + w.nt - 1) / w.nt; o.ksplit = 1; o.y0 = y0; o.act = act; o.planes = planes;
      o.xp = img(in); o.xp_bs = in.bs; o.xp_kc = in.kc;
      o.yp = img(out); o.yp_bs = out.bs; o.yp_kc = out.kc;
    };
    auto reduce = [&](int parts, int C, const float* bias, const float* res, const float* g, const float* bta, int pre_relu,
                      int post_relu, float* y, const Planes* out) {  // res / y: [batch][C][ld]
      EncStep& e = st[ns++];
      e.type = ES_REDUCE;
      EncReduceOp& r = e.red;
      r.x = PART; r.x_bs = (long long)part_c * ld; r.part_stride = part_stride; r.ld = ld; r.parts = parts; r.C = C;
      r.pre_relu = pre_relu; r.post_relu = post_relu; r.bias = bias; r.res = res; r.res_bs = (long long)C * ld;
      r.gamma = g; r.beta = bta; r.y = y; r.y_bs = (long long)C * ld;
      if (out) { r.yp = img(*out); r.yp_bs = out->bs; r.yp_kc = out->kc; }
    };
    // prenet: 3 x (conv k5 -> LN -> ReLU), 1x1 proj + residual              layers.py:73-80
    for (int i = 0; i < 3; ++i) {
      EncStep& e = st[ns++];
      e.type = ES_CONV; e.n_conv = 1;
      conv_partial(e.conv[0], m->pre_conv[i], i == 0 ? X_p : P_p[(i - 1) & 1], 2, 3, 0);
      reduce(3, H, m->pre_conv[i].b, nullptr, m->pre_lng[i], m->pre_lnb[i], 0, 1, nullptr, &P_p[i & 1]);
    }
    {
      EncStep& e = st[ns++];
      e.type = ES_CONV; e.n_conv = 1;
      conv_partial(e.conv[0], m->pre_proj, P_p[0], 0, 3, 0);
      reduce(3, H, m->pre_proj.b, X, nullptr, nullptr, 0, 0, X, &X_p);
    }
    // encoder layers                                                          attentions.py:62-74
    for (int i = 0; i < m->n_layers; ++i) {
      const GlowEnc& L = m->enc[i];
      {
        EncStep& e = st[ns++];
        e.type = ES_CONV; e.n_conv = 1;
        conv_image(e.conv[0], L.qkv, X_p, 0, ACT_NONE, QKV_p, 0);
      }
      {
        EncStep& e = st[ns++];
        e.type = ES_ATTN;
        EncAttnOp& a = e.att;
        a.qkv_p = img(QKV_p); a.qkv_bs = QKV_p.bs; a.qkv_kc = QKV_p.kc;
        a.relk_img = (const unsigned char*)L.relk_img; a.relv_img = (const unsigned char*)L.relv_img;
        a.window = m->window; a.H = H; a.n_heads = m->n_heads; a.y0 = 0; a.planes = planes;
        a.out_p = img(ATT_p); a.out_bs = ATT_p.bs; a.out_kc = ATT_p.kc;
      }
      {
        EncStep& e = st[ns++];
        e.type = ES_CONV; e.n_conv = 1;
        conv_partial(e.conv[0], L.o, ATT_p, 0, 3, m->n_heads);  // not on the attention CTAs: their weight ring was on hold
        reduce(3, H, L.o.b, X, L.ln1g, L.ln1b, 0, 0, X, &X_p);
      }
      {
        EncStep& e = st[ns++];
        e.type = ES_CONV; e.n_conv = 1;
        conv_image(e.conv[0], L.ffn1, X_p, m->K / 2, ACT_RELU, FF_p, 0);
      }
      {
        EncStep& e = st[ns++];
        e.type = ES_CONV; e.n_conv = 1;
        conv_partial(e.conv[0], L.ffn2, FF_p, m->K / 2, 4, 0);
        reduce(4, H, L.ffn2.b, X, L.ln2g, L.ln2b, 0, 0, X, &X_p);
      }
    }
    // proj_m and the duration predictor                                        models.py:133, 39-49
    planes = 3;  // fp32-grade whatever the mode: the durations decide the alignment
    {
      EncStep& e = st[ns++];
      e.type = ES_CONV; e.n_conv = 2;
      conv_partial(e.conv[0], m->dp1, X_p, m->K / 2, 2, 0);
      EncConvOp& pm = e.conv[1];
      pm.w = m->proj_m.wu; pm.bias = m->proj_m.b; pm.Cin = H; pm.Cout = mel; pm.K = 1; pm.pad = 0; pm.NT = m->proj_m.nt;
      pm.ntiles = (mel + pm.NT - 1) / pm.NT; pm.ksplit = 1; pm.y0 = e.conv[0].ntiles * 2; pm.act = ACT_NONE; pm.planes = 3;
      pm.xp = img(X_p); pm.xp_bs = X_p.bs; pm.xp_kc = X_p.kc;
      pm.y = c->xm_d; pm.y_bs = (long long)mel * ld; pm.y_ld = ld;
      reduce(2, Fd, m->dp1.b, nullptr, m->dp_ln1g, m->dp_ln1b, 1, 0, nullptr, &D1_p);
    }
    {
      EncStep& e = st[ns++];
      e.type = ES_CONV; e.n_conv = 1;
      conv_partial(e.conv[0], m->dp2, D1_p, m->K / 2, 2, 0);
      reduce(2, Fd, m->dp2.b, nullptr, m->dp_ln2g, m->dp_ln2b, 1, 0, D2, nullptr);
    }
    B2_CHECK(ns == n_enc_steps, "encoder step table has an unexpected length");
    B2_CUDA(cudaMemcpyAsync(c->enc_steps_d, st, sizeof(EncStep) * ns, cudaMemcpyHostToDevice, s));
    EncChainParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.steps = c->enc_steps_d; fp.n_steps = ns; fp.batch = batch; fp.rows = prow; fp.lens = lens; fp.counter = counters;
    static const bool trace = dbg_flag("chain_trace");
    if (trace) {
      static unsigned long long* tr = nullptr;
      if (!tr) B2_CUDA(cudaMallocManaged(&tr, sizeof(unsigned long long) * 8 * EC_YT * EC_MAX_STEPS));
      B2_CUDA(cudaDeviceSynchronize());
      memset(tr, 0, sizeof(unsigned long long) * 8 * EC_YT * EC_MAX_STEPS);
      fp.trace = tr;
      c->launches += launch_enc_chain(st, fp, max_tx, s);
      B2_CUDA(cudaDeviceSynchronize());
      unsigned long long t_start = ~0ull, t_end = 0;
      for (int i = 0; i < 8 * EC_YT * ns; ++i)
        if (tr[i]) { t_start = std::min(t_start, tr[i]); t_end = std::max(t_end, tr[i]); }
      for (int L = 0; L < ns && L < 16; ++L) {  // prenet + first layer
        fprintf(stderr, "[enc] S%2d type %d:", L, st[L].type);
        for (int y = 0; y < EC_YT; ++y) {
          const unsigned long long* e = tr + ((size_t)L * EC_YT + y) * 8;
          if (!e[5]) continue;
          fprintf(stderr, " | y%d start %6.2f A %5.2f mma %5.2f acc %5.2f epi %5.2f pub %5.2f", y, ((e[0] ? e[0] : e[4]) - t_start) * 1e-3,
                  e[0] && e[1] ? (e[1] - e[0]) * 1e-3 : 0.0, e[3] && e[1] ? (e[3] - e[1]) * 1e-3 : 0.0, e[3] ? (e[4] - e[3]) * 1e-3 : 0.0,
                  (e[5] - e[4]) * 1e-3, (e[6] - e[5]) * 1e-3);
        }
        fprintf(stderr, "\n");
      }
      fprintf(stderr, "[enc] %d steps, %.1f us from first stamp to last publish (%.2f us per step)\n", ns, (t_end - t_start) * 1e-3,
              (t_end - t_start) * 1e-3 / ns);
    } else {
      c->launches += launch_enc_chain(st, fp, max_tx, s);
    }
    launch_duration(D2, (long long)Fd * ld, ld, Fd, m->dp_pw, m->dp_pb, lens, batch, max_tx, length_scale, m->n_sqz,
                    c->logw_d, c->wceil_d, c->cum_d, c->ylen_d, s);
    c->launches += 1;
  } else {
  // embedding * sqrt(H)                                         models.py:119-121
  launch_embed(ids_d, lens, batch, max_tx, m->emb, m->V, H, X, (long long)H * ld, ld, bad_id_d, s);
  c->launches++;
  // prenet: 3 x (conv k5 -> LN -> ReLU), 1x1 proj + residual   layers.py:73-80
  {
    const float* h = X;
    float* bufs[2] = {A, Bf};
    for (int i = 0; i < 3; ++i) {
      float* o = bufs[i & 1];
      ConvParams p = mk(h, H, o, H);
      p.pad_left = 2;
      if (um && i > 0) set_in(p, P_p[(i - 1) & 1]);  // the embedding itself is gathered from fp32
      conv(c, p, m->pre_conv[i], batch, max_tx, s);
      ln(o, nullptr, m->pre_lng[i], m->pre_lnb[i], o, H, 0, 1, um ? &P_p[i & 1] : nullptr);
      h = o;
    }
    ConvParams p = mk(h, H, X, H);
    p.res = X; p.res_bs = (long long)H * ld; p.res_ld = ld;
    if (um) {
      set_in(p, P_p[0]);
      set_out(p, X_p, 1.f);
    }
    conv(c, p, m->pre_proj, batch, max_tx, s);
  }
  // encoder layers                                              attentions.py:62-74
  for (int i = 0; i < m->n_layers; ++i) {
    const GlowEnc& e = m->enc[i];
    ConvParams p = mk(X, H, QKV, 3 * H);
    if (um) set_in(p, X_p);
    if (um_att && e.relk_img) {
      set_out(p, QKV_p, 1.f);
      p.y = nullptr;  // q, k, v are only ever MMA operands of the attention kernel
    }
    conv(c, p, e.qkv, batch, max_tx, s);
    if (um_att && e.relk_img) {
      launch_attn_umma(QKV_p.ptr, QKV_p.bs, QKV_p.kc, QKV_p.rows, e.relk_img, e.relv_img, m->window, H, m->n_heads, lens, batch,
                       ATT_p.ptr, ATT_p.bs, ATT_p.kc, ATT_p.rows, nullptr, 0, 0, s, max_tx);
    } else {
      launch_rel_attention(QKV, (long long)3 * H * ld, ld, H, m->n_heads, e.rel_k, e.rel_v, m->window, lens, batch,
                           max_tx, ATT, (long long)H * ld, ld, s);
    }
    c->launches++;
    p = mk(ATT, H, A, H);
    if (um_att && e.relk_img) set_in(p, ATT_p);
    conv(c, p, e.o, batch, max_tx, s);
    ln(X, A, e.ln1g, e.ln1b, X, H, 0, 0, um ? &X_p : nullptr);
    p = mk(X, H, FF, F);
    p.pad_left = m->K / 2; p.act = ACT_RELU;
    if (um && e.ffn1.wu && e.ffn2.wu) {
      set_in(p, X_p);
      set_out(p, FF_p, 1.f);
      p.y = nullptr;  // the FFN hidden activation is only ever read as an MMA operand
    }
    conv(c, p, e.ffn1, batch, max_tx, s);
    p = mk(FF, F, A, H);
    p.pad_left = m->K / 2;
    if (um && e.ffn1.wu && e.ffn2.wu) set_in(p, FF_p);
    conv(c, p, e.ffn2, batch, max_tx, s);
    ln(X, A, e.ln2g, e.ln2b, X, H, 0, 0, um ? &X_p : nullptr);
  }
  // proj_m                                                      models.py:133
  {
    ConvParams p = mk(X, H, c->xm_d, mel);
    if (um) set_in(p, X_p);
    conv(c, p, m->proj_m, batch, max_tx, s);
  }
  // duration predictor                                          models.py:39-49
  {
    ConvParams p = mk(X, H, D1, Fd);
    p.pad_left = m->K / 2;
    if (um) set_in(p, X_p);
    conv(c, p, m->dp1, batch, max_tx, s);
    ln(D1, nullptr, m->dp_ln1g, m->dp_ln1b, D1, Fd, 1, 0, um ? &D1_p : nullptr);
    p = mk(D1, Fd, D2, Fd);
    p.pad_left = m->K / 2;
    if (um) set_in(p, D1_p);
    conv(c, p, m->dp2, batch, max_tx, s);
    ln(D2, nullptr, m->dp_ln2g, m->dp_ln2b, D2, Fd, 1, 0, nullptr);
    launch_duration(D2, (long long)Fd * ld, ld, Fd, m->dp_pw, m->dp_pb, lens, batch, max_tx, length_scale, m->n_sqz,
                    c->logw_d, c->wceil_d, c->cum_d, c->ylen_d, s);
    c->launches += 1;
  }
  }  // classic (one launch per layer) path
  c->y_len_h.resize(batch);
  B2_CUDA(cudaMemcpyAsync(c->y_len_h.data(), c->ylen_d, sizeof(int) * B, cudaMemcpyDeviceToHost, s));
  int bad_id = 0;
  B2_CUDA(cudaMemcpyAsync(&bad_id, bad_id_d, sizeof(int), cudaMemcpyDeviceToHost, s));
  tm.stop_and_sync();
  B2_CHECK(bad_id == 0, "phoneme id out of range for this voice (0 <= id < num_symbols)");
  c->enc_model = m;
  c->batch = batch;
  c->max_tx = max_tx;
  c->tx_ld = ld;
  c->ty_max = 0;
  for (int b = 0; b < batch; ++b) {
    c->ty_max = std::max(c->ty_max, c->y_len_h[b]);
    if (y_len_out) y_len_out[b] = c->y_len_h[b];
  }
}

// Runs expand+noise+flow on c->stream.  Leaves the un-squeezed result in the squeezed state buffer
// (c->state_d) and returns it; the caller launches the unsqueeze epilogue it needs.
static void glow_flow_impl(const b200tts_glow* m, b200tts_call* c, float noise_scale, const float* eps, int eps_ld,
                           uint64_t seed, size_t extra_bytes) {
  B2_CHECK(c->enc_model == m, "b200tts_glow_decode must follow b200tts_glow_encode with the same model and call");
  cudaStream_t s = c->stream;
  const int batch = c->batch, mel = m->mel, C = mel * m->n_sqz, Hd = m->Hd;
  const long long B = batch;
  const int Ty = c->ty_max;
  const int Tq = Ty / m->n_sqz;
  const int ldq = round_up(std::max(Tq, 1), 32);
  c->tq_ld = ldq;
  c->idx_ld = round_up(std::max(Ty, 1), 32);
  const bool eps_host = eps && !is_device_ptr(eps);
  size_t need = Arena::pad(B * c->idx_ld, 4) + Arena::pad((size_t)B * C * ldq, 4) + 3 * Arena::pad((size_t)B * Hd * ldq, 4);
  const int prow = plane_rows(Tq, 1);
  const bool um = use_umma() && C % 32 == 0 && Hd % 32 == 0;  // plane images hold whole 32-channel chunks
  if (um) need += Arena::pad(B * plane_item_bytes(C, prow, 2), 1) + 3 * Arena::pad(B * plane_item_bytes(Hd, prow, 2), 1);
  if (eps_host) need += Arena::pad((size_t)B * mel * eps_ld, 4);
  if (c->debug) need += Arena::pad((size_t)m->n_blocks * B * C * ldq, 4);
  // the persistent kernel serves any batch (CTA groups walk the utterances) as long as one utterance's CTA group fits
  const bool chain = um && m->flow_dev != nullptr && flow_chain_enabled() && Tq > 0 && flow_chain_fits(Tq);
  const size_t n_counters = (size_t)B * m->n_blocks * (2 + 2 * m->n_block_layers);
  if (chain) need += Arena::pad(n_counters, 4);
  c->dec.reserve(need + extra_bytes + 4096);
  c->idx_d = c->dec.take<int>(B * c->idx_ld);
  float* S = c->dec.take<float>((size_t)B * C * ldq);
  float* Hh = c->dec.take<float>((size_t)B * Hd * ldq);
  float* ACT = c->dec.take<float>((size_t)B * Hd * ldq);
  float* SK = c->dec.take<float>((size_t)B * Hd * ldq);
  Planes S_p, H_p, A_p, K_p;  // bf16 operand images of the flow state, WN hidden state, gate output, skip sum
  if (um) {
    S_p = take_planes(c->dec, batch, C, prow, 2);
    H_p = take_planes(c->dec, batch, Hd, prow, 2);
    A_p = take_planes(c->dec, batch, Hd, prow, 2);
    K_p = take_planes(c->dec, batch, Hd, prow, 2);
  }
  const float* eps_d = eps;
  if (eps_host) {
    B2_CHECK(eps_ld >= Ty, "eps_ld smaller than the longest item");
    float* e = c->dec.take<float>((size_t)B * mel * eps_ld);
    B2_CUDA(cudaMemcpyAsync(e, eps, sizeof(float) * B * mel * eps_ld, cudaMemcpyHostToDevice, s));
    eps_d = e;
  }
  c->taps_d = c->debug ? c->dec.take<float>((size_t)m->n_blocks * B * C * ldq) : nullptr;
  c->state_d = S;

  launch_align_index(c->cum_d, c->xlen_d, c->ylen_d, batch, c->max_tx, c->idx_d, c->idx_ld, c->sqlen_d, m->n_sqz, s);
  c->launches++;
  if (Ty <= 0) return;
  launch_expand_squeeze(c->xm_d, (long long)mel * c->tx_ld, c->tx_ld, c->idx_d, c->idx_ld, c->ylen_d, eps_d,
                        (long long)mel * eps_ld, eps_ld, noise_scale, seed, mel, m->n_sqz, S, (long long)C * ldq, ldq,
                        batch, Ty, s);
  c->launches++;
  const int* lens = c->sqlen_d;
  auto mk = [&](const float* x, int cin_rows, float* y, int cout_rows) {
    ConvParams p;
    p.x = x; p.x_bs = (long long)cin_rows * ldq; p.x_ld = ldq;
    p.y = y; p.y_bs = (long long)cout_rows * ldq; p.y_ld = ldq;
    p.lens = lens; p.len_mul = 1; p.Lin = Tq;
    return p;
  };
  if (chain) {
    // expand wrote the fp32 state; its operand image feeds the first `start` conv
    launch_mean3_image(S, nullptr, nullptr, (long long)C * ldq, ldq, C, 1.f, 1.f, lens, 1, batch, S_p.ptr, S_p.bs, S_p.kc,
                       S_p.rows, s);
    c->launches++;
    FlowChainParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.blocks = m->flow_dev; fp.n_blocks = m->n_blocks; fp.n_wn = m->n_block_layers; fp.batch = batch;
    fp.half = C / 2; fp.Hd = Hd; fp.K = m->kd; fp.dil_rate = m->dil_rate; fp.nt_end = m->flow_nt_end;
    fp.lens = lens;
    fp.S = S; fp.S_bs = (long long)C * ldq; fp.S_ld = ldq;
    fp.taps = c->taps_d;
    fp.S_p = (unsigned char*)S_p.ptr; fp.H_p = (unsigned char*)H_p.ptr; fp.A_p = (unsigned char*)A_p.ptr; fp.K_p = (unsigned char*)K_p.ptr;
    fp.Sp_bs = S_p.bs; fp.Hp_bs = H_p.bs; fp.Sp_kc = S_p.kc; fp.Hp_kc = H_p.kc; fp.rows = prow;
    fp.counter = c->dec.take<unsigned int>(n_counters);
    static const bool trace = dbg_flag("chain_trace");  // debug: per-layer phase stamps of utterance 0, time tile 0
    const int n_layers = m->n_blocks * (2 + 2 * m->n_block_layers);
    if (trace) {
      static unsigned long long* tr = nullptr;
      if (!tr) B2_CUDA(cudaMallocManaged(&tr, sizeof(unsigned long long) * 8 * 6 * 256));
      B2_CUDA(cudaDeviceSynchronize());
      memset(tr, 0, sizeof(unsigned long long) * 8 * 6 * 256);
      fp.trace = tr;
      c->launches += launch_flow_chain(fp, Tq, s);
      B2_CUDA(cudaDeviceSynchronize());
      const int n_steps = 2 + 2 * m->n_block_layers;
      const unsigned long long t_start = tr[0] ? tr[0] : tr[1];
      for (int L = 0; L < n_layers && L < 2 * n_steps; ++L) {  // first two blocks
        fprintf(stderr, "[chain] L%3d step %d:", L, L % n_steps);
        for (int y = 0; y < 6; ++y) {
          const unsigned long long* e = tr + ((size_t)L * 6 + y) * 8;
          if (!e[4]) { fprintf(stderr, " | y%d idle", y); continue; }
          fprintf(stderr, " | y%d bar %6.2f A %5.2f W %5.2f mma %5.2f acc %5.2f epi %5.2f pub %5.2f", y, (e[0] - t_start) * 1e-3,
                  (e[1] - e[0]) * 1e-3, (e[2] > e[1] ? e[2] - e[1] : 0) * 1e-3, (e[3] - e[2]) * 1e-3, (e[4] - e[3]) * 1e-3, (e[5] - e[4]) * 1e-3,
                  (e[6] - e[5]) * 1e-3);
        }
        fprintf(stderr, "\n");
      }
      for (int c = 0; c < 6; ++c) {  // micro-trace of the MMA warp of CTA (0,0): first res_skip layer and first in_layer, per chunk
        const unsigned long long *e = tr + ((size_t)(200 + c) * 6) * 8, *f = tr + ((size_t)(210 + c) * 6) * 8;
        fprintf(stderr, "[chain]   rs chunk %d: wait A %6.2f..%6.2f issued+committed %6.2f | in chunk %d: wait A %6.2f..%6.2f W %6.2f issued+committed %6.2f\n", c,
                (e[0] - t_start) * 1e-3, (e[1] - t_start) * 1e-3, (e[2] - t_start) * 1e-3, c, (f[0] - t_start) * 1e-3, (f[1] - t_start) * 1e-3,
                (f[2] - t_start) * 1e-3, (f[3] - t_start) * 1e-3);
      }
      unsigned long long t_end = 0;
      for (int i = 0; i < 8 * 6 * n_layers; ++i) t_end = std::max(t_end, tr[i]);
      fprintf(stderr, "[chain] %d layers, %.1f us from first barrier to last publish (%.2f us per layer)\n", n_layers,
              (t_end - t_start) * 1e-3, (t_end - t_start) * 1e-3 / n_layers);
      return;
    }
    c->launches += launch_flow_chain(fp, Tq, s);
    return;
  }
  auto run = [&](ConvParams& p, const ConvW& w) { conv(c, p, w, batch, Tq, s); };
  for (int blk = m->n_blocks - 1; blk >= 0; --blk) {
    const GlowFlow& f = m->flow[blk];
    // CouplingBlock reverse                                     attentions.py:119-142
    ConvParams p = mk(S, C, Hh, Hd);  // start reads x0 = state[:, :C/2]
    if (um) {
      if (blk != m->n_blocks - 1) set_in(p, S_p);  // the first state comes from the expand kernel (fp32 gather)
      set_out(p, H_p, 1.f);
    }
    run(p, f.start);
    for (int l = 0; l < m->n_block_layers; ++l) {  // WN       layers.py:138-162
      int dil = 1;
      for (int j = 0; j < l; ++j) dil *= m->dil_rate;
      p = mk(Hh, Hd, ACT, Hd);
      p.dil = dil; p.pad_left = (m->kd * dil - dil) / 2; p.epi = EPI_GATE;
      if (um && f.in[l].wu) {
        set_in(p, H_p);
        set_out(p, A_p, 1.f);
        p.y = nullptr;  // the gate output is only ever read as an MMA operand
      }
      run(p, f.in[l]);
      p = mk(ACT, Hd, Hh, Hd);
      p.epi = EPI_RESSKIP; p.aux = SK; p.aux_bs = (long long)Hd * ldq; p.aux_ld = ldq;
      p.split = l < m->n_block_layers - 1 ? Hd : 0;
      p.first = l == 0;
      if (um && f.rs[l].wu) {
        set_in(p, A_p);
        set_out(p, H_p, 1.f, &K_p);
      }
      run(p, f.rs[l]);
    }
    p = mk(SK, Hd, S, C);
    p.epi = EPI_COUPLING; p.inv_w = f.inv; p.an_bias = f.an_bias; p.an_scale = f.an_scale; p.half = C / 2;
    if (um && f.end.wu) {
      set_in(p, K_p);
      set_out(p, S_p, 1.f);
    }
    run(p, f.end);
    if (c->taps_d)
      B2_CUDA(cudaMemcpyAsync(c->taps_d + (size_t)blk * B * C * ldq, S, sizeof(float) * B * C * ldq,
                              cudaMemcpyDeviceToDevice, s));
  }
}

// copy a device [batch][rows][ld_d] result into a caller buffer [batch][rows][ld_o] (host or device)
static void copy_out_2d(void* dst, int ld_o, const void* src, int ld_d, int rows_total, int width, size_t elt,
                        cudaStream_t s) {
  if (!dst || rows_total <= 0 || width <= 0) return;
  B2_CUDA(cudaMemcpy2DAsync(dst, (size_t)ld_o * elt, src, (size_t)ld_d * elt, (size_t)width * elt, rows_total,
                            cudaMemcpyDefault, s));
}

static void glow_decode_impl(const b200tts_glow* m, b200tts_call* c, float noise_scale, const float* eps, int eps_ld,
                             uint64_t seed, float* mel_out, int mel_ld) {
  B2_CUDA(cudaSetDevice(c->device));
  const int batch = c->batch, mel = m->mel;
  B2_CHECK(mel_out != nullptr, "mel_out is null");
  B2_CHECK(mel_ld >= c->ty_max, "mel_ld smaller than the longest item");
  Timer tm(c);
  const bool out_dev = is_device_ptr(mel_out);
  const size_t extra = out_dev ? 0 : Arena::pad((size_t)batch * mel * std::max(mel_ld, 1), 4);
  glow_flow_impl(m, c, noise_scale, eps, eps_ld, seed, extra);
  cudaStream_t s = c->stream;
  float* dst = out_dev ? mel_out : c->dec.take<float>((size_t)batch * mel * std::max(mel_ld, 1));
  launch_unsqueeze(c->state_d, (long long)mel * m->n_sqz * c->tq_ld, c->tq_ld, c->ylen_d, mel, m->n_sqz, dst,
                   (long long)mel * mel_ld, mel_ld, nullptr, 0, 0, nullptr, batch, mel_ld, s);
  c->launches++;
  if (!out_dev && mel_ld > 0)
    B2_CUDA(cudaMemcpyAsync(mel_out, dst, sizeof(float) * batch * mel * mel_ld, cudaMemcpyDeviceToHost, s));
  tm.stop_and_sync();
}

// ------------------------------------------------------------------------------------------------
// HiFi-GAN
// ------------------------------------------------------------------------------------------------
static void hifigan_bind(b200tts_hifigan* m) {
  const Weights& W = m->W;
  const int32_t* c = W.cfg;
  B2_CHECK(W.n_cfg >= 40, "hifigan blob config too short");
  m->resblock = c[0]; m->n_ups = c[1]; m->C0 = c[2]; m->n_kernels = c[3]; m->mel = c[4];
  B2_CHECK(m->n_ups <= 8 && m->n_kernels == 3, "unsupported HiFi-GAN shape");
  m->hop = 1;
  for (int i = 0; i < m->n_ups; ++i) {
    m->up_rate[i] = c[8 + i];
    m->up_k[i] = c[16 + i];
    B2_CHECK(m->up_k[i] == 2 * m->up_rate[i], "only k = 2*stride transposed convs are supported");
    m->hop *= m->up_rate[i];
  }
  m->conv_pre = conv_w(W, "conv_pre", m->mel, m->C0, 7);
  int ch = m->C0;
  m->rb.resize(m->n_ups * m->n_kernels);
  for (int i = 0; i < m->n_ups; ++i) {
    const int cin = m->C0 >> i, cout = m->C0 >> (i + 1);
    m->ups[i] = conv_w(W, "ups" + std::to_string(i), cin, cout, 2, m->up_rate[i]);
    {  // tensor-core form: 3-tap conv with cout*u pixel-shuffled channels
      ConvW& u3 = m->ups_ps[i];
      u3.cin = cin; u3.cout = cout * m->up_rate[i]; u3.k = 3;
      const std::string nm = "ups" + std::to_string(i) + ".ps";
      if (W.has(nm + ".u")) {
        u3.b = W.get(nm + ".b", u3.cout);
        bind_umma(W, nm, u3);
      }
    }
    ch = cout;
    B2_CHECK(ch % 8 == 0, "channel counts must be multiples of 8");
    for (int j = 0; j < m->n_kernels; ++j) {
      HifiRB& r = m->rb[i * m->n_kernels + j];
      r.k = c[24 + j];
      r.ndil = m->resblock == 1 ? 3 : 2;
      std::string s = "rb" + std::to_string(i * m->n_kernels + j);
      for (int d = 0; d < r.ndil; ++d) {
        r.dil[d] = c[28 + j * 3 + d];
        r.c1[d] = conv_w(W, s + ".c1_" + std::to_string(d), ch, ch, r.k);
        if (m->resblock == 1) r.c2[d] = conv_w(W, s + ".c2_" + std::to_string(d), ch, ch, r.k);
      }
    }
  }
  m->conv_post.cin = ch; m->conv_post.cout = 1; m->conv_post.k = 7;
  m->conv_post.w = W.get("conv_post.w", (size_t)ch * 7);
  m->conv_post.b = W.get("conv_post.b", 1);
}

// mel_d: device [batch][mel][mel_ld]; frames_d device [batch]; results -> pcm buffers (host or device)
static void hifigan_run(const b200tts_hifigan* m, b200tts_call* c, const float* mel_d, int mel_ld, const int* frames_d,
                        const int* frames_h, int batch, float* pcm_f32, int16_t* pcm_i16, int pcm_ld) {
  cudaStream_t s = c->stream;
  const long long B = batch;
  int Tmax = 0;
  for (int b = 0; b < batch; ++b) Tmax = std::max(Tmax, frames_h[b]);
  const int hop = m->hop;
  B2_CHECK(pcm_ld >= Tmax * hop, "pcm_ld smaller than the longest item");
  const int ld0 = round_up(std::max(Tmax, 1), 32);
  // ---- workspace: conv_pre output + two alternating stage buffer sets of 7 activations each
  size_t stage_max = 0;
  {
    int L = ld0;
    for (int i = 0; i < m->n_ups; ++i) {
      L *= m->up_rate[i];
      stage_max = std::max(stage_max, (size_t)(m->C0 >> (i + 1)) * L);
    }
  }
  const int Lout_ld = ld0 * hop;
  size_t need = Arena::pad((size_t)B * m->C0 * ld0, 4) + 14 * Arena::pad(B * stage_max, 4) +
                Arena::pad((size_t)B * Lout_ld, 4) + Arena::pad(B, 4) + Arena::pad((size_t)B * Lout_ld, 2);
  // denoiser: STFT frames over the longest waveform, windowed inverse frames in scratch
  const bool dn = c->denoise > 0.f && m->dn_bias != nullptr;
  const int dn_nt = dn && Tmax * hop > DN_FFT ? (Tmax * hop - DN_FFT + DN_HOP - 1) / DN_HOP : 0;
  if (dn) {
    B2_CHECK(hop % DN_HOP == 0, "denoiser expects a vocoder hop that is a multiple of 256");
    need += Arena::pad((size_t)B * std::max(dn_nt, 1) * DN_FFT, 4);
  }
  // bf16 operand images: conv_pre output + 7 per stage set (two alternating sets)
  const bool um = use_umma() && m->conv_pre.wu != nullptr && m->C0 % 32 == 0;
  size_t plane_max = 0;
  if (um) {
    int Lin = Tmax;
    for (int i = 0; i < m->n_ups; ++i) {
      plane_max = std::max(plane_max, plane_item_bytes(m->C0 >> (i + 1), plane_rows(Lin, m->up_rate[i]), 2));
      Lin *= m->up_rate[i];
    }
    need += Arena::pad(B * plane_item_bytes(m->C0, plane_rows(Tmax, 1), 2), 1) + 14 * Arena::pad(B * plane_max, 1);
  }
  c->voc.reserve(need + 4096);
  float* P = c->voc.take<float>((size_t)B * m->C0 * ld0);
  float* setb[2][7];
  for (int a = 0; a < 2; ++a)
    for (int j = 0; j < 7; ++j) setb[a][j] = c->voc.take<float>(B * stage_max);
  float* pcm_d = c->voc.take<float>((size_t)B * Lout_ld);
  Planes P_p;
  unsigned char* pset[2][7] = {};
  if (um) {
    P_p = take_planes(c->voc, batch, m->C0, plane_rows(Tmax, 1), 2);
    for (int a = 0; a < 2; ++a)
      for (int j = 0; j < 7; ++j) pset[a][j] = c->voc.take<unsigned char>(B * plane_max);
  }
  unsigned int* peak_d = c->voc.take<unsigned int>(B);
  short* i16_d = c->voc.take<short>((size_t)B * Lout_ld);
  float* dn_fr = dn ? c->voc.take<float>((size_t)B * std::max(dn_nt, 1) * DN_FFT) : nullptr;

  if (Tmax > 0) {
    B2_CUDA(cudaMemsetAsync(peak_d, 0, sizeof(unsigned int) * B, s));
    // conv_pre                                                   hifi_gan/models.py:187
    {
      ConvParams p;
      p.x = mel_d; p.x_bs = (long long)m->mel * mel_ld; p.x_ld = mel_ld;
      p.y = P; p.y_bs = (long long)m->C0 * ld0; p.y_ld = ld0;
      p.lens = frames_d; p.len_mul = 1; p.Lin = Tmax; p.pad_left = 3;
      if (um) set_out(p, P_p, 0.1f);  // ups[0] consumes leaky_relu(conv_pre(x), 0.1)
      conv(c, p, m->conv_pre, batch, Tmax, s);
    }
    const float *in1 = P, *in2 = nullptr, *in3 = nullptr;
    float in_div = 1.f;
    int Lin_ld = ld0, len_mul = 1, cin = m->C0;
    Planes prev_U_p;
    bool prev_ump = false;
    for (int i = 0; i < m->n_ups; ++i) {
      const int u = m->up_rate[i], cout = m->C0 >> (i + 1);
      const int Lld = Lin_ld * u;
      float** bs_ = setb[i & 1];
      float* U = bs_[0];
      // operand images of this stage: U, R[3], T[3] (leaky_relu(.,0.1) already applied: every consumer wants it)
      Planes U_p, R_p[3], T_p[3];
      const int Lin_max = Tmax * len_mul;
      // images hold whole 32-channel chunks: narrower stages (vctk_medium tail) stay on the fp32 gather path
      const bool ump = um && cout % 32 == 0;
      if (ump) {
        auto mkp = [&](unsigned char* ptr) {
          Planes pl;
          pl.kc = (cout + 31) / 32 * 4;
          pl.rows = plane_rows(Lin_max, u);
          pl.nsplit = 2;
          pl.bs = (long long)plane_item_bytes(cout, pl.rows, 2);
          pl.ptr = ptr;
          return pl;
        };
        U_p = mkp(pset[i & 1][0]);
        for (int j = 0; j < 3; ++j) {
          R_p[j] = mkp(pset[i & 1][1 + j]);
          T_p[j] = mkp(pset[i & 1][4 + j]);
        }
      }
      // x = leaky_relu(x, 0.1); x = ups[i](x)                    hifi_gan/models.py:189-190
      {
        ConvParams p;
        p.x = in1; p.x2 = in2; p.x3 = in3; p.in_div = in_div;
        p.x_bs = (long long)cin * Lin_ld; p.x_ld = Lin_ld;
        p.pre = PRE_LRELU; p.pre_slope = 0.1f;
        p.y = U; p.y_bs = (long long)cout * Lld; p.y_ld = Lld;
        p.lens = frames_d; p.len_mul = len_mul; p.Lin = Tmax * len_mul;
        if (m->ups_ps[i].wu && use_umma()) {
          p.pix = u; p.pad_left = 1;  // one 3-tap conv, cout*u pixel-shuffled channels
          if (um && i == 0) {
            set_in(p, P_p);
          } else if (um && prev_ump && in2 && in3) {
            // later stages read the MRF mean of three tensors: build its operand image once (into the previous stage's
            // dead U image) instead of letting each of the upsampler's n-tile CTAs gather and split it again
            launch_mean3_image(in1, in2, in3, (long long)cin * Lin_ld, Lin_ld, cin, in_div, 0.1f, frames_d, len_mul, batch,
                               prev_U_p.ptr, prev_U_p.bs, prev_U_p.kc, prev_U_p.rows, s);
            c->launches++;
            set_in(p, prev_U_p);
          }
          if (ump) set_out(p, U_p, 0.1f);
          conv(c, p, m->ups_ps[i], batch, Tmax * len_mul, s);
        } else {
          p.n_phase = u; p.tpad = (m->up_k[i] - u) / 2;
          conv(c, p, m->ups[i], batch, Tmax * len_mul, s);
        }
      }
      len_mul *= u;
      const int L = Tmax * len_mul;
      // three MRF resblocks                                         hifi_gan/models.py:191-197
      // tensor-core path: the three branches (k = 3/7/11) share one launch per conv position (blockIdx.z = branch);
      // fp32 SIMT cross-check path: one launch per conv, branch after branch
      const int nk = m->n_kernels;
      auto fill_c1 = [&](int j, int d, ConvParams& p, bool& planes) {
        const HifiRB& r = m->rb[i * nk + j];
        float* R = bs_[1 + j];
        float* T = bs_[4 + j];
        const float* xin = d == 0 ? U : R;
        p = ConvParams();
        p.x = xin; p.x_bs = (long long)cout * Lld; p.x_ld = Lld;
        p.pre = PRE_LRELU; p.pre_slope = 0.1f;
        p.lens = frames_d; p.len_mul = len_mul; p.Lin = L;
        p.dil = r.dil[d]; p.pad_left = (r.k * r.dil[d] - r.dil[d]) / 2;
        p.y = T; p.y_bs = (long long)cout * Lld; p.y_ld = Lld;
        planes = ump && r.c1[d].wu && r.c2[d].wu;
        if (planes) {
          set_in(p, d == 0 ? U_p : R_p[j]);
          set_out(p, T_p[j], 0.1f);
          p.y = nullptr;  // T is only ever read as an MMA operand
        }
      };
      auto fill_c2 = [&](int j, int d, ConvParams& q, bool planes) {
        const HifiRB& r = m->rb[i * nk + j];
        float* R = bs_[1 + j];
        float* T = bs_[4 + j];
        const float* xin = d == 0 ? U : R;
        q = ConvParams();
        q.x = T; q.x_bs = (long long)cout * Lld; q.x_ld = Lld;
        q.pre = PRE_LRELU; q.pre_slope = 0.1f;
        q.lens = frames_d; q.len_mul = len_mul; q.Lin = L;
        q.dil = 1; q.pad_left = (r.k - 1) / 2;
        q.y = R; q.y_bs = (long long)cout * Lld; q.y_ld = Lld;
        q.res = xin; q.res_bs = (long long)cout * Lld; q.res_ld = Lld;
        if (planes) {
          set_in(q, T_p[j]);
          if (d + 1 < r.ndil) set_out(q, R_p[j], 0.1f);
          {
            // the block input already exists as a hi/lo image (c1's operand): read the residual from it and keep the
            // fp32 copy of x only where a non-MMA consumer needs it (last dilation: upsampler gather / conv_post)
            const Planes& rpl = d == 0 ? U_p : R_p[j];
            q.rp = rpl.ptr; q.rp_bs = rpl.bs; q.rp_kc = rpl.kc; q.rp_rows = rpl.rows; q.rp_inv_slope = 10.f;
            q.res = nullptr;
            if (d + 1 < r.ndil) q.y = nullptr;
          }
        }
      };
      auto bind = [&](ConvParams& p, const ConvW& w) {
        p.w = w.w; p.bias = w.b; p.Cin = w.cin; p.Cout = w.cout; p.K = w.k;
        p.wu = w.wu; p.um_nt = w.nt; p.um_nsplit = w.nsplit;
      };
      const bool umma_rb = m->resblock == 1 && use_umma() && nk == 3 && m->rb[i * nk].c1[0].wu != nullptr &&
                           !conv_trace();
      bool staged = false;
      if (umma_rb) {
        // all ResBlock convs of the stage in one persistent launch with tile-level dataflow (stage_umma.cuh)
        static const bool no_stage = dbg_flag("no_stage");
        const int ndil = m->rb[i * nk].ndil;
        if (!no_stage && 2 * ndil * 3 <= SG_MAX_CONVS && i < 8) {
          ConvParams tab[SG_MAX_CONVS];
          bool all_planes = true;
          for (int d = 0; d < ndil; ++d)
            for (int j = 0; j < 3; ++j) {
              bool pl = false;
              fill_c1(j, d, tab[(2 * d) * 3 + j], pl);
              bind(tab[(2 * d) * 3 + j], m->rb[i * nk + j].c1[d]);
              all_planes = all_planes && pl && m->rb[i * nk + j].ndil == ndil;
              fill_c2(j, d, tab[(2 * d + 1) * 3 + j], pl);
              bind(tab[(2 * d + 1) * 3 + j], m->rb[i * nk + j].c2[d]);
            }
          // 32-channel stage: whole ResBlock steps per item (T in shared memory); x ping-pongs U -> R -> T(free) -> fp32 R
          ConvParams f1[9], f2[9];
          static const bool no_pair_stage = dbg_flag("no_stage_pair");
          bool pair = !no_pair_stage && all_planes && ndil <= 3 && cout <= 32;
          if (pair) {
            for (int d = 0; d < ndil; ++d)
              for (int j = 0; j < 3; ++j) {
                const Planes& in = d == 0 ? U_p : ((d & 1) ? R_p[j] : T_p[j]);
                const Planes& out = (d & 1) ? T_p[j] : R_p[j];
                ConvParams& a = f1[d * 3 + j];
                ConvParams& q = f2[d * 3 + j];
                a = tab[(2 * d) * 3 + j];
                q = tab[(2 * d + 1) * 3 + j];
                set_in(a, in);
                a.yp = nullptr;
                a.yp_slope = 0.1f;  // T = leaky_relu(c1(.), 0.1), kept on the SM
                q.xp = nullptr;
                q.rp = in.ptr; q.rp_bs = in.bs; q.rp_kc = in.kc; q.rp_rows = in.rows; q.rp_inv_slope = 10.f;
                q.yp = nullptr;
                if (d + 1 < ndil) set_out(q, out, 0.1f);
              }
            pair = stage_pair_supported(f1, f2, ndil, 3);
          }
          if (pair || (all_planes && stage_umma_supported(tab, 2 * ndil, 3))) {
            if (!c->stage_tab_h[i]) {
              B2_CUDA(cudaMallocHost(&c->stage_tab_h[i], stage_umma_table_bytes()));
              B2_CUDA(cudaMalloc(&c->stage_tab_d[i], stage_umma_table_bytes()));
            }
            const size_t need_flags = (size_t)2 * ndil * 3 * batch * ((L + 127) / 128 + 1);
            if (need_flags > c->stage_flags_cap[i]) {
              if (c->stage_flags[i]) B2_CUDA(cudaFree(c->stage_flags[i]));
              c->stage_flags[i] = nullptr;
              B2_CUDA(cudaMalloc((void**)&c->stage_flags[i], sizeof(unsigned int) * need_flags * 2));
              c->stage_flags_cap[i] = need_flags * 2;
            }
            if (pair)
              c->launches += launch_stage_pair(f1, f2, ndil, 3, batch, frames_h, c->stage_tab_h[i], c->stage_tab_d[i], c->stage_flags[i],
                                               c->stage_flags_cap[i], s);
            else
              c->launches += launch_stage_umma(tab, 2 * ndil, 3, batch, frames_h, c->stage_tab_h[i], c->stage_tab_d[i], c->stage_flags[i],
                                               c->stage_flags_cap[i], s, dbg_flag("stage_trace"));
            staged = true;
          }
        }
      }
      if (staged) {
      } else if (umma_rb) {
        const int ndil = m->rb[i * nk].ndil;
        static const bool no_pair = dbg_flag("no_rbpair");
        for (int d = 0; d < ndil; ++d) {
          ConvParams ps[3], qs[3];
          bool pl[3];
          for (int j = 0; j < 3; ++j) {
            fill_c1(j, d, ps[j], pl[j]);
            bind(ps[j], m->rb[i * nk + j].c1[d]);
            fill_c2(j, d, qs[j], pl[j]);
            bind(qs[j], m->rb[i * nk + j].c2[d]);
          }
          // narrow stages: c1 -> c2 fused, T stays in shared memory (rbpair_umma.cuh).  Its tiles overlap, so a step must not
          // rewrite the image it reads: x ping-pongs U -> R -> T(free in this mode) -> fp32 R.
          ConvParams f1[3], f2[3];
          bool fused = !no_pair && pl[0] && pl[1] && pl[2] && 3 * batch <= 148;
          if (fused) {
            for (int j = 0; j < 3; ++j) {
              const Planes& in = d == 0 ? U_p : ((d & 1) ? R_p[j] : T_p[j]);
              const Planes& out = (d & 1) ? T_p[j] : R_p[j];
              f1[j] = ps[j];
              f2[j] = qs[j];
              set_in(f1[j], in);
              f1[j].yp = nullptr;
              f1[j].yp_slope = 0.1f;  // T = leaky_relu(c1(.), 0.1), kept on the SM
              f2[j].xp = nullptr;
              f2[j].rp = in.ptr; f2[j].rp_bs = in.bs; f2[j].rp_kc = in.kc; f2[j].rp_rows = in.rows; f2[j].rp_inv_slope = 10.f;
              if (d + 1 < ndil) set_out(f2[j], out, 0.1f);
            }
            fused = rbpair_umma_supported(f1, f2, 3);
          }
          if (fused) {
            c->launches += launch_rbpair_umma(f1, f2, 3, batch, L, s);
          } else {
            c->launches += launch_conv_umma_batch(ps, 3, batch, L, s);
            c->launches += launch_conv_umma_batch(qs, 3, batch, L, s);
          }
        }
      } else {
        for (int j = 0; j < nk; ++j) {
          const HifiRB& r = m->rb[i * nk + j];
          float* R = bs_[1 + j];
          float* T = bs_[4 + j];
          for (int d = 0; d < r.ndil; ++d) {
            if (m->resblock == 1) {
              ConvParams p, q;
              bool pl;
              fill_c1(j, d, p, pl);
              conv(c, p, r.c1[d], batch, L, s);
              fill_c2(j, d, q, pl);
              conv(c, q, r.c2[d], batch, L, s);
            } else {  // ResBlock2                                   hifi_gan/models.py:137-142
              ConvParams p;
              float* dst = (d & 1) ? R : T;  // ping-pong: conv input and output must differ
              const float* src = d == 0 ? U : ((d & 1) ? T : R);
              p.x = src; p.x_bs = (long long)cout * Lld; p.x_ld = Lld;
              p.pre = PRE_LRELU; p.pre_slope = 0.1f;
              p.lens = frames_d; p.len_mul = len_mul; p.Lin = L;
              p.dil = r.dil[d]; p.pad_left = (r.k * r.dil[d] - r.dil[d]) / 2;
              p.y = dst; p.y_bs = (long long)cout * Lld; p.y_ld = Lld;
              p.res = src; p.res_bs = (long long)cout * Lld; p.res_ld = Lld;
              conv(c, p, r.c1[d], batch, L, s);
            }
          }
        }
      }
      // the resblock result buffer: ResBlock1 -> R; ResBlock2 with ndil=2 -> last dst = R (d=1)
      in1 = bs_[1]; in2 = bs_[2]; in3 = bs_[3];
      prev_U_p = U_p;
      prev_ump = ump && m->n_kernels == 3;
      in_div = (float)m->n_kernels;
      Lin_ld = Lld;
      cin = cout;
    }
    // leaky_relu(0.01) -> conv_post -> tanh                       hifi_gan/models.py:198-200
    launch_conv_post(in1, in2, in3, (long long)cin * Lin_ld, Lin_ld, in_div, 0.01f, m->conv_post.w, m->conv_post.b, cin,
                     m->conv_post.k, frames_d, len_mul, batch, Tmax * len_mul, pcm_d, Lout_ld, peak_d, s);
    c->launches++;
    if (dn) {
      // magnitude-domain bias subtraction frame by frame, then overlap-add in place        larynx/hifi_gan.py:171-179
      const long long fr_bs = (long long)std::max(dn_nt, 1) * DN_FFT;
      B2_CUDA(cudaMemsetAsync(peak_d, 0, sizeof(unsigned int) * B, s));
      if (dn_nt > 0) {
        launch_stft_denoise(pcm_d, Lout_ld, frames_d, len_mul, batch, dn_nt, m->dn_hann, m->dn_tw, m->dn_bias, c->denoise,
                            dn_fr, fr_bs, nullptr, s);
        c->launches++;
      }
      launch_overlap_add(dn_fr, fr_bs, frames_d, len_mul, batch, std::min(dn_nt * DN_HOP + DN_FFT, Lout_ld), pcm_d, Lout_ld,
                         peak_d, s);
      c->launches++;
    }
    if (pcm_i16) {
      launch_pcm_to_i16(pcm_d, Lout_ld, peak_d, frames_d, len_mul, batch, i16_d, Lout_ld, s);
      c->launches++;
    }
    if (pcm_f32) {
      launch_zero_tail(pcm_d, Lout_ld, Lout_ld, 1, frames_d, len_mul, batch, Lout_ld, s);
      c->launches++;
    }
  } else {
    B2_CUDA(cudaMemsetAsync(pcm_d, 0, sizeof(float) * B * Lout_ld, s));
    B2_CUDA(cudaMemsetAsync(i16_d, 0, sizeof(short) * B * Lout_ld, s));
  }
  const int w = std::min(pcm_ld, Lout_ld);
  copy_out_2d(pcm_f32, pcm_ld, pcm_d, Lout_ld, batch, w, sizeof(float), s);
  copy_out_2d(pcm_i16, pcm_ld, i16_d, Lout_ld, batch, w, sizeof(short), s);
}

// Denoiser tables and bias spectrum (larynx/hifi_gan.py:181-203): the bias is the magnitude of the first STFT frame of
// the vocoder's response to an all-zero mel of 88 frames.  Computed once at model creation so that the model object stays
// immutable afterwards (the reference initialises it lazily on first use).
static void hifigan_init_denoiser(b200tts_hifigan* m) {
  if (m->hop % DN_HOP != 0) return;  // no reference voice has such a vocoder; the denoiser stays unavailable
  const int dev = m->W.device;
  B2_CUDA(cudaSetDevice(dev));
  std::vector<float> hann(DN_FFT), tw(2 * DN_FFT);
  const double pi = 3.14159265358979323846;
  for (int n = 0; n < DN_FFT; ++n) {
    hann[n] = (float)(0.5 - 0.5 * std::cos(2.0 * pi * n / (DN_FFT - 1)));  // numpy.hanning: symmetric window
    tw[2 * n] = (float)std::cos(2.0 * pi * n / DN_FFT);
    tw[2 * n + 1] = (float)(-std::sin(2.0 * pi * n / DN_FFT));
  }
  B2_CUDA(cudaMalloc(&m->dn_hann, sizeof(float) * DN_FFT));
  B2_CUDA(cudaMalloc(&m->dn_tw, sizeof(float) * 2 * DN_FFT));
  B2_CUDA(cudaMemcpy(m->dn_hann, hann.data(), sizeof(float) * DN_FFT, cudaMemcpyHostToDevice));
  B2_CUDA(cudaMemcpy(m->dn_tw, tw.data(), sizeof(float) * 2 * DN_FFT, cudaMemcpyHostToDevice));
  const int T = 88, N = T * m->hop;
  float *mel_d = nullptr, *pcm_d = nullptr, *bias_d = nullptr;
  int* frames_d = nullptr;
  b200tts_call* c = nullptr;
  auto cleanup = [&]() {
    if (c) b200tts_call_destroy(c);
    cudaFree(mel_d);
    cudaFree(pcm_d);
    cudaFree(frames_d);
  };
  try {
    if (b200tts_call_create(dev, &c) != 0) throw Error(g_err);
    B2_CUDA(cudaMalloc(&mel_d, sizeof(float) * m->mel * T));
    B2_CUDA(cudaMalloc(&pcm_d, sizeof(float) * N));
    B2_CUDA(cudaMalloc(&frames_d, sizeof(int)));
    B2_CUDA(cudaMalloc(&bias_d, sizeof(float) * DN_BINS));
    B2_CUDA(cudaMemset(mel_d, 0, sizeof(float) * m->mel * T));
    B2_CUDA(cudaMemcpy(frames_d, &T, sizeof(int), cudaMemcpyHostToDevice));
    hifigan_run(m, c, mel_d, T, frames_d, &T, 1, pcm_d, nullptr, N);
    launch_stft_denoise(pcm_d, N, frames_d, m->hop, 1, 1, m->dn_hann, m->dn_tw, nullptr, 0.f, nullptr, 0, bias_d, c->stream);
    B2_CUDA(cudaStreamSynchronize(c->stream));
  } catch (...) {
    cleanup();
    cudaFree(bias_d);
    throw;
  }
  cleanup();
  m->dn_bias = bias_d;
}

// HiFiGanVocoder.denoise (larynx/hifi_gan.py:171-179) as a call of its own on float waveforms
static void denoise_impl(const b200tts_hifigan* m, b200tts_call* c, float strength, const float* bias, const float* pcm_in,
                         const int32_t* n_samples, int batch, int in_ld, float* out_f32, int16_t* out_i16, int out_ld) {
  B2_CHECK(batch > 0, "empty batch");
  B2_CHECK(m->W.device == c->device, "model and call live on different devices");
  B2_CHECK(m->dn_hann != nullptr, "this vocoder has no denoiser (hop is not a multiple of 256)");
  B2_CHECK(strength >= 0.f, "denoiser strength must be >= 0");
  B2_CUDA(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  std::vector<int> nh(batch), oh(batch);
  if (is_device_ptr(n_samples)) B2_CUDA(cudaMemcpy(nh.data(), n_samples, sizeof(int) * batch, cudaMemcpyDeviceToHost));
  else memcpy(nh.data(), n_samples, sizeof(int) * batch);
  int nt_max = 0, n_max = 0;
  for (int b = 0; b < batch; ++b) {
    B2_CHECK(nh[b] >= 0 && nh[b] <= in_ld, "n_samples[b] out of range");
    const int nt = nh[b] > DN_FFT ? (nh[b] - DN_FFT + DN_HOP - 1) / DN_HOP : 0;
    oh[b] = nt * DN_HOP + DN_FFT;
    nt_max = std::max(nt_max, nt);
    n_max = std::max(n_max, nh[b]);
  }
  const int olen_max = nt_max * DN_HOP + DN_FFT;
  B2_CHECK(out_ld >= olen_max, "out_ld smaller than the longest denoised item (n_t*256 + 1024)");
  const long long B = batch;
  const bool in_dev = is_device_ptr(pcm_in), bias_dev = bias == nullptr || is_device_ptr(bias);
  const long long fr_bs = (long long)std::max(nt_max, 1) * DN_FFT;
  c->stg.reserve(2 * Arena::pad(B, 4) + (in_dev ? 0 : Arena::pad((size_t)B * std::max(in_ld, 1), 4)) + Arena::pad(DN_BINS, 4) +
                 Arena::pad((size_t)B * olen_max, 4) + Arena::pad((size_t)B * fr_bs, 4) + Arena::pad(B, 4) +
                 Arena::pad((size_t)B * olen_max, 2) + 4096);
  int* n_d = c->stg.take<int>(B);
  int* o_d = c->stg.take<int>(B);
  Timer tm(c);
  B2_CUDA(cudaMemcpyAsync(n_d, nh.data(), sizeof(int) * B, cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaMemcpyAsync(o_d, oh.data(), sizeof(int) * B, cudaMemcpyHostToDevice, s));
  const float* in_d = pcm_in;
  if (!in_dev) {
    float* st = c->stg.take<float>((size_t)B * std::max(in_ld, 1));
    if (in_ld > 0) B2_CUDA(cudaMemcpyAsync(st, pcm_in, sizeof(float) * B * in_ld, cudaMemcpyHostToDevice, s));
    in_d = st;
  }
  const float* bias_d = bias ? bias : m->dn_bias;
  if (!bias_dev) {
    float* st = c->stg.take<float>(DN_BINS);
    B2_CUDA(cudaMemcpyAsync(st, bias, sizeof(float) * DN_BINS, cudaMemcpyHostToDevice, s));
    bias_d = st;
  }
  B2_CHECK(bias_d != nullptr, "no bias spectrum");
  float* work = c->stg.take<float>((size_t)B * olen_max);
  float* fr = c->stg.take<float>((size_t)B * fr_bs);
  unsigned int* peak_d = c->stg.take<unsigned int>(B);
  short* i16_d = c->stg.take<short>((size_t)B * olen_max);
  B2_CUDA(cudaMemsetAsync(peak_d, 0, sizeof(unsigned int) * B, s));
  B2_CUDA(cudaMemsetAsync(work, 0, sizeof(float) * B * olen_max, s));
  if (nt_max > 0) {
    launch_stft_denoise(in_d, in_ld, n_d, 1, batch, nt_max, m->dn_hann, m->dn_tw, bias_d, strength, fr, fr_bs, nullptr, s);
    c->launches++;
  }
  launch_overlap_add(fr, fr_bs, n_d, 1, batch, olen_max, work, olen_max, peak_d, s);
  c->launches++;
  if (out_i16) {
    launch_pcm_to_i16(work, olen_max, peak_d, o_d, 1, batch, i16_d, olen_max, s);
    c->launches++;
  }
  copy_out_2d(out_f32, out_ld, work, olen_max, batch, olen_max, sizeof(float), s);
  copy_out_2d(out_i16, out_ld, i16_d, olen_max, batch, olen_max, sizeof(short), s);
  tm.stop_and_sync();
  // the host vectors nh / oh outlive the async copies: stop_and_sync() waited for the stream
}

static void hifigan_infer_impl(const b200tts_hifigan* m, b200tts_call* c, const float* mel, const int32_t* frames,
                               int batch, int mel_ld, float* pcm_f32, int16_t* pcm_i16, int pcm_ld) {
  B2_CHECK(batch > 0, "empty batch");
  B2_CHECK(m->W.device == c->device, "model and call live on different devices");
  B2_CUDA(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  std::vector<int> fh(batch);
  if (is_device_ptr(frames)) B2_CUDA(cudaMemcpy(fh.data(), frames, sizeof(int) * batch, cudaMemcpyDeviceToHost));
  else memcpy(fh.data(), frames, sizeof(int) * batch);
  int Tmax = 0;
  for (int b = 0; b < batch; ++b) {
    B2_CHECK(fh[b] >= 0 && fh[b] <= mel_ld, "frames[b] out of range");
    Tmax = std::max(Tmax, fh[b]);
  }
  // stage inputs in the decode arena (the vocoder arena is reserved inside hifigan_run)
  const bool mel_dev = is_device_ptr(mel);
  c->stg.reserve(Arena::pad(batch, 4) + (mel_dev ? 0 : Arena::pad((size_t)batch * m->mel * std::max(mel_ld, 1), 4)) + 4096);
  int* frames_d = c->stg.take<int>(batch);
  Timer tm(c);
  B2_CUDA(cudaMemcpyAsync(frames_d, fh.data(), sizeof(int) * batch, cudaMemcpyHostToDevice, s));
  const float* mel_d = mel;
  if (!mel_dev) {
    float* md = c->stg.take<float>((size_t)batch * m->mel * std::max(mel_ld, 1));
    if (mel_ld > 0)
      B2_CUDA(cudaMemcpyAsync(md, mel, sizeof(float) * batch * m->mel * mel_ld, cudaMemcpyHostToDevice, s));
    mel_d = md;
  }
  hifigan_run(m, c, mel_d, mel_ld, frames_d, fh.data(), batch, pcm_f32, pcm_i16, pcm_ld);
  tm.stop_and_sync();
}

static void tts_finish_impl(const b200tts_glow* g, const b200tts_hifigan* v, b200tts_call* c, const b200tts_glue* glue,
                            float noise_scale, const float* eps, int eps_ld, uint64_t seed, float* mel_out, int mel_ld,
                            float* pcm_f32, int16_t* pcm_i16, int pcm_ld) {
  B2_CHECK(glue != nullptr, "glue settings are required");
  B2_CHECK(g->W.device == c->device && v->W.device == c->device, "models and call live on different devices");
  B2_CHECK(g->mel == v->mel, "mel channel mismatch between the two models");
  B2_CUDA(cudaSetDevice(c->device));
  const int batch = c->batch, mel = g->mel;
  const int Ty = c->ty_max;
  const int gl_ld = round_up(std::max(Ty, 1), 32);
  if (mel_out) B2_CHECK(mel_ld >= Ty, "mel_ld smaller than the longest item");
  Timer tm(c);
  const bool mel_host = mel_out && !is_device_ptr(mel_out);
  size_t extra = Arena::pad((size_t)batch * mel * gl_ld, 4);
  if (mel_host) extra += Arena::pad((size_t)batch * mel * std::max(mel_ld, 1), 4);
  glow_flow_impl(g, c, noise_scale, eps, eps_ld, seed, extra);
  cudaStream_t s = c->stream;
  float* glued = c->dec.take<float>((size_t)batch * mel * gl_ld);
  float* mdst = nullptr;
  if (mel_out) mdst = mel_host ? c->dec.take<float>((size_t)batch * mel * std::max(mel_ld, 1)) : mel_out;
  const int width = std::max(mel_out ? mel_ld : 0, Ty);
  launch_unsqueeze(c->state_d, (long long)mel * g->n_sqz * c->tq_ld, c->tq_ld, c->ylen_d, mel, g->n_sqz, mdst,
                   (long long)mel * mel_ld, mel_ld, glued, (long long)mel * gl_ld, gl_ld, glue, batch, width, s);
  c->launches++;
  if (mel_host && mel_ld > 0)
    B2_CUDA(cudaMemcpyAsync(mel_out, mdst, sizeof(float) * batch * mel * mel_ld, cudaMemcpyDeviceToHost, s));
  hifigan_run(v, c, glued, gl_ld, c->ylen_d, c->y_len_h.data(), batch, pcm_f32, pcm_i16, pcm_ld);
  tm.stop_and_sync();
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* b200tts_last_error(void) { return g_err.c_str(); }
int b200tts_version(void) { return B200TTS_VERSION; }

int b200tts_device_count(int* count) {
  B2_API_BEGIN
  B2_CHECK(count != nullptr, "null argument");
  B2_CUDA(cudaGetDeviceCount(count));
  B2_API_END
}

int b200tts_call_create(int device, b200tts_call** out) {
  B2_API_BEGIN
  B2_CHECK(out != nullptr, "null argument");
  B2_CUDA(cudaSetDevice(device));
  conv1d_init_attributes();
  conv_umma_init_attributes();
  rbpair_umma_init_attributes();
  stage_umma_init_attributes();
  stage_set_striped(!dbg_flag("no_stage_stripe"));
  attention_init_attributes();
  attn_umma_init_attributes();
  flow_chain_init_attributes();
  enc_chain_init_attributes();
  b200tts_call* c = new b200tts_call();
  c->device = device;
  B2_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    B2_CUDA(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
    B2_CUDA(cudaEventCreateWithFlags(&c->join[i], cudaEventDisableTiming));
  }
  B2_CUDA(cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming));
  B2_CUDA(cudaEventCreate(&c->ev0));
  B2_CUDA(cudaEventCreate(&c->ev1));
  *out = c;
  B2_API_END
}

int b200tts_call_destroy(b200tts_call* c) {
  B2_API_BEGIN
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 2; ++i) {
    cudaStreamSynchronize(c->aux[i]);
    cudaStreamDestroy(c->aux[i]);
    cudaEventDestroy(c->join[i]);
  }
  cudaFreeHost(c->enc_steps_h);
  cudaFree(c->enc_steps_d);
  for (int i = 0; i < 8; ++i) {
    cudaFreeHost(c->stage_tab_h[i]);
    cudaFree(c->stage_tab_d[i]);
    cudaFree(c->stage_flags[i]);
  }
  cudaStreamDestroy(c->stream);
  cudaEventDestroy(c->fork);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  delete c;
  B2_API_END
}

int b200tts_call_last_ms(b200tts_call* c, float* ms) {
  B2_API_BEGIN
  B2_CHECK(c && ms, "null argument");
  *ms = c->last_ms;
  B2_API_END
}
int b200tts_call_launches(b200tts_call* c, int64_t* n) {
  B2_API_BEGIN
  B2_CHECK(c && n, "null argument");
  *n = c->launches;
  B2_API_END
}
int b200tts_call_stream(b200tts_call* c, void** stream) {
  B2_API_BEGIN
  B2_CHECK(c && stream, "null argument");
  *stream = (void*)c->stream;
  B2_API_END
}
int b200tts_call_set_debug(b200tts_call* c, int on) {
  B2_API_BEGIN
  B2_CHECK(c, "null argument");
  c->debug = on;
  B2_API_END
}

int b200tts_glow_create(const void* blob, size_t nbytes, int device, b200tts_glow** out) {
  B2_API_BEGIN
  B2_CHECK(blob && out, "null argument");
  std::unique_ptr<b200tts_glow> m(new b200tts_glow());
  m->W.load(blob, nbytes, device, 1);
  glow_bind(m.get());
  *out = m.release();
  B2_API_END
}
int b200tts_glow_destroy(b200tts_glow* m) {
  B2_API_BEGIN
  if (m) {
    cudaSetDevice(m->W.device);
    delete m;
  }
  B2_API_END
}
int b200tts_glow_info(const b200tts_glow* m, int32_t* num_symbols, int32_t* mel_channels, int32_t* n_sqz) {
  B2_API_BEGIN
  B2_CHECK(m, "null argument");
  if (num_symbols) *num_symbols = m->V;
  if (mel_channels) *mel_channels = m->mel;
  if (n_sqz) *n_sqz = m->n_sqz;
  B2_API_END
}

int b200tts_glow_encode(const b200tts_glow* m, b200tts_call* call, const int64_t* ids, const int32_t* x_len, int batch,
                        int max_tx, float length_scale, int32_t* y_len_out) {
  B2_API_BEGIN
  B2_CHECK(m && call && ids && x_len, "null argument");
  if (!is_device_ptr(x_len))
    for (int b = 0; b < batch; ++b) B2_CHECK(x_len[b] >= 0 && x_len[b] <= max_tx, "x_len[b] out of range");
  glow_encode_impl(m, call, ids, x_len, batch, max_tx, length_scale, y_len_out);
  B2_API_END
}

int b200tts_glow_decode(const b200tts_glow* m, b200tts_call* call, float noise_scale, const float* eps, int eps_ld,
                        uint64_t seed, float* mel_out, int mel_ld) {
  B2_API_BEGIN
  B2_CHECK(m && call, "null argument");
  glow_decode_impl(m, call, noise_scale, eps, eps_ld, seed, mel_out, mel_ld);
  B2_API_END
}

int b200tts_glow_alignment(const b200tts_glow* m, b200tts_call* c, float* logw, int32_t* w_ceil, int32_t* idx,
                           int idx_ld) {
  B2_API_BEGIN
  B2_CHECK(m && c && c->enc_model == m, "no encode state on this call");
  B2_CUDA(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  const size_t n = (size_t)c->batch * c->max_tx;
  if (logw) B2_CUDA(cudaMemcpyAsync(logw, c->logw_d, sizeof(float) * n, cudaMemcpyDefault, s));
  if (w_ceil) B2_CUDA(cudaMemcpyAsync(w_ceil, c->wceil_d, sizeof(int) * n, cudaMemcpyDefault, s));
  if (idx) {
    B2_CHECK(c->idx_d != nullptr, "alignment indices exist only after decode");
    B2_CHECK(idx_ld >= c->ty_max, "idx_ld smaller than the longest item");
    copy_out_2d(idx, idx_ld, c->idx_d, c->idx_ld, c->batch, std::min(idx_ld, c->idx_ld), sizeof(int), s);
  }
  B2_CUDA(cudaStreamSynchronize(s));
  B2_API_END
}

int b200tts_debug_noise(b200tts_call* c, uint64_t seed, int batch, int mel, int n, float* out, int ld) {
  B2_API_BEGIN
  B2_CHECK(c && out, "null argument");
  B2_CHECK(batch > 0 && mel > 0 && n >= 0 && ld >= n, "bad noise shape");
  B2_CUDA(cudaSetDevice(c->device));
  const bool dev = is_device_ptr(out);
  float* d = out;
  if (!dev) {
    c->stg.reserve(Arena::pad((size_t)batch * mel * ld, 4) + 4096);
    d = c->stg.take<float>((size_t)batch * mel * ld);
  }
  launch_noise_fill(seed, batch, mel, n, d, (long long)mel * ld, ld, c->stream);
  c->launches++;
  if (!dev) B2_CUDA(cudaMemcpyAsync(out, d, sizeof(float) * batch * mel * ld, cudaMemcpyDeviceToHost, c->stream));
  B2_CUDA(cudaStreamSynchronize(c->stream));
  B2_API_END
}

int b200tts_debug_stage_partition(const int32_t* strip_tiles, const int64_t* strip_cost, int n_strips, int nctas,
                                  int32_t* cta_of_tile, int32_t* n_ranges) {
  B2_API_BEGIN
  static_assert(sizeof(long long) == sizeof(int64_t) && sizeof(int) == sizeof(int32_t), "ABI types");
  const int n = stage_partition_debug(reinterpret_cast<const int*>(strip_tiles), reinterpret_cast<const long long*>(strip_cost), n_strips, nctas,
                                      reinterpret_cast<int*>(cta_of_tile));
  if (n_ranges) *n_ranges = n;
  B2_API_END
}

int b200tts_glow_flow_tap(const b200tts_glow* m, b200tts_call* c, int blk, float* out, int ld) {
  B2_API_BEGIN
  B2_CHECK(m && c && out, "null argument");
  B2_CHECK(c->taps_d != nullptr, "flow taps were not recorded (set debug before decode)");
  B2_CHECK(blk >= 0 && blk < m->n_blocks, "block index out of range");
  B2_CUDA(cudaSetDevice(c->device));
  const int C = m->mel * m->n_sqz;
  const int Tq = c->ty_max / m->n_sqz;
  B2_CHECK(ld >= Tq, "ld too small");
  copy_out_2d(out, ld, c->taps_d + (size_t)blk * c->batch * C * c->tq_ld, c->tq_ld, c->batch * C, Tq, sizeof(float),
              c->stream);
  B2_CUDA(cudaStreamSynchronize(c->stream));
  B2_API_END
}

int b200tts_hifigan_create(const void* blob, size_t nbytes, int device, b200tts_hifigan** out) {
  B2_API_BEGIN
  B2_CHECK(blob && out, "null argument");
  std::unique_ptr<b200tts_hifigan> m(new b200tts_hifigan());
  m->W.load(blob, nbytes, device, 2);
  hifigan_bind(m.get());
  hifigan_init_denoiser(m.get());
  *out = m.release();
  B2_API_END
}
int b200tts_hifigan_destroy(b200tts_hifigan* m) {
  B2_API_BEGIN
  if (m) {
    cudaSetDevice(m->W.device);
    delete m;
  }
  B2_API_END
}
int b200tts_hifigan_info(const b200tts_hifigan* m, int32_t* mel_channels, int32_t* hop) {
  B2_API_BEGIN
  B2_CHECK(m, "null argument");
  if (mel_channels) *mel_channels = m->mel;
  if (hop) *hop = m->hop;
  B2_API_END
}

int b200tts_hifigan_infer(const b200tts_hifigan* m, b200tts_call* call, const float* mel, const int32_t* frames,
                          int batch, int mel_ld, float* pcm_f32, int16_t* pcm_i16, int pcm_ld) {
  B2_API_BEGIN
  B2_CHECK(m && call && mel && frames, "null argument");
  hifigan_infer_impl(m, call, mel, frames, batch, mel_ld, pcm_f32, pcm_i16, pcm_ld);
  B2_API_END
}

int b200tts_call_set_denoiser(b200tts_call* call, float strength) {
  B2_API_BEGIN
  B2_CHECK(call, "null argument");
  B2_CHECK(strength == strength, "denoiser strength is NaN");
  call->denoise = strength > 0.f ? strength : 0.f;
  B2_API_END
}

int b200tts_call_set_encoder_bf16(b200tts_call* call, int on) {
  B2_API_BEGIN
  B2_CHECK(call, "null argument");
  call->enc_bf16 = on ? 1 : 0;
  B2_API_END
}

int b200tts_hifigan_denoise(const b200tts_hifigan* m, b200tts_call* call, float strength, const float* bias,
                            const float* pcm_in, const int32_t* n_samples, int batch, int in_ld, float* pcm_f32,
                            int16_t* pcm_i16, int out_ld) {
  B2_API_BEGIN
  B2_CHECK(m && call && pcm_in && n_samples, "null argument");
  denoise_impl(m, call, strength, bias, pcm_in, n_samples, batch, in_ld, pcm_f32, pcm_i16, out_ld);
  B2_API_END
}

int b200tts_hifigan_bias_spectrum(const b200tts_hifigan* m, float* out, int n) {
  B2_API_BEGIN
  B2_CHECK(m && out, "null argument");
  B2_CHECK(m->dn_bias != nullptr, "this vocoder has no denoiser (hop is not a multiple of 256)");
  B2_CHECK(n >= DN_BINS, "bias spectrum needs room for 513 values");
  B2_CUDA(cudaSetDevice(m->W.device));
  B2_CUDA(cudaMemcpy(out, m->dn_bias, sizeof(float) * DN_BINS, cudaMemcpyDefault));
  B2_API_END
}

int b200tts_tts_finish(const b200tts_glow* g, const b200tts_hifigan* v, b200tts_call* call, const b200tts_glue* glue,
                       float noise_scale, const float* eps, int eps_ld, uint64_t seed, float* mel_out, int mel_ld,
                       float* pcm_f32, int16_t* pcm_i16, int pcm_ld) {
  B2_API_BEGIN
  B2_CHECK(g && v && call, "null argument");
  tts_finish_impl(g, v, call, glue, noise_scale, eps, eps_ld, seed, mel_out, mel_ld, pcm_f32, pcm_i16, pcm_ld);
  B2_API_END
}

}  // extern "C"
