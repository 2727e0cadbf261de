// conv_umma.cu — launcher of the tcgen05 implicit-GEMM conv (conv_umma.cuh).
#include <algorithm>

#include "conv_umma.cuh"
#include "kernels.cuh"

namespace b2 {

template <int MT, int NSPLIT, int EW, int WST, int AST, int G, int NBUF>
static void set_attr_umma() {
  B2_CUDA(cudaFuncSetAttribute(conv_umma_kernel<MT, NSPLIT, EW, WST, AST, G, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
}

// kernel variants (MT, NSPLIT, EW, WST, AST, G, NBUF)
#define UMMA_BIG 1, 2, 8, 3, 2, 1, 1        /* large grids: two CTAs per SM */
#define UMMA_BIG2 2, 2, 8, 3, 2, 1, 1       /* 256-row tiles, NT <= 64 (HiFi-GAN stage 3): half the weight traffic and per-CTA set-up, still two CTAs per SM */
#define UMMA_BIG4 4, 2, 8, 3, 1, 1, 1       /* 512-row tiles, NT <= 32 and one input chunk (stage 4) */
#define UMMA_SMALL128 1, 2, 16, 3, 2, 2, 1  /* small grids, NT = 128: 2 taps (32 KB) per weight slot */
#define UMMA_SMALL64 1, 2, 16, 3, 3, 5, 1   /* small grids, NT <= 64: up to 5 taps (40 KB) per slot, 3 activation stages */
#define UMMA_ENC128 1, 3, 16, 3, 2, 1, 1    /* 3-way split (text encoder) */
#define UMMA_ENC64 1, 3, 16, 3, 2, 3, 1
/* persistent launches: one CTA per SM walks several time tiles; two accumulator buffers (all 512 TMEM columns), 16
 * epilogue warps drain one buffer while the MMA warp fills the other; deep weight rings (the SM is not shared) */
#define UMMA_P128 1, 2, 16, 9, 2, 1, 2      /* NT = 128: 9 x 16 KB weight slots */
#define UMMA_P64 2, 2, 16, 6, 2, 2, 2       /* NT <= 64, 256-row tiles: 6 x 2 taps x 8 KB */
#define UMMA_P32 4, 2, 16, 3, 2, 4, 2       /* NT <= 32, 512-row tiles, one input chunk: 3 x 4 taps x 4 KB */

void conv_umma_init_attributes() {
  set_attr_umma<UMMA_BIG>();
  set_attr_umma<UMMA_BIG2>();
  set_attr_umma<UMMA_BIG4>();
  set_attr_umma<UMMA_P128>();
  set_attr_umma<UMMA_P64>();
  set_attr_umma<UMMA_P32>();
  set_attr_umma<UMMA_SMALL128>();
  set_attr_umma<UMMA_SMALL64>();
  set_attr_umma<UMMA_ENC128>();
  set_attr_umma<UMMA_ENC64>();
}

template <int MT, int NSPLIT, int EW, int WST, int AST, int G, int NBUF>
static void launch_umma_cfg(const ConvBatch& pb, int batch, int Lmax, cudaStream_t s, int grid_x = 0) {
  const ConvParams& p = pb.p[0];
  for (int i = 0; i < pb.n; ++i) {  // images are sized for whole tiles of this variant
    const ConvParams& q = pb.p[i];
    const int Lr = (Lmax + MT * 128 - 1) / (MT * 128) * (MT * 128);
    if (q.xp) B2_CHECK(q.xp_rows >= UM_PADL + Lr + UM_HALO, "conv_umma: input plane image too short");
    if (q.yp) B2_CHECK(q.yp_rows >= UM_PADL + Lr * (q.pix > 1 ? q.pix : 1) + UM_TAIL, "conv_umma: output plane image too short");
  }
  const size_t smem = UmmaCfg<MT>::smem_bytes(NSPLIT, p.um_nt, WST, AST, G);
  B2_CHECK(smem <= 227 * 1024, "conv_umma: shared memory budget exceeded");
  const int ntiles = (p.Cout + p.um_nt - 1) / p.um_nt;
  const int m_tiles = (Lmax + MT * 128 - 1) / (MT * 128);
  B2_CHECK(NBUF * NSPLIT * MT * p.um_nt <= 512, "conv_umma: accumulators exceed the 512 TMEM columns");
  B2_CHECK(NBUF == 2 || grid_x == 0, "conv_umma: single-buffer variants take one tile per CTA");
  dim3 grid(grid_x > 0 ? std::min(grid_x, m_tiles) : m_tiles, ntiles, batch * pb.n);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(EW * 32 + 64 + (NBUF == 2 ? 32 : 0));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL: overlap our prologue with the previous kernel
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B2_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<MT, NSPLIT, EW, WST, AST, G, NBUF>, pb));
}

int launch_conv_umma_batch(const ConvParams* ps, int n, int batch, int Lmax, cudaStream_t stream) {
  B2_CHECK(n >= 1 && n <= 3, "conv_umma: 1..3 convs per launch");
  ConvBatch pb;
  pb.n = n;
  pb.zb = batch;
  for (int i = 0; i < n; ++i) {
    const ConvParams& p = ps[i];
    B2_CHECK(p.wu != nullptr && p.um_nt >= 16 && p.um_nt <= 128 && p.um_nt % 16 == 0, "conv_umma: bad weight image");
    // rows of a tile's bulk copies: from the 8-row boundary below its first row, a whole number of 8-row groups
    B2_CHECK(((p.K - 1) * p.dil + ((UM_PADL - p.pad_left) & 7) + 7) / 8 * 8 <= UM_HALO, "conv_umma: receptive field too wide");
    B2_CHECK(p.pad_left <= UM_PADL, "conv_umma: left padding too wide");
    B2_CHECK(p.n_phase == 1, "conv_umma: transposed convs are packed as pixel-shuffle convs");
    if (p.xp) B2_CHECK(p.xp_kc >= (p.Cin + 31) / 32 * 4, "conv_umma: input plane image too narrow");
    B2_CHECK(p.Cout == ps[0].Cout && p.um_nt == ps[0].um_nt && p.um_nsplit == ps[0].um_nsplit && p.Cin == ps[0].Cin,
             "conv_umma: convs of one launch must share their geometry");
    pb.p[i] = p;
  }
  if (Lmax <= 0 || batch <= 0) return 0;
  const ConvParams& p = ps[0];
  const int ntiles = (p.Cout + p.um_nt - 1) / p.um_nt;
  const long long ctas1 = (long long)((Lmax + 127) / 128) * ntiles * batch * n;
  // Tuned constants (profiles/r1_tcgen05_summary.md).  Narrow layers with very many time tiles (HiFi-GAN stages 3 and 4)
  // are bound by per-CTA set-up and by re-streaming the weights for every 128 rows: taller tiles amortise both.
  constexpr long long mt_min = 149;    // grids from this many 128-row CTAs on take the taller-tile variants
  constexpr long long ew16_max = 148;  // grids up to this size use the 16-epilogue-warp variants (one CTA per SM)
  // Persistent launch: one CTA per SM; the 148 CTAs are split between the convs (branches) of the launch in proportion
  // to their cost per tile, max(mainloop ~ K * input chunks, epilogue), and each CTA walks the time tiles of its
  // (n-tile, item, branch) with a stride of gx[branch].
  constexpr int persist_min = 3;       // tiles per CTA from which a persistent launch pays
  constexpr int persist_ctas = 148;
  constexpr int epi_units = 15;        // epilogue cost of a tile in units of one (tap, 32-channel chunk) of main loop
  const long long per = (long long)ntiles * batch;
  if (p.um_nsplit == 2 && p.xp != nullptr && per * n <= persist_ctas) {
    const int mt = p.um_nt <= 32 && p.Cin <= 32 ? 4 : (p.um_nt <= 64 ? 2 : 1);
    const int m_tiles = (Lmax + mt * 128 - 1) / (mt * 128);
    int a[3] = {1, 1, 1}, w[3];
    for (int i = 0; i < n; ++i) w[i] = std::max(ps[i].K * ((ps[i].Cin + 31) / 32), epi_units);
    long long total = per * n;
    while (total + per <= persist_ctas) {
      int worst = 0;
      long long worst_cost = -1;
      for (int i = 0; i < n; ++i) {
        const long long cost = (long long)((m_tiles + a[i] - 1) / a[i]) * w[i];
        if (cost > worst_cost) { worst_cost = cost; worst = i; }
      }
      if (a[worst] >= m_tiles) break;
      a[worst]++;
      total += per;
    }
    int min_tpc = 1 << 30, gmax = 0;
    for (int i = 0; i < n; ++i) {
      const int tpc = (m_tiles + a[i] - 1) / a[i];
      pb.gx[i] = (m_tiles + tpc - 1) / tpc;
      min_tpc = std::min(min_tpc, tpc);
      gmax = std::max(gmax, pb.gx[i]);
    }
    if (min_tpc >= persist_min) {
      if (mt == 4) launch_umma_cfg<UMMA_P32>(pb, batch, Lmax, stream, gmax);
      else if (mt == 2) launch_umma_cfg<UMMA_P64>(pb, batch, Lmax, stream, gmax);
      else launch_umma_cfg<UMMA_P128>(pb, batch, Lmax, stream, gmax);
      B2_CUDA(cudaGetLastError());
      return 1;
    }
    for (int i = 0; i < n; ++i) pb.gx[i] = 0;
  }
  if (p.um_nsplit == 3) {
    if (p.um_nt <= 64) launch_umma_cfg<UMMA_ENC64>(pb, batch, Lmax, stream);
    else launch_umma_cfg<UMMA_ENC128>(pb, batch, Lmax, stream);
  } else if (ctas1 >= mt_min && p.um_nt <= 32 && p.Cin <= 32) {
    launch_umma_cfg<UMMA_BIG4>(pb, batch, Lmax, stream);
  } else if (ctas1 >= mt_min && p.um_nt <= 64) {
    launch_umma_cfg<UMMA_BIG2>(pb, batch, Lmax, stream);
  } else if (ctas1 > ew16_max) {
    launch_umma_cfg<UMMA_BIG>(pb, batch, Lmax, stream);
  } else if (p.um_nt <= 64) {
    launch_umma_cfg<UMMA_SMALL64>(pb, batch, Lmax, stream);
  } else {
    launch_umma_cfg<UMMA_SMALL128>(pb, batch, Lmax, stream);
  }
  B2_CUDA(cudaGetLastError());
  return 1;
}

// Operand image of leaky_relu(((x1 + x2) + x3) / div, slope): the MRF mean that feeds the next upsampler
// (hifi_gan/models.py:191-197 followed by :189), written ONCE instead of being re-gathered by every n-tile CTA of the
// upsampler conv.  Covers every row of the image: zero rows in front (UM_PADL), behind each item's length and in the tail.
// Same expression order as the gather path of conv_umma_kernel, so both paths produce identical bits.
__global__ void __launch_bounds__(256) mean3_image_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                          const float* __restrict__ x3, long long bs, int ld, float div,
                                                          float slope, const int* __restrict__ lens, int len_mul,
                                                          unsigned char* __restrict__ img, long long img_bs, int kc_total,
                                                          int rows) {
  const int b = blockIdx.z, kc = blockIdx.y;
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= rows) return;
  const int t = r - UM_PADL;
  const bool ok = t >= 0 && t < lens[b] * len_mul;
  float v[8];
  const size_t base = (size_t)b * bs + (size_t)(kc * 8) * ld + (ok ? t : 0);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const size_t off = base + (size_t)e * ld;
    float a = x1[off];
    if (x2) a = (a + x2[off]) + x3[off];  // x2 and x3 come together (MRF mean); alone: the image of x1 itself
    if (div != 1.f) a = a / div;
    a = a > 0.f ? a : a * slope;
    v[e] = ok ? a : 0.f;
  }
  uint4 q[2];
  split8<2>(v, q);
  unsigned char* o = img + (size_t)b * img_bs + ((size_t)kc * rows + r) * 16;
  *reinterpret_cast<uint4*>(o) = q[0];
  *reinterpret_cast<uint4*>(o + (size_t)kc_total * rows * 16) = q[1];
}

void launch_mean3_image(const float* x1, const float* x2, const float* x3, long long bs, int ld, int C, float div, float slope,
                        const int* lens, int len_mul, int batch, void* img, long long img_bs, int kc_total, int rows,
                        cudaStream_t s) {
  B2_CHECK(C % 8 == 0 && C / 8 <= kc_total, "mean image: channel count does not fit the image");
  if (batch <= 0 || rows <= 0) return;
  dim3 grid((rows + 255) / 256, C / 8, batch);
  mean3_image_kernel<<<grid, 256, 0, s>>>(x1, x2, x3, bs, ld, div, slope, lens, len_mul, reinterpret_cast<unsigned char*>(img),
                                          img_bs, kc_total, rows);
  B2_CUDA(cudaGetLastError());
}

int launch_conv_umma(const ConvParams& p, int batch, int Lmax, cudaStream_t stream) {
  return launch_conv_umma_batch(&p, 1, batch, Lmax, stream);
}

}  // namespace b2
