// rbpair_umma.cu — launcher of the fused ResBlock1 step kernel (rbpair_umma.cuh).
#include <algorithm>

#include "kernels.cuh"
#include "rbpair_umma.cuh"

namespace b2 {

void rbpair_umma_init_attributes() {
  B2_CUDA(cudaFuncSetAttribute(rbpair_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  B2_CUDA(cudaFuncSetAttribute(rbpair_umma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
}

// tile height (in 128-row MMA tiles) of the fused kernel for this geometry, or 0 when it does not apply
static int rbpair_mt(const ConvParams& c1, const ConvParams& c2) {
  const int C = c1.Cin;
  if (C % 32 != 0 || C > 64) return 0;
  if (c1.Cout != C || c2.Cin != C || c2.Cout != C || c1.um_nt != C || c2.um_nt != C) return 0;
  if (c1.um_nsplit != 2 || c2.um_nsplit != 2 || !c1.wu || !c2.wu || !c1.xp || !c2.rp) return 0;
  if (c2.dil != 1 || c2.K - 1 > RP_TPAD || (c1.K - 1) * c1.dil > UM_HALO - 16 || c1.pad_left + c2.pad_left > UM_PADL) return 0;
  if (c1.epi != EPI_PLAIN || c2.epi != EPI_PLAIN || c1.pix > 1 || c2.pix > 1 || c2.act != ACT_NONE) return 0;
  if (c1.lens != c2.lens || c1.len_mul != c2.len_mul) return 0;
  if (c2.yp && (c2.yp == c1.xp || c2.yp == c2.rp)) return 0;  // tiles overlap: the input image must not be rewritten
  return C <= 32 ? 4 : 2;
}

bool rbpair_umma_supported(const ConvParams* c1, const ConvParams* c2, int n) {
  if (n < 1 || n > 3) return false;
  const int mt = rbpair_mt(c1[0], c2[0]);
  if (mt == 0) return false;
  for (int i = 1; i < n; ++i)
    if (rbpair_mt(c1[i], c2[i]) != mt || c1[i].Cin != c1[0].Cin) return false;
  return true;
}

template <int MT>
static void launch_cfg(const PairBatch& pb, int batch, int gmax, cudaStream_t s) {
  const size_t smem = RbPairCfg<MT>::smem_bytes(pb.c2[0].um_nt, pb.c2[0].Cin);
  B2_CHECK(smem <= 227 * 1024, "rbpair_umma: shared memory budget exceeded");
  B2_CHECK(2 * MT * pb.c2[0].um_nt <= 256, "rbpair_umma: accumulators exceed their TMEM columns");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gmax, 1, batch * pb.n);
  cfg.blockDim = dim3(RP_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL: overlap our prologue with the previous kernel
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B2_CUDA(cudaLaunchKernelEx(&cfg, rbpair_umma_kernel<MT>, pb));
}

// n <= 3 independent ResBlock steps (the MRF branches of a stage) of identical geometry in one persistent launch.
// The 148 CTAs are split between the branches in proportion to their cost per tile.
int launch_rbpair_umma(const ConvParams* c1, const ConvParams* c2, int n, int batch, int Lmax, cudaStream_t stream) {
  B2_CHECK(rbpair_umma_supported(c1, c2, n), "rbpair_umma: unsupported geometry");
  if (Lmax <= 0 || batch <= 0) return 0;
  const int mt = rbpair_mt(c1[0], c2[0]);
  PairBatch pb;
  pb.n = n;
  pb.zb = batch;
  int tiles[3], w[3], a[3] = {1, 1, 1};
  for (int i = 0; i < n; ++i) {
    pb.c1[i] = c1[i];
    pb.c2[i] = c2[i];
    const int S = mt * 128 - (c2[i].K - 1);
    tiles[i] = (Lmax + S - 1) / S;
    const int Lr = tiles[i] * S + mt * 128;
    B2_CHECK(c1[i].xp_rows >= UM_PADL + Lr + UM_HALO, "rbpair_umma: input plane image too short");
    if (c2[i].yp) B2_CHECK(c2[i].yp_rows >= UM_PADL + Lr + UM_TAIL, "rbpair_umma: output plane image too short");
    const int nch = (c1[i].Cin + 31) / 32;
    w[i] = (c1[i].K + c2[i].K) * nch * mt + 8 * mt;  // main loops + the two epilogues, per tile
  }
  constexpr int persist_ctas = 148;
  long long total = (long long)batch * n;
  B2_CHECK(total <= persist_ctas, "rbpair_umma: more (branch, item) pairs than SMs");
  while (total + batch <= persist_ctas) {
    int worst = -1;
    long long worst_cost = -1;
    for (int i = 0; i < n; ++i) {
      if (a[i] >= tiles[i]) continue;
      const long long cost = (long long)((tiles[i] + a[i] - 1) / a[i]) * w[i];
      if (cost > worst_cost) { worst_cost = cost; worst = i; }
    }
    if (worst < 0) break;
    a[worst]++;
    total += batch;
  }
  int gmax = 0;
  for (int i = 0; i < n; ++i) {
    const int tpc = (tiles[i] + a[i] - 1) / a[i];
    pb.gx[i] = (tiles[i] + tpc - 1) / tpc;
    gmax = std::max(gmax, pb.gx[i]);
  }
  if (mt == 4) launch_cfg<4>(pb, batch, gmax, stream);
  else launch_cfg<2>(pb, batch, gmax, stream);
  B2_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace b2
