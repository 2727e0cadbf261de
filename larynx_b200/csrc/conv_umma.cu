// conv_umma.cu — launcher of the tcgen05 implicit-GEMM conv (conv_umma.cuh).
#include "conv_umma.cuh"
#include "kernels.cuh"

namespace b2 {

template <int MT, int NSPLIT, int EW, int WST>
static void set_attr_umma() {
  B2_CUDA(cudaFuncSetAttribute(conv_umma_kernel<MT, NSPLIT, EW, WST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
}

void conv_umma_init_attributes() {
  set_attr_umma<1, 2, 8, 3>();
  set_attr_umma<2, 2, 8, 6>();
  set_attr_umma<1, 2, 16, 6>();
  set_attr_umma<1, 3, 16, 4>();
}

template <int MT, int NSPLIT, int EW, int WST>
static void launch_umma_cfg(const ConvParams& p, int batch, int Lmax, cudaStream_t s) {
  const size_t smem = UmmaCfg<MT>::smem_bytes(NSPLIT, p.um_nt, WST);
  B2_CHECK(smem <= 227 * 1024, "conv_umma: shared memory budget exceeded");
  const int ntiles = (p.Cout + p.um_nt - 1) / p.um_nt;
  dim3 grid((Lmax + MT * 128 - 1) / (MT * 128), ntiles, batch);
  conv_umma_kernel<MT, NSPLIT, EW, WST><<<grid, EW * 32 + 64, smem, s>>>(p);
}

int launch_conv_umma(const ConvParams& p, int batch, int Lmax, cudaStream_t stream) {
  B2_CHECK(p.wu != nullptr && p.um_nt >= 16 && p.um_nt <= 128 && p.um_nt % 16 == 0, "conv_umma: bad weight image");
  B2_CHECK((p.K - 1) * p.dil <= UM_HALO, "conv_umma: receptive field too wide");
  B2_CHECK(p.pad_left <= UM_PADL, "conv_umma: left padding too wide");
  B2_CHECK(p.n_phase == 1, "conv_umma: transposed convs are packed as pixel-shuffle convs");
  if (p.xp) B2_CHECK(p.xp_kc >= (p.Cin + 31) / 32 * 4, "conv_umma: input plane image too narrow");
  if (Lmax <= 0 || batch <= 0) return 0;
  const int ntiles = (p.Cout + p.um_nt - 1) / p.um_nt;
  const long long ctas1 = (long long)((Lmax + 127) / 128) * ntiles * batch;
  const long long ctas2 = (long long)((Lmax + 255) / 256) * ntiles * batch;
  if (p.um_nsplit == 3) {
    launch_umma_cfg<1, 3, 16, 4>(p, batch, Lmax, stream);
  } else if (ctas2 >= 148) {
    launch_umma_cfg<2, 2, 8, 6>(p, batch, Lmax, stream);
  } else if (ctas1 > 148) {
    launch_umma_cfg<1, 2, 8, 3>(p, batch, Lmax, stream);
  } else {
    launch_umma_cfg<1, 2, 16, 6>(p, batch, Lmax, stream);
  }
  B2_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace b2
