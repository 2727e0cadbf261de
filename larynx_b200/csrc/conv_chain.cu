// conv_chain.cu — launcher of the chained conv kernel (conv_chain.cuh).  Round-2 work in progress: see the header.
#include <stdlib.h>

#include <vector>

#include "conv_chain.cuh"
#include "kernels.cuh"

namespace b2 {

#define CHAIN_CFG 16, 3, 3, 5 /* EW, WST, AST, G — the shape of UMMA_SMALL64, the variant these layers use one by one */

void conv_chain_init_attributes() {
  B2_CUDA(cudaFuncSetAttribute(conv_chain_kernel<CHAIN_CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)ChainCfg<CHAIN_CFG>::smem_bytes));
}

// true if the layers can run as one chain on this device: plane in/out, 2-way split, NT <= 64, grid fits the SMs
bool conv_chain_supported(const ConvParams* layers, int n_layers, int batch, int Lmax) {
  if (n_layers <= 0 || batch <= 0 || Lmax <= 0) return false;
  int max_ntiles = 0;
  for (int i = 0; i < n_layers; ++i) {
    const ConvParams& p = layers[i];
    if (!p.wu || !p.xp || p.um_nsplit != 2 || p.um_nt > CH_NT_MAX || p.um_nt % 16 != 0 || p.n_phase != 1 || p.pix > 1) return false;
    if ((p.K - 1) * p.dil > UM_HALO || p.pad_left > UM_PADL) return false;
    if (p.y == nullptr && p.yp == nullptr) return false;
    max_ntiles = std::max(max_ntiles, (p.Cout + p.um_nt - 1) / p.um_nt);
  }
  const long long ctas = (long long)((Lmax + 127) / 128) * max_ntiles * batch;
  int dev = 0, sms = 0;
  B2_CUDA(cudaGetDevice(&dev));
  B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return ctas <= sms;  // cooperative launch: the whole grid must be resident, one CTA per SM
}

// layers_d: device copy of `layers` (n_layers entries); counter_d: n_layers uints (zeroed here).  Returns the launches made.
int launch_conv_chain(const ConvParams* layers, const ConvParams* layers_d, int n_layers, unsigned int* counter_d, int batch,
                      int Lmax, cudaStream_t s) {
  B2_CHECK(conv_chain_supported(layers, n_layers, batch, Lmax), "conv_chain: unsupported chain");
  conv_chain_init_attributes();  // per launch: cheap, and keeps this experimental path out of b200tts_call_create
  int max_ntiles = 0;
  for (int i = 0; i < n_layers; ++i) max_ntiles = std::max(max_ntiles, (layers[i].Cout + layers[i].um_nt - 1) / layers[i].um_nt);
  dim3 grid((Lmax + 127) / 128, max_ntiles, batch);
  B2_CUDA(cudaMemsetAsync(counter_d, 0, sizeof(unsigned int) * n_layers, s));
  ChainParams cp;
  cp.layers = layers_d;
  cp.n_layers = n_layers;
  cp.counter = counter_d;
  cp.n_ctas = grid.x * grid.y * grid.z;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(16 * 32 + 96);
  cfg.dynamicSmemBytes = ChainCfg<CHAIN_CFG>::smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs resident at once: the layer barrier spins
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B2_CUDA(cudaLaunchKernelEx(&cfg, conv_chain_kernel<CHAIN_CFG>, cp));
  B2_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace b2
