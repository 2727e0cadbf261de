// flow_chain.cu — launcher of the persistent flow-decoder kernel (flow_chain.cuh).
#include <algorithm>

#include "flow_chain.cuh"
#include "kernels.cuh"

namespace b2 {

void flow_chain_init_attributes() {
  B2_CUDA(cudaFuncSetAttribute(flow_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FC_SMEM));
}

// geometry the kernel is built for (every released voice: SURVEY.md §2.2); anything else takes the per-layer launches
bool flow_chain_geometry_ok(int half, int Hd, int K, int dil_rate, int n_wn, int nt_hidden, int nt_end) {
  if (Hd != FC_NT * FC_YT / 2 || nt_hidden != FC_NT) return false;
  if (n_wn < 1 || n_wn > FC_MAX_WN || K < 1 || K > FC_KMAX) return false;
  int dil = 1;
  for (int l = 1; l < n_wn; ++l) dil *= dil_rate;
  if ((K - 1) * dil > FC_HALO || (K * dil - dil) / 2 > UM_PADL) return false;
  if (half % 8 != 0 || 2 * half > 256 || nt_end < 16 || nt_end > FC_NT || nt_end % 32 != 0) return false;
  if ((2 * half + nt_end - 1) / nt_end > FC_YT) return false;
  return true;
}

// CTAs that can be resident at once (cooperative launch: the per-utterance barriers spin)
static int resident_ctas() {
  static const int n = [] {
    int dev = 0, sms = 0, per_sm = 0;
    B2_CUDA(cudaGetDevice(&dev));
    B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flow_chain_kernel, FC_THREADS, FC_SMEM));
    return sms * per_sm;
  }();
  return n;
}

bool flow_chain_fits(int Lmax) {
  if (Lmax <= 0) return false;
  return (long long)((Lmax + 127) / 128) * FC_YT <= resident_ctas();
}

// fp.counter: batch * n_blocks * (2 + 2 n_wn) uints (zeroed here).  Returns the launches made.
int launch_flow_chain(const FlowChainParams& fp, int Lmax, cudaStream_t s) {
  B2_CHECK(flow_chain_fits(Lmax), "flow_chain: utterance too long for a resident CTA group");
  B2_CHECK(fp.rows >= UM_PADL + (Lmax + 127) / 128 * 128 + UM_TAIL, "flow_chain: plane images too short");
  const int m_tiles = (Lmax + 127) / 128;
  const int groups = std::max(1, std::min(fp.batch, resident_ctas() / (m_tiles * FC_YT)));
  const int n_steps = 2 + 2 * fp.n_wn;
  B2_CUDA(cudaMemsetAsync(fp.counter, 0, sizeof(unsigned int) * (size_t)fp.batch * fp.n_blocks * n_steps, s));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(m_tiles, FC_YT, groups);
  cfg.blockDim = dim3(FC_THREADS);
  cfg.dynamicSmemBytes = FC_SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs resident at once: the layer barriers spin
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  spin_gate_begin(s);
  const cudaError_t le = cudaLaunchKernelEx(&cfg, flow_chain_kernel, fp);
  spin_gate_end(s);
  B2_CUDA(le);
  B2_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace b2
