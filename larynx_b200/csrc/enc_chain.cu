// enc_chain.cu — launcher of the persistent text-encoder kernel (enc_chain.cuh) + the embedding kernel that feeds it.
#include <algorithm>

#include "enc_chain.cuh"
#include "kernels.cuh"

namespace b2 {

void enc_chain_init_attributes() {
  B2_CUDA(cudaFuncSetAttribute(enc_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EC_SMEM));
}

// CTAs that can be resident at once (cooperative launch: the per-utterance barriers spin)
static int enc_resident_ctas() {
  static const int n = [] {
    int dev = 0, sms = 0, per_sm = 0;
    B2_CUDA(cudaGetDevice(&dev));
    B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, enc_chain_kernel, EC_THREADS, EC_SMEM));
    return sms * per_sm;
  }();
  return n;
}

// the attention step handles one 128-query tile per utterance
bool enc_chain_fits(int max_tx) { return max_tx > 0 && max_tx <= 128 && EC_YT <= enc_resident_ctas(); }

// steps: host copy of the table (n_steps entries, validated here); steps_d: its device copy.  fp.counter is zeroed here.
int launch_enc_chain(const EncStep* steps, const EncChainParams& fp, int max_tx, cudaStream_t s) {
  B2_CHECK(enc_chain_fits(max_tx), "enc_chain: utterance too long for the chained encoder");
  B2_CHECK(fp.n_steps >= 1 && fp.n_steps <= EC_MAX_STEPS, "enc_chain: too many steps");
  const int m_tiles = (max_tx + 127) / 128;
  B2_CHECK(fp.rows >= UM_PADL + m_tiles * 128 + UM_TAIL + 128, "enc_chain: operand images too short");
  for (int i = 0; i < fp.n_steps; ++i) {
    const EncStep& st = steps[i];
    if (st.type == ES_CONV) {
      B2_CHECK(st.n_conv >= 1 && st.n_conv <= 2, "enc_chain: 1 or 2 convs per step");
      int rows_used = 0;
      for (int j = 0; j < st.n_conv; ++j) {
        const EncConvOp& o = st.conv[j];
        const int nch = (o.Cin + UM_BKC - 1) / UM_BKC;
        B2_CHECK(o.w && o.xp && (o.y || o.yp), "enc_chain: conv without operands");
        B2_CHECK(o.planes == 1 || o.planes == 3, "enc_chain: 1 or 3 activation planes");
        B2_CHECK(o.NT >= 16 && o.NT <= EC_NT_MAX && o.NT % 16 == 0, "enc_chain: bad N tile");
        B2_CHECK(o.K >= 1 && o.K - 1 <= EC_HALO && o.pad <= UM_PADL, "enc_chain: receptive field too wide");
        B2_CHECK(o.ksplit >= 1 && nch % o.ksplit == 0, "enc_chain: K-split must divide the input chunks");
        B2_CHECK(o.ntiles == (o.Cout + o.NT - 1) / o.NT, "enc_chain: N tiles do not cover the output channels");
        B2_CHECK(o.ksplit == 1 || (o.y && !o.yp && !o.bias && !o.act), "enc_chain: K-split convs write raw fp32 partials");
        B2_CHECK(o.y0 >= rows_used && o.y0 + o.ntiles * o.ksplit <= EC_YT, "enc_chain: conv does not fit the CTA rows");
        B2_CHECK(o.xp_kc >= nch * UM_KCH, "enc_chain: input image too narrow");
        rows_used = o.y0 + o.ntiles * o.ksplit;
      }
    } else if (st.type == ES_REDUCE) {
      const EncReduceOp& r = st.red;
      B2_CHECK(r.x && r.parts >= 1 && r.C % 8 == 0 && r.C <= 256 && (r.y || r.yp), "enc_chain: bad reduce step");
    } else {
      const EncAttnOp& a = st.att;
      B2_CHECK(st.type == ES_ATTN && a.H % a.n_heads == 0 && a.H / a.n_heads == EA_DK && a.window <= 4 &&
                   a.y0 + a.n_heads <= EC_YT && a.relk_img && a.relv_img && (a.planes == 1 || a.planes == 3),
               "enc_chain: unsupported attention");
    }
  }
  const int groups = std::max(1, std::min(fp.batch, enc_resident_ctas() / (m_tiles * EC_YT)));
  B2_CUDA(cudaMemsetAsync(fp.counter, 0, sizeof(unsigned int) * (size_t)fp.batch * fp.n_steps, s));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(m_tiles, EC_YT, groups);
  cfg.blockDim = dim3(EC_THREADS);
  cfg.dynamicSmemBytes = EC_SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs resident at once: the step barriers spin
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  spin_gate_begin(s);
  const cudaError_t le = cudaLaunchKernelEx(&cfg, enc_chain_kernel, fp);
  spin_gate_end(s);
  B2_CUDA(le);
  B2_CUDA(cudaGetLastError());
  return 1;
}

// ------------------------------------------------------------------------------------------------
// embedding (glow_tts/models.py:119-120) written as fp32 [H][ld] AND as the 3-plane operand image the first prenet conv
// bulk-loads: x[b][c][t] = emb[ids[b][t]][c] * sqrt(H); one thread per (image row, 8-channel group); rows outside
// [0, len) are zero rows of the image.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) embed_image_kernel(const long long* __restrict__ ids, const int* __restrict__ lens, int max_tx,
                                                          const float* __restrict__ emb, int n_vocab, int H, float scale,
                                                          float* __restrict__ x, long long x_bs, int x_ld, unsigned char* __restrict__ img,
                                                          long long img_bs, int kc_total, int rows, int* __restrict__ bad_id) {
  const int b = blockIdx.z, g = blockIdx.y;
  const int r = blockIdx.x * 128 + threadIdx.x;
  if (r >= rows) return;
  const int t = r - UM_PADL;
  const bool ok = t >= 0 && t < lens[b];
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  if (ok) {
    long long id = ids[(size_t)b * max_tx + t];
    if (id < 0 || id >= n_vocab) {  // nn.Embedding raises IndexError: flag it, the host throws
      if (g == 0) atomicExch(bad_id, 1);
      id = 0;
    }
    const float4* e4 = reinterpret_cast<const float4*>(emb + (size_t)id * H + g * 8);
    const float4 a = e4[0], c = e4[1];
    v[0] = a.x * scale; v[1] = a.y * scale; v[2] = a.z * scale; v[3] = a.w * scale;
    v[4] = c.x * scale; v[5] = c.y * scale; v[6] = c.z * scale; v[7] = c.w * scale;
    float* xb = x + (size_t)b * x_bs + (size_t)(g * 8) * x_ld + t;
#pragma unroll
    for (int e = 0; e < 8; ++e) xb[(size_t)e * x_ld] = v[e];
  }
  uint4 q[3];
  split8<3>(v, q);
  unsigned char* o = img + (size_t)b * img_bs + ((size_t)g * rows + r) * 16;
  const size_t pstride = (size_t)kc_total * rows * 16;
  *reinterpret_cast<uint4*>(o) = q[0];
  *reinterpret_cast<uint4*>(o + pstride) = q[1];
  *reinterpret_cast<uint4*>(o + 2 * pstride) = q[2];
}

void launch_embed_image(const long long* ids, const int* lens, int batch, int max_tx, const float* emb, int n_vocab, int H, float* x,
                        long long x_bs, int x_ld, void* img, long long img_bs, int kc_total, int rows, int* bad_id, cudaStream_t s) {
  B2_CHECK(H % 8 == 0 && H / 8 <= kc_total, "embed: channel count does not fit the image");
  if (batch <= 0 || rows <= 0) return;
  dim3 grid((rows + 127) / 128, H / 8, batch);
  embed_image_kernel<<<grid, 128, 0, s>>>(ids, lens, max_tx, emb, n_vocab, H, sqrtf((float)H), x, x_bs, x_ld,
                                          reinterpret_cast<unsigned char*>(img), img_bs, kc_total, rows, bad_id);
  B2_CUDA(cudaGetLastError());
}

}  // namespace b2
