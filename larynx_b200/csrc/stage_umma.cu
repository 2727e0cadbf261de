// stage_umma.cu — host side of the per-stage dataflow kernel (stage_umma.cuh): work partition + launch.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "kernels.cuh"
#include "stage_umma.cuh"

#ifndef SG_COST_A2
#define SG_COST_A2 55  // ... of a 256 x 64 tile (x 1/2)
#endif
#ifndef SG_COST_AP
#define SG_COST_AP 100  // ... of a ResBlock step (c1 + c2) on a 512 x 32 tile
#endif
#ifndef SG_COST_A1
#define SG_COST_A1 26  // modelled cost (10 ns) of one (tap, 32-channel chunk) of a 128 x 128 tile
#endif

namespace b2 {

void stage_umma_init_attributes() {
  B2_CUDA(cudaFuncSetAttribute(stage_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  B2_CUDA(cudaFuncSetAttribute(stage_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  B2_CUDA(cudaFuncSetAttribute(stage_umma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  B2_CUDA(cudaFuncSetAttribute(stage_pair_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
}

static int stage_mt(int nt) { return nt <= 32 ? 4 : (nt <= 64 ? 2 : 1); }

template <int MT>
static int resident_ctas(size_t smem) {
  // (per process: every GPU of a box is the same part; the query costs several microseconds per launch otherwise)
  static std::mutex m;
  static size_t last_smem = 0;
  static int last = 0;
  std::lock_guard<std::mutex> g(m);
  if (last > 0 && last_smem == smem) return last;
  int dev = 0, sms = 0, per_sm = 0;
  B2_CUDA(cudaGetDevice(&dev));
  B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stage_umma_kernel<MT>, SG_EW * 32 + 128, smem));
  last_smem = smem;
  last = sms * per_sm;
  return last;
}

// Contiguous, equal-cost tile ranges.  strips: (col, item, n-tile, tiles, cost per tile) in launch order; a tile goes to the CTA
// its cumulative cost (at the tile's middle) falls into, so every CTA's total differs from total / nctas by less than one tile
// and consecutive tiles of a strip stay together.  seg_off[c] .. seg_off[c+1] are CTA c's segments (a CTA the walk skips over
// — a single tile dearer than a whole share — gets none).  Returns the number of segments.
struct StageStrip {
  int col, b, ntile, tiles;
  long long cost;
  int t0 = 0;  // index of the strip's first tile
};

// Item order inside a level.  Contiguous ranges of ONE branch make a CTA's next level wait for the tile it has just issued (its
// own neighbour in time): MMA + epilogue + publication of that tile = ~11 us of a 12.5 us item in profiles/r2d_stage_trace.log.
// Striped: the tiles of a level are ordered (item, time tile, branch, n-tile) and cut into equal-cost ranges in THAT order, so a
// CTA alternates between the branches and the tiles a next-level item needs from its own CTA were issued several items earlier;
// the ones it needs from its neighbours are the last / first items of their ranges.  One segment per tile: used while the
// segment table holds them (SG_MAX_SEGS), which covers single utterances and small batches — large batches have many tiles per
// CTA and level, and the contiguous order (better weight reuse in L1/L2) loses nothing there.
static bool g_striped = true;
void stage_set_striped(bool on) { g_striped = on; }
constexpr int SG_MAX_SEGS = 2048;
int stage_partition(const StageStrip* strips, int n_strips, int nctas, int* seg_off, StageSeg* segs, int max_segs) {
  long long total = 0;
  for (int i = 0; i < n_strips; ++i) total += (long long)strips[i].tiles * strips[i].cost;
  B2_CHECK(total > 0 && nctas >= 1, "stage partition: no work");
  int nseg = 0, cur_cta = -1;
  long long cum = 0;
  for (int i = 0; i < n_strips; ++i) {
    const StageStrip& st = strips[i];
    bool open = false;
    for (int t = 0; t < st.tiles; ++t) {
      const int cta = (int)std::min<long long>(nctas - 1, (cum + st.cost / 2) * nctas / total);
      cum += st.cost;
      if (cta != cur_cta) {
        for (int c = cur_cta + 1; c <= cta; ++c) seg_off[c] = nseg;
        cur_cta = cta;
        open = false;
      }
      if (!open) {
        B2_CHECK(nseg < max_segs, "stage partition: too many segments");
        segs[nseg].col = st.col; segs[nseg].b = st.b; segs[nseg].ntile = st.ntile; segs[nseg].t_begin = st.t0 + t; segs[nseg].t_end = st.t0 + t;
        ++nseg;
        open = true;
      }
      segs[nseg - 1].t_end = st.t0 + t + 1;
    }
  }
  for (int c = cur_cta + 1; c <= nctas; ++c) seg_off[c] = nseg;
  return nseg;
}

bool stage_umma_supported(const ConvParams* convs, int n_levels, int n_cols) {
  const int n = n_levels * n_cols;
  if (n < 1 || n > SG_MAX_CONVS) return false;
  const ConvParams& p0 = convs[0];
  const int C = p0.Cin, NT = p0.um_nt;
  if (C % 32 != 0 || NT < 32 || NT > 128 || NT % 32 != 0 || p0.Cout % NT != 0) return false;
  const int mt = stage_mt(NT);
  for (int i = 0; i < n; ++i) {
    const ConvParams& p = convs[i];
    if (p.Cin != C || p.Cout != p0.Cout || p.um_nt != NT || p.um_nsplit != 2 || !p.wu || !p.xp || !p.bias) return false;
    if (p.epi != EPI_PLAIN || p.pix > 1 || p.n_phase != 1 || p.res || p.x2 || p.x3) return false;
    if (!(p.yp || p.y) || (p.yp && p.yp_nsplit != 2)) return false;
    if (((p.K - 1) * p.dil + ((UM_PADL - p.pad_left) & 7) + 7) / 8 * 8 > UM_HALO || p.pad_left > UM_PADL) return false;
    if ((p.K - 1) * p.dil >= mt * 128) return false;  // a tile only depends on its producer's three neighbouring tiles
    if (p.lens != p0.lens || p.len_mul != p0.len_mul || !p.lens) return false;
    if (p.xp_kc < (C + 31) / 32 * 4) return false;
  }
  size_t smem = 0;
  if (mt == 1) smem = StageCfg<1>::smem_bytes(NT, p0.Cout);
  else if (mt == 2) smem = StageCfg<2>::smem_bytes(NT, p0.Cout);
  else smem = StageCfg<4>::smem_bytes(NT, p0.Cout);
  return smem <= 227 * 1024;
}

size_t stage_umma_table_bytes() {
  return sizeof(ConvParams) * SG_MAX_CONVS + sizeof(int) * SG_MAX_CONVS + sizeof(int) * 1024 + sizeof(StageSeg) * 2048;
}

// convs: host table [n_levels][n_cols]; lens_h: host copy of the item lengths (x len_mul = rows); tab_h / tab_d: pinned host and
// device scratch of stage_umma_table_bytes() (the host block must stay untouched until the stream has consumed the copy);
// flags_d: at least n_levels * n_cols * batch * ceil(Lmax / (MT*128)) counters.
int launch_stage_umma(const ConvParams* convs, int n_levels, int n_cols, int batch, const int* lens_h, void* tab_h, void* tab_d,
                      unsigned int* flags_d, size_t flags_cap, cudaStream_t s, bool trace) {
  B2_CHECK(stage_umma_supported(convs, n_levels, n_cols), "stage_umma: unsupported stage geometry");
  const int n = n_levels * n_cols;
  const ConvParams& p0 = convs[0];
  const int NT = p0.um_nt, ntiles = p0.Cout / NT, mt = stage_mt(NT);
  const int nch = (p0.Cin + 31) / 32;
  // ---- modelled cost of one tile of a column through all levels (units: 10 ns).  Fitted to per-CTA timelines of the full
  //      launch (profiles/r2d_stage_trace.log: all 148 CTAs busy, L2 shared): the wide stages are main-loop bound (k = 3: epilogue
  //      bound), the 256-row tiles of the 64-channel stage overlap the two poorly (sum, not max), the
  //      512-row tiles of the 32-channel stage are epilogue bound for every kernel size.
  std::vector<long long> wcol(n_cols, 0);
  for (int j = 0; j < n_cols; ++j)
    for (int lv = 0; lv < n_levels; ++lv) {
      const long long kc = (long long)convs[lv * n_cols + j].K * nch;
      long long cost;
      if (mt == 1) cost = std::max((long long)SG_COST_A1 * kc + 120, 600LL);
      else if (mt == 2) cost = (long long)SG_COST_A2 * kc / 2 + 370;
      else cost = std::max(50 * kc + 150, 600LL);
      wcol[j] += cost;
    }
  int tiles_max = 0;
  long long total = 0, n_tiles_all = 0;
  std::vector<int> mtl(batch);
  for (int b = 0; b < batch; ++b) {
    const int Lb = lens_h[b] * p0.len_mul;
    mtl[b] = (Lb + mt * 128 - 1) / (mt * 128);
    tiles_max = std::max(tiles_max, mtl[b]);
    for (int j = 0; j < n_cols; ++j) {
      total += (long long)mtl[b] * ntiles * wcol[j];
      n_tiles_all += (long long)mtl[b] * ntiles;
    }
  }
  if (n_tiles_all == 0) return 0;
  B2_CHECK((size_t)n * batch * tiles_max <= flags_cap, "stage_umma: flag buffer too small");
  size_t smem = 0;
  int resident = 0;
  if (mt == 1) { smem = StageCfg<1>::smem_bytes(NT, p0.Cout); resident = resident_ctas<1>(smem); }
  else if (mt == 2) { smem = StageCfg<2>::smem_bytes(NT, p0.Cout); resident = resident_ctas<2>(smem); }
  else { smem = StageCfg<4>::smem_bytes(NT, p0.Cout); resident = resident_ctas<4>(smem); }
  B2_CHECK(resident >= 1, "stage_umma: kernel does not fit an SM");
  const int nctas = (int)std::min<long long>(std::min(resident, 1000), n_tiles_all);
  // ---- contiguous, equal-cost ranges: tile -> CTA by its cumulative cost
  char* th = reinterpret_cast<char*>(tab_h);
  ConvParams* convs_t = reinterpret_cast<ConvParams*>(th);
  int* dep_t = reinterpret_cast<int*>(th + sizeof(ConvParams) * SG_MAX_CONVS);
  int* off_t = dep_t + SG_MAX_CONVS;
  StageSeg* seg_t = reinterpret_cast<StageSeg*>(off_t + 1024);
  B2_CHECK(nctas + 1 <= 1024, "stage_umma: too many CTAs");
  memcpy(convs_t, convs, sizeof(ConvParams) * n);
  for (int lv = 0; lv < n_levels; ++lv)
    for (int j = 0; j < n_cols; ++j) dep_t[lv * n_cols + j] = lv > 0 ? (lv - 1) * n_cols + j : -1;
  std::vector<StageStrip> strips;
  if (g_striped && n_tiles_all + nctas <= SG_MAX_SEGS) {
    for (int b = 0; b < batch; ++b)
      for (int t = 0; t < mtl[b]; ++t)
        for (int j = 0; j < n_cols; ++j)
          for (int nt = 0; nt < ntiles; ++nt) strips.push_back({j, b, nt, 1, wcol[j], t});
  } else {
    for (int j = 0; j < n_cols; ++j)
      for (int b = 0; b < batch; ++b)
        for (int nt = 0; nt < ntiles; ++nt)
          if (mtl[b] > 0) strips.push_back({j, b, nt, mtl[b], wcol[j]});
  }
  const int nseg = stage_partition(strips.data(), (int)strips.size(), nctas, off_t, seg_t, SG_MAX_SEGS);
  char* td = reinterpret_cast<char*>(tab_d);
  // only the bytes in use: conv table + dep + offsets + the segments actually written (~10 of 52 KB)
  B2_CUDA(cudaMemcpyAsync(td, th, (size_t)(reinterpret_cast<char*>(seg_t + nseg) - th), cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaMemsetAsync(flags_d, 0, sizeof(unsigned int) * (size_t)n * batch * tiles_max, s));
  StageParams sp;
  sp.convs = reinterpret_cast<const ConvParams*>(td);
  sp.convs1 = nullptr;
  sp.dep = reinterpret_cast<const int*>(td + sizeof(ConvParams) * SG_MAX_CONVS);
  sp.seg_off = sp.dep + SG_MAX_CONVS;
  sp.segs = reinterpret_cast<const StageSeg*>(sp.seg_off + 1024);
  sp.flags = flags_d;
  sp.n_levels = n_levels; sp.n_cols = n_cols; sp.zb = batch; sp.tiles_max = tiles_max; sp.ntiles = ntiles;
  sp.trace = nullptr; sp.trace_cta = 0; sp.trace_items = 0;
  static unsigned long long* tr = nullptr;
  constexpr int TR_ITEMS = 64;
  int trace_cta = 0;
  if (trace) {  // debug (single-threaded use): per-item stamps of two CTAs, printed after the launch
    if (!tr) B2_CUDA(cudaMallocManaged(&tr, sizeof(unsigned long long) * 8 * TR_ITEMS));
    B2_CUDA(cudaDeviceSynchronize());
    memset(tr, 0, sizeof(unsigned long long) * 8 * TR_ITEMS);
    const char* e = getenv("B200TTS_TRACE_CTA");  // debug only: which CTA to trace (default: the middle one)
    trace_cta = e ? std::min(std::max(atoi(e), 0), nctas - 1) : nctas / 2;
    sp.trace = tr; sp.trace_cta = trace_cta; sp.trace_items = TR_ITEMS;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nctas);
  cfg.blockDim = dim3(SG_EW * 32 + 128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs resident at once: they wait for each other's tiles
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  spin_gate_begin(s);
  cudaError_t le;
  if (mt == 1) le = cudaLaunchKernelEx(&cfg, stage_umma_kernel<1>, sp);
  else if (mt == 2) le = cudaLaunchKernelEx(&cfg, stage_umma_kernel<2>, sp);
  else le = cudaLaunchKernelEx(&cfg, stage_umma_kernel<4>, sp);
  spin_gate_end(s);
  B2_CUDA(le);
  B2_CUDA(cudaGetLastError());
  if (trace) {
    B2_CUDA(cudaDeviceSynchronize());
    unsigned long long t0 = ~0ull, t1 = 0;
    for (int i = 0; i < 8 * TR_ITEMS; ++i)
      if (tr[i]) { t0 = std::min(t0, tr[i]); t1 = std::max(t1, tr[i]); }
    const int c0 = off_t[trace_cta], c1 = off_t[trace_cta + 1];
    fprintf(stderr, "[stage] NT %d nch %d mt %d: %d CTAs, CTA %d has %d segment(s):", NT, nch, mt, nctas, trace_cta, c1 - c0);
    for (int q = c0; q < c1; ++q) fprintf(stderr, " (col %d K %d tiles %d-%d)", seg_t[q].col, convs[seg_t[q].col].K, seg_t[q].t_begin, seg_t[q].t_end);
    fprintf(stderr, "; first stamp to last %.1f us\n", (t1 - t0) * 1e-3);
    for (int i = 0; i < TR_ITEMS; ++i) {
      const unsigned long long* e = tr + (size_t)i * 8;
      if (!e[5]) break;
      fprintf(stderr, "[stage]  item %2d: dep-wait %6.2f..%6.2f | mma start %6.2f A0 landed %6.2f W0 landed %6.2f issued %6.2f | epi %6.2f..%6.2f\n", i,
              (e[0] - t0) * 1e-3, (e[1] - t0) * 1e-3, (e[3] - t0) * 1e-3, (e[7] - t0) * 1e-3, (e[2] - t0) * 1e-3, (e[4] - t0) * 1e-3, (e[5] - t0) * 1e-3,
              (e[6] - t0) * 1e-3);
    }
  }
  return 1;
}

// ------------------------------------------------------------------------------------------------
// pair mode (stage_pair_kernel): items are whole ResBlock steps; c1[lv][col], c2[lv][col] with lv = dilation step
// ------------------------------------------------------------------------------------------------
static size_t pair_smem(int NT, int C, int n) { return RbPairCfg<4>::smem_bytes(NT, C) + 8 * (size_t)n * C; }

bool stage_pair_supported(const ConvParams* c1, const ConvParams* c2, int n_levels, int n_cols) {
  const int n = n_levels * n_cols;
  if (n < 1 || 2 * n > SG_MAX_CONVS) return false;
  const int C = c1[0].Cin;
  if (C != 32) return false;  // the 512-row tile variant; the 64-channel stage gains nothing (its main loops dominate)
  for (int i = 0; i < n; ++i) {
    const ConvParams &a = c1[i], &b = c2[i];
    if (a.Cin != C || a.Cout != C || b.Cin != C || b.Cout != C || a.um_nt != C || b.um_nt != C) return false;
    if (a.um_nsplit != 2 || b.um_nsplit != 2 || !a.wu || !b.wu || !a.xp || !b.rp || !a.bias || !b.bias) return false;
    if (b.dil != 1 || b.K - 1 > RP_TPAD || (a.K - 1) * a.dil > UM_HALO - 16 || a.pad_left + b.pad_left > UM_PADL) return false;
    if (a.epi != EPI_PLAIN || b.epi != EPI_PLAIN || a.pix > 1 || b.pix > 1 || b.act != ACT_NONE || b.res) return false;
    if (!a.lens || a.lens != c1[0].lens || b.lens != a.lens || a.len_mul != c1[0].len_mul || b.len_mul != a.len_mul) return false;
    if (!(b.yp || b.y) || (b.yp && (b.yp == a.xp || b.yp_nsplit != 2))) return false;
    if (b.K != c2[i % n_cols].K || a.K != b.K) return false;  // one tile stride per branch
  }
  return pair_smem(C, C, n) <= 227 * 1024;
}

int launch_stage_pair(const ConvParams* c1, const ConvParams* c2, int n_levels, int n_cols, int batch, const int* lens_h, void* tab_h,
                      void* tab_d, unsigned int* flags_d, size_t flags_cap, cudaStream_t s) {
  B2_CHECK(stage_pair_supported(c1, c2, n_levels, n_cols), "stage_pair: unsupported stage geometry");
  constexpr int mt = 4;
  const int n = n_levels * n_cols;
  const int C = c1[0].Cin, nch = (C + 31) / 32;
  // cost of one tile of a branch through all steps (10 ns): main loops of c1 and c2 + the T hand-over in between, or the
  // output epilogue, whichever is longer (profiles/r2d_stage_trace.log: 0.5 us per tap and chunk, ~5 us per stored tile)
  std::vector<long long> wcol(n_cols, 0);
  std::vector<int> S(n_cols);
  for (int j = 0; j < n_cols; ++j) {
    S[j] = mt * 128 - (c2[j].K - 1);
    for (int lv = 0; lv < n_levels; ++lv) wcol[j] += std::max((long long)SG_COST_AP * c2[lv * n_cols + j].K * nch + 500, 550LL);
  }
  int tiles_max = 0;
  long long total = 0, n_tiles_all = 0;
  std::vector<int> mtl((size_t)batch * n_cols);
  for (int b = 0; b < batch; ++b)
    for (int j = 0; j < n_cols; ++j) {
      const int Lb = lens_h[b] * c2[0].len_mul;
      const int t = (Lb + S[j] - 1) / S[j];
      mtl[(size_t)b * n_cols + j] = t;
      tiles_max = std::max(tiles_max, t);
      total += (long long)t * wcol[j];
      n_tiles_all += t;
      B2_CHECK(c1[j].xp_rows >= UM_PADL + t * S[j] + mt * 128 + UM_HALO, "stage_pair: input plane image too short");
    }
  if (n_tiles_all == 0) return 0;
  B2_CHECK((size_t)n * batch * tiles_max <= flags_cap, "stage_pair: flag buffer too small");
  const size_t smem = pair_smem(C, C, n);
  int dev = 0, sms = 0, per_sm = 0;
  B2_CUDA(cudaGetDevice(&dev));
  B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stage_pair_kernel<4>, SG_EW * 32 + 128, smem));
  B2_CHECK(per_sm >= 1, "stage_pair: kernel does not fit an SM");
  const int nctas = (int)std::min<long long>(std::min(sms * per_sm, 1000), n_tiles_all);
  char* th = reinterpret_cast<char*>(tab_h);
  ConvParams* convs_t = reinterpret_cast<ConvParams*>(th);  // [0, n): c2 of each step, [n, 2n): c1
  int* dep_t = reinterpret_cast<int*>(th + sizeof(ConvParams) * SG_MAX_CONVS);
  int* off_t = dep_t + SG_MAX_CONVS;
  StageSeg* seg_t = reinterpret_cast<StageSeg*>(off_t + 1024);
  memcpy(convs_t, c2, sizeof(ConvParams) * n);
  memcpy(convs_t + n, c1, sizeof(ConvParams) * n);
  for (int lv = 0; lv < n_levels; ++lv)
    for (int j = 0; j < n_cols; ++j) dep_t[lv * n_cols + j] = lv > 0 ? (lv - 1) * n_cols + j : -1;
  std::vector<StageStrip> strips;
  if (g_striped && n_tiles_all + nctas <= SG_MAX_SEGS) {  // (the branches' tile strides differ by a few rows: same index ~ same time)
    for (int b = 0; b < batch; ++b)
      for (int t = 0; t < tiles_max; ++t)
        for (int j = 0; j < n_cols; ++j)
          if (t < mtl[(size_t)b * n_cols + j]) strips.push_back({j, b, 0, 1, wcol[j], t});
  } else {
    for (int j = 0; j < n_cols; ++j)
      for (int b = 0; b < batch; ++b)
        if (mtl[(size_t)b * n_cols + j] > 0) strips.push_back({j, b, 0, mtl[(size_t)b * n_cols + j], wcol[j]});
  }
  const int nseg = stage_partition(strips.data(), (int)strips.size(), nctas, off_t, seg_t, SG_MAX_SEGS);
  char* td = reinterpret_cast<char*>(tab_d);
  // only the bytes in use: conv table + dep + offsets + the segments actually written (~10 of 52 KB)
  B2_CUDA(cudaMemcpyAsync(td, th, (size_t)(reinterpret_cast<char*>(seg_t + nseg) - th), cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaMemsetAsync(flags_d, 0, sizeof(unsigned int) * (size_t)n * batch * tiles_max, s));
  StageParams sp;
  sp.convs = reinterpret_cast<const ConvParams*>(td);
  sp.convs1 = sp.convs + n;
  sp.dep = reinterpret_cast<const int*>(td + sizeof(ConvParams) * SG_MAX_CONVS);
  sp.seg_off = sp.dep + SG_MAX_CONVS;
  sp.segs = reinterpret_cast<const StageSeg*>(sp.seg_off + 1024);
  sp.flags = flags_d;
  sp.n_levels = n_levels; sp.n_cols = n_cols; sp.zb = batch; sp.tiles_max = tiles_max; sp.ntiles = 1;
  sp.trace = nullptr; sp.trace_cta = 0; sp.trace_items = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nctas);
  cfg.blockDim = dim3(SG_EW * 32 + 128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs resident at once: they wait for each other's tiles
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  spin_gate_begin(s);
  const cudaError_t le = cudaLaunchKernelEx(&cfg, stage_pair_kernel<4>, sp);
  spin_gate_end(s);
  B2_CUDA(le);
  B2_CUDA(cudaGetLastError());
  return 1;
}

// tests: cta_of_tile[k] = CTA of the k-th tile in strip order (strip i = column i, `strip_tiles[i]` tiles of `strip_cost[i]`)
int stage_partition_debug(const int* strip_tiles, const long long* strip_cost, int n_strips, int nctas, int* cta_of_tile) {
  B2_CHECK(strip_tiles && strip_cost && cta_of_tile && n_strips >= 1 && nctas >= 1 && nctas < 4096, "bad arguments");
  std::vector<StageStrip> strips;
  std::vector<long long> first(n_strips + 1, 0);
  for (int i = 0; i < n_strips; ++i) {
    B2_CHECK(strip_tiles[i] >= 0 && strip_cost[i] > 0, "bad strip");
    strips.push_back({i, 0, 0, strip_tiles[i], strip_cost[i]});
    first[i + 1] = first[i] + strip_tiles[i];
  }
  std::vector<int> off(nctas + 1, 0);
  std::vector<StageSeg> segs((size_t)first[n_strips] + 1);
  const int nseg = stage_partition(strips.data(), n_strips, nctas, off.data(), segs.data(), (int)segs.size());
  for (long long k = 0; k < first[n_strips]; ++k) cta_of_tile[k] = -1;
  for (int c = 0; c < nctas; ++c)
    for (int q = off[c]; q < off[c + 1]; ++q)
      for (int t = segs[q].t_begin; t < segs[q].t_end; ++t) {
        B2_CHECK(cta_of_tile[first[segs[q].col] + t] == -1, "tile assigned twice");
        cta_of_tile[first[segs[q].col] + t] = c;
      }
  return nseg;
}

}  // namespace b2
