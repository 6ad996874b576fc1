// C ABI of audiopure_b200 (see include/audiopure_b200.h).  Host-side launch logic only: tensor-map
// construction, workspace carving, the per-step launch sequence, the purifier loops and the NCCL shim.
#include "../../include/audiopure_b200.h"

#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "diffwave_kernels.cuh"
#include "mel_kernel.cuh"

namespace {

thread_local std::string g_err;

int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}

#define AP_CUDA(expr)                                                                                  \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      return fail(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" +        \
                  std::to_string(__LINE__) + ")");                                                     \
  } while (0)

#define AP_CHECK(cond, msg)         \
  do {                              \
    if (!(cond)) return fail(msg);  \
  } while (0)

// ------------------------------------------------------------------ tensor maps (driver entry point) --
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 (or, for the tf32 build, fp32) tensor, innermost dim contiguous, box = {128 bytes of elements, box1, 1, ...},
// 128-byte swizzle, out-of-bounds elements read as zero.
int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, uint32_t box1, bool f32 = false) {
  EncodeTiledFn fn = encode_tiled_fn();
  AP_CHECK(fn, "cuTensorMapEncodeTiled entry point not available from the driver");
  cuuint64_t gdim[5];
  cuuint64_t gstride[4];
  cuuint32_t box[5], estr[5];
  uint64_t stride = f32 ? 4 : 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    stride *= dims[i];
    if (i < rank - 1) gstride[i] = stride;
    box[i] = (i == 0) ? (f32 ? 32 : 64) : (i == 1 ? box1 : 1);
    estr[i] = 1;
  }
  CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                  gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
  return 0;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// SM count of the current device (the grid-stride kernels are sized in multiples of it), cached per device.
int current_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// Blocks for a grid-stride kernel over n items: enough to cover them, at most ctas_per_sm per SM.
unsigned grid_for(long long n, int threads, int ctas_per_sm) {
  long long blocks = (n + threads - 1) / threads;
  const long long cap = static_cast<long long>(current_sm_count()) * ctas_per_sm;
  if (blocks > cap) blocks = cap;
  return static_cast<unsigned>(blocks < 1 ? 1 : blocks);
}

}  // namespace

struct ap_net {
  int layers = 0, cycle = 0, T = 0, max_chunk = 64;
  bool tf32 = false;          // AP_FLAG_TF32: fp32 storage + kind::tf32 MMAs instead of bf16
  uint32_t round_bias = 0;    // tf32: added to fp32 operand bits so that the tensor core's narrowing rounds to nearest
  std::vector<float> h_b1, h_c2;  // host copies of the bias tables (passed to the kernels by value)
  ap::TailBias tail_bias{};
  ap_weights w{};
  std::vector<float> alpha, alpha_bar, sigma, sde_beta, sde_acp;
  CUtensorMap tm_w1, tm_w2, tm_ws, tm_wf;
  int num_sms = 0, device = 0;
  // activation maps for one (workspace, Bc, L); two sets are cached so that a batch that is not a multiple of
  // max_chunk (a full chunk and a remainder chunk alternate on every evaluation) re-encodes nothing
  struct ActMaps {
    const void* ws = nullptr;
    int B = 0, L = 0;
    uint64_t last_use = 0;
    CUtensorMap tm_h[2], tm_h_st[2], tm_gate, tm_gate_st;  // loads use [128-row] boxes, epilogue stores [32-row]
  };
  ActMaps maps[2];
  uint64_t map_clock = 0;
  // measurement hook (ap_profile_*)
  bool profile = false;
  struct Span {
    cudaEvent_t a, b;
    int kind;
  };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t take_event() {
    cudaEvent_t e = nullptr;
    if (!pool.empty()) {
      e = pool.back();
      pool.pop_back();
    } else {
      cudaEventCreate(&e);
    }
    return e;
  }
  ~ap_net() {
    for (auto& s : spans) {
      cudaEventDestroy(s.a);
      cudaEventDestroy(s.b);
    }
    for (auto e : pool) cudaEventDestroy(e);
  }
};

namespace {
// Brackets one kernel launch with events when profiling is on.
struct ProfSpan {
  ap_net* n;
  cudaStream_t st;
  cudaEvent_t b = nullptr;
  int kind;
  ProfSpan(ap_net* n_, cudaStream_t st_, int kind_) : n(n_), st(st_), kind(kind_) {
    if (!n->profile || n->spans.size() > (1u << 20)) return;
    cudaEvent_t a = n->take_event();
    b = n->take_event();
    cudaEventRecord(a, st);
    n->spans.push_back({a, b, kind});
  }
  ~ProfSpan() {
    if (b) cudaEventRecord(b, st);
  }
};
}  // namespace

namespace {

struct WsLayout {
  size_t h_bytes, off_h[2], off_gate, off_x[2], total;
};

WsLayout ws_layout(const ap_net* n, int Bc, int L) {
  WsLayout w;
  w.h_bytes = static_cast<size_t>(Bc) * L * ap::kC * (n->tf32 ? 4 : 2);  // a multiple of 512: enough for TMA (16 B) and 128 B lines
  w.off_h[0] = 0;
  w.off_h[1] = w.h_bytes;
  w.off_gate = 2 * w.h_bytes;
  const size_t xb = align_up(static_cast<size_t>(Bc) * L * 4, 1024);
  w.off_x[0] = w.off_gate + static_cast<size_t>(n->layers) * w.h_bytes;
  w.off_x[1] = w.off_x[0] + xb;
  w.total = w.off_x[1] + xb;
  return w;
}

int ensure_maps(ap_net* n, uint8_t* ws, int Bc, int L, const ap_net::ActMaps** out) {
  ap_net::ActMaps* victim = &n->maps[0];
  for (auto& m : n->maps) {
    if (m.ws == ws && m.B == Bc && m.L == L) {
      m.last_use = ++n->map_clock;
      *out = &m;
      return 0;
    }
    if (m.last_use < victim->last_use) victim = &m;
  }
  ap_net::ActMaps& m = *victim;
  m.ws = nullptr;  // invalid until every map below is encoded
  const WsLayout w = ws_layout(n, Bc, L);
  const uint64_t d3[3] = {static_cast<uint64_t>(ap::kC), static_cast<uint64_t>(L), static_cast<uint64_t>(Bc)};
  for (int i = 0; i < 2; ++i)
    if (make_map(&m.tm_h[i], ws + w.off_h[i], 3, d3, ap::kTileT, n->tf32) ||
        make_map(&m.tm_h_st[i], ws + w.off_h[i], 3, d3, 32, n->tf32))
      return 1;
  const uint64_t d4[4] = {static_cast<uint64_t>(ap::kC), static_cast<uint64_t>(L), static_cast<uint64_t>(Bc),
                          static_cast<uint64_t>(n->layers)};
  if (make_map(&m.tm_gate, ws + w.off_gate, 4, d4, ap::kTileT, n->tf32) ||
      make_map(&m.tm_gate_st, ws + w.off_gate, 4, d4, 32, n->tf32))
    return 1;
  m.ws = ws;
  m.B = Bc;
  m.L = L;
  m.last_use = ++n->map_clock;
  *out = &m;
  return 0;
}

// One epsilon-network evaluation of Bc clips (+ the fused output update described by `tail`).
int run_eval(ap_net* n, const float* x, int Bc, int L, int t, ap::TailArgs tail, uint8_t* ws, cudaStream_t st) {
  AP_CHECK(t >= 0 && t < n->T, "diffusion step t out of range [0, T)");
  const ap_net::ActMaps* mp = nullptr;
  if (ensure_maps(n, ws, Bc, L, &mp)) return 1;
  const WsLayout w = ws_layout(n, Bc, L);
  void* h0 = ws + w.off_h[0];
  const long long rows = static_cast<long long>(Bc) * L;
  {
    long long blocks = (rows + 7) / 8;
    const long long cap = static_cast<long long>(n->num_sms) * 16;
    if (blocks > cap) blocks = cap;
    ProfSpan span(n, st, 2);
    const float* part0 = n->w.part0 + static_cast<size_t>(t) * ap::kC;
    if (n->tf32)
      ap::prologue_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, n->w.w0, n->w.b0, part0, h0, rows,
                                                                                n->round_bias);
    else
      ap::prologue_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, n->w.w0, n->w.b0, part0, h0, rows, 0u);
  }
  const int tiles_per_clip = (L + ap::kTileT - 1) / ap::kTileT;
  const int num_tiles = tiles_per_clip * Bc;
  // persistent grid of CTA pairs (clusters of 2, one per TPC); each pair walks pairs of tiles
  const int units = (num_tiles + 1) / 2, clusters = n->num_sms / 2;
  const int grid = 2 * (units < clusters ? units : clusters);
  // Ablation switches of profiles/r01_ablation.md change results; they exist only in builds made with
  // -DAP_ENABLE_ABLATION (python -m audiopure_b200.build --ablation), never in the default library.
#ifdef AP_ENABLE_ABLATION
  static const int dbg = getenv("AP_DEBUG") ? atoi(getenv("AP_DEBUG")) : 0;
#else
  const int dbg = 0;
#endif
  // Programmatic dependent launch: each tensor-core kernel may begin its set-up (barriers, TMEM, descriptor
  // prefetch) while its predecessor drains; griddepcontrol.wait in the kernel orders every global access.
  static const bool pdl = getenv("AP_NO_PDL") == nullptr;
  cudaLaunchConfig_t lc{};
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  lc.gridDim = dim3(grid);
  lc.blockDim = dim3(ap::kThreads);
  lc.stream = st;
  lc.attrs = &attr;
  lc.numAttrs = pdl ? 1 : 0;
  for (int l = 0; l < n->layers; ++l) {
    ap::LayerArgs a;
    a.L = L;
    a.tiles_per_clip = tiles_per_clip;
    a.num_tiles = num_tiles;
    a.dilation = 1 << (l % n->cycle);
    a.layer = l;
    a.write_h = (l + 1 < n->layers) ? 1 : 0;
    a.debug = dbg;
    a.round_bias = n->round_bias;
    ap::LayerBias bias;
    memcpy(bias.b1, n->h_b1.data() + static_cast<size_t>(l) * 512, sizeof(bias.b1));
    memcpy(bias.c2, n->h_c2.data() + (static_cast<size_t>(t) * n->layers + l) * ap::kC, sizeof(bias.c2));
    ProfSpan span(n, st, 0);
    if (n->tf32) {
      lc.dynamicSmemBytes = ap::Mode<true>::kLayerSmem;
      AP_CUDA(cudaLaunchKernelEx(&lc, ap::layer_kernel<true>, mp->tm_h[l & 1], n->tm_w1, n->tm_w2, mp->tm_gate_st,
                                 mp->tm_h_st[(l + 1) & 1], bias, a));
    } else {
      lc.dynamicSmemBytes = ap::Mode<false>::kLayerSmem;
      AP_CUDA(cudaLaunchKernelEx(&lc, ap::layer_kernel<false>, mp->tm_h[l & 1], n->tm_w1, n->tm_w2, mp->tm_gate_st,
                                 mp->tm_h_st[(l + 1) & 1], bias, a));
    }
  }
  tail.bo = n->w.bo;
  tail.round_bias = n->round_bias;
  tail.B = Bc;
  tail.L = L;
  tail.tiles_per_clip = tiles_per_clip;
  tail.num_tiles = num_tiles;
  tail.num_layers = n->layers;
  {
    ProfSpan span(n, st, 1);
    if (n->tf32) {
      lc.dynamicSmemBytes = ap::Mode<true>::kTailSmem;
      AP_CUDA(cudaLaunchKernelEx(&lc, ap::tail_kernel<true>, mp->tm_gate, n->tm_ws, n->tm_wf, n->tail_bias, tail));
    } else {
      lc.dynamicSmemBytes = ap::Mode<false>::kTailSmem;
      AP_CUDA(cudaLaunchKernelEx(&lc, ap::tail_kernel<false>, mp->tm_gate, n->tm_ws, n->tm_wf, n->tail_bias, tail));
    }
  }
  AP_CUDA(cudaGetLastError());
  return 0;
}

ap::TailArgs tail_update(const float* x_in, float* x_out, float ca, float cb, float cc, const float* z,
                         uint64_t seed, uint32_t stream_id, long long elem_offset) {
  ap::TailArgs t{};
  t.x_in = x_in;
  t.x_out = x_out;
  t.eps_out = nullptr;
  t.z = z;
  t.ca = ca;
  t.cb = cb;
  t.cc = cc;
  t.seed = seed;
  t.stream_lo = stream_id;
  t.elem_offset = elem_offset;
  return t;
}

int launch_axpbz(const float* x, const float* z, float* y, float ca, float cb, long long n, uint64_t seed,
                 uint32_t stream_id, long long elem_offset, int num_sms, cudaStream_t st) {
  long long blocks = (n + 255) / 256;
  const long long cap = static_cast<long long>(num_sms) * 8;
  if (blocks > cap) blocks = cap;
  ap::axpbz_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, z, y, ca, cb, n, seed, stream_id, elem_offset);
  AP_CUDA(cudaGetLastError());
  return 0;
}

int check_call(const ap_net* n, int B, int L, const void* ws, size_t ws_bytes) {
  AP_CHECK(n, "null handle");
  AP_CHECK(B > 0 && L > 0, "B and L must be positive");
  AP_CHECK(ws, "null workspace");
  AP_CHECK(reinterpret_cast<uintptr_t>(ws) % 1024 == 0, "workspace must be 1024-byte aligned");
  AP_CHECK(ws_bytes >= ap_workspace_bytes(n, B, L), "workspace too small: see ap_workspace_bytes");
  return 0;
}

// Purifier loop shared by DDPM and SDE: diffuse, then `steps` fused eval+update launches per chunk.
struct StepCoef {
  float ca, cb, cc;
  int t;
};

int purify_loop(ap_net* n, const float* x_in, float* x_out, int B, int L, float diff_a, float diff_b,
                const std::vector<StepCoef>& steps, const float* z, uint64_t seed, int64_t clip_offset, uint8_t* ws,
                cudaStream_t st) {
  const size_t BL = static_cast<size_t>(B) * L;
  for (int c0 = 0; c0 < B; c0 += n->max_chunk) {
    const int Bc = (B - c0) < n->max_chunk ? (B - c0) : n->max_chunk;
    const WsLayout w = ws_layout(n, Bc, L);
    float* xb[2] = {reinterpret_cast<float*>(ws + w.off_x[0]), reinterpret_cast<float*>(ws + w.off_x[1])};
    const size_t off = static_cast<size_t>(c0) * L;
    const long long elem_off = (clip_offset + c0) * static_cast<long long>(L);
    if (launch_axpbz(x_in + off, z ? z + off : nullptr, xb[0], diff_a, diff_b, static_cast<long long>(Bc) * L, seed,
                     0u, elem_off, n->num_sms, st))
      return 1;
    int cur = 0;
    for (size_t i = 0; i < steps.size(); ++i) {
      const bool last = (i + 1 == steps.size());
      float* dst = last ? x_out + off : xb[cur ^ 1];
      const float* zi = (z && steps[i].cc != 0.f) ? z + (i + 1) * BL + off : nullptr;
      ap::TailArgs tail = tail_update(xb[cur], dst, steps[i].ca, steps[i].cb, steps[i].cc, zi, seed,
                                      static_cast<uint32_t>(i + 1), elem_off);
      if (run_eval(n, xb[cur], Bc, L, steps[i].t, tail, ws, st)) return 1;
      cur ^= 1;
    }
  }
  return 0;
}

int run_debug_gemm(bool tf32, const void* a, const void* b, float* d, int K, cudaStream_t st) {
  const int subk = tf32 ? 32 : 64;
  AP_CHECK(K > 0 && K % subk == 0, "K must be a positive multiple of one 128-byte operand row");
  CUtensorMap ta, tb;
  const uint64_t da[2] = {static_cast<uint64_t>(K), 128}, db[2] = {static_cast<uint64_t>(K), 256};
  if (make_map(&ta, a, 2, da, 128, tf32) || make_map(&tb, b, 2, db, 256, tf32)) return 1;
  const int smem = ap::kABytes + 256 * 128 + 64 + 1024;
  if (tf32) {
    AP_CUDA(cudaFuncSetAttribute(ap::debug_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ap::debug_gemm_kernel<true><<<1, 128, smem, st>>>(ta, tb, d, K);
  } else {
    AP_CUDA(cudaFuncSetAttribute(ap::debug_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ap::debug_gemm_kernel<false><<<1, 128, smem, st>>>(ta, tb, d, K);
  }
  AP_CUDA(cudaGetLastError());
  return 0;
}

// How does tcgen05 kind::tf32 narrow an fp32 operand word?  1 + 0.75 ulp_tf32 times 1 comes back as 1 if the low 13
// mantissa bits are dropped (then the kernels add half a tf32 ulp to every operand they write, which makes the
// drop a round-to-nearest), as 1 + ulp if the tensor core rounds by itself (then they add nothing).
int probe_tf32_round_bias(uint32_t* bias) {
  float *a = nullptr, *b = nullptr, *d = nullptr;
  const size_t na = 128 * 32, nb = 256 * 32, nd = 128 * 256;
  AP_CUDA(cudaMalloc(&a, (na + nb + nd) * sizeof(float)));
  b = a + na;
  d = b + nb;
  const float va = 1.0f + 0.75f / 1024.0f, vb = 1.0f;
  float got = -1.f;
  cudaError_t e = cudaMemset(a, 0, (na + nb + nd) * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(a, &va, sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(b, &vb, sizeof(float), cudaMemcpyHostToDevice);
  int rc = 0;
  if (e == cudaSuccess) rc = run_debug_gemm(true, a, b, d, 32, nullptr);
  if (e == cudaSuccess && rc == 0) e = cudaMemcpy(&got, d, sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(a);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(std::string("tf32 probe: ") + cudaGetErrorString(e));
  if (got == 1.0f)
    *bias = 0x1000u;
  else if (got == 1.0f + 1.0f / 1024.0f)
    *bias = 0u;
  else
    return fail("tf32 probe: unexpected product " + std::to_string(got));
  return 0;
}

}  // namespace

extern "C" {

const char* ap_last_error(void) { return g_err.c_str(); }
int ap_abi_version(void) { return AP_ABI_VERSION; }

int ap_create(const ap_config* cfg, const ap_weights* w, ap_net** out) {
  AP_CHECK(cfg && w && out, "null argument");
  AP_CHECK(cfg->num_res_layers > 0 && cfg->num_res_layers <= 256, "num_res_layers out of range");
  AP_CHECK(cfg->dilation_cycle > 0 && cfg->dilation_cycle <= 16, "dilation_cycle out of range");
  AP_CHECK(cfg->T > 0, "T must be positive");
  AP_CHECK(cfg->alpha && cfg->alpha_bar && cfg->sigma && cfg->sde_beta && cfg->sde_alphas_cumprod,
           "schedule tables missing");
  int dev = 0;
  AP_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  AP_CUDA(cudaGetDeviceProperties(&prop, dev));
  AP_CHECK(prop.major == 10, std::string("audiopure_b200 needs an sm_100 device, found sm_") +
                                 std::to_string(prop.major) + std::to_string(prop.minor));
  ap_net* n = new ap_net();
  n->layers = cfg->num_res_layers;
  n->cycle = cfg->dilation_cycle;
  n->T = cfg->T;
  if ((cfg->flags & ~static_cast<uint32_t>(AP_FLAG_TF32)) != 0) {
    delete n;
    return fail("unknown ap_config.flags");
  }
  n->tf32 = (cfg->flags & AP_FLAG_TF32) != 0;
  n->max_chunk = cfg->max_chunk > 0 ? cfg->max_chunk : (n->tf32 ? 32 : 64);
  if (n->tf32 && probe_tf32_round_bias(&n->round_bias)) {
    delete n;
    return 1;
  }
  n->w = *w;
  n->alpha.assign(cfg->alpha, cfg->alpha + cfg->T);
  n->alpha_bar.assign(cfg->alpha_bar, cfg->alpha_bar + cfg->T);
  n->sigma.assign(cfg->sigma, cfg->sigma + cfg->T);
  n->sde_beta.assign(cfg->sde_beta, cfg->sde_beta + cfg->T);
  n->sde_acp.assign(cfg->sde_alphas_cumprod, cfg->sde_alphas_cumprod + cfg->T);
  n->num_sms = prop.multiProcessorCount;
  n->device = dev;
  const uint64_t L = static_cast<uint64_t>(n->layers);
  const uint64_t dw1[2] = {768, L * 512}, dw2[2] = {256, L * 256}, dws[2] = {L * 256, 256}, dwf[2] = {256, 256};
  const uint32_t brows = ap::Tc::kBRows;  // weight rows staged per CTA per K step (half of N = 256)
  int rc = make_map(&n->tm_w1, w->w1, 2, dw1, brows, n->tf32) || make_map(&n->tm_w2, w->w2, 2, dw2, brows, n->tf32) ||
           make_map(&n->tm_ws, w->ws, 2, dws, brows, n->tf32) || make_map(&n->tm_wf, w->wf, 2, dwf, brows, n->tf32);
  if (rc) {
    delete n;
    return 1;
  }
  n->h_b1.resize(static_cast<size_t>(n->layers) * 512);
  n->h_c2.resize(static_cast<size_t>(n->T) * n->layers * ap::kC);
  cudaError_t es[] = {
      cudaMemcpy(n->h_b1.data(), w->b1, n->h_b1.size() * sizeof(float), cudaMemcpyDeviceToHost),
      cudaMemcpy(n->h_c2.data(), w->c2, n->h_c2.size() * sizeof(float), cudaMemcpyDeviceToHost),
      cudaMemcpy(n->tail_bias.bs, w->bs, sizeof(n->tail_bias.bs), cudaMemcpyDeviceToHost),
      cudaMemcpy(n->tail_bias.bf, w->bf, sizeof(n->tail_bias.bf), cudaMemcpyDeviceToHost),
      cudaMemcpy(n->tail_bias.wo, w->wo, sizeof(n->tail_bias.wo), cudaMemcpyDeviceToHost),
      cudaFuncSetAttribute(ap::layer_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ap::Mode<false>::kLayerSmem),
      cudaFuncSetAttribute(ap::tail_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ap::Mode<false>::kTailSmem),
      cudaFuncSetAttribute(ap::layer_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ap::Mode<true>::kLayerSmem),
      cudaFuncSetAttribute(ap::tail_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ap::Mode<true>::kTailSmem)};
  for (cudaError_t e : es)
    if (e != cudaSuccess) {
      delete n;
      return fail(std::string("ap_create: copying bias tables / raising the shared-memory limit: ") + cudaGetErrorString(e));
    }
  *out = n;
  return 0;
}

void ap_destroy(ap_net* net) { delete net; }

size_t ap_workspace_bytes(const ap_net* net, int B, int L) {
  if (!net || B <= 0 || L <= 0) return 0;
  const int Bc = B < net->max_chunk ? B : net->max_chunk;
  return ws_layout(net, Bc, L).total;
}

int ap_eps(ap_net* net, const float* x, int B, int L, int t, float* eps_out, void* ws, size_t ws_bytes,
           void* stream) {
  if (check_call(net, B, L, ws, ws_bytes)) return 1;
  AP_CHECK(x && eps_out, "null tensor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int c0 = 0; c0 < B; c0 += net->max_chunk) {
    const int Bc = (B - c0) < net->max_chunk ? (B - c0) : net->max_chunk;
    ap::TailArgs tail{};
    tail.x_in = x + static_cast<size_t>(c0) * L;
    tail.eps_out = eps_out + static_cast<size_t>(c0) * L;
    if (run_eval(net, x + static_cast<size_t>(c0) * L, Bc, L, t, tail, static_cast<uint8_t*>(ws), st)) return 1;
  }
  return 0;
}

int ap_step(ap_net* net, const float* x_in, float* x_out, int B, int L, int t, float ca, float cb, float cc,
            const float* z, uint64_t seed, uint32_t stream_id, int64_t clip_offset, void* ws, size_t ws_bytes,
            void* stream) {
  if (check_call(net, B, L, ws, ws_bytes)) return 1;
  AP_CHECK(x_in && x_out, "null tensor");
  AP_CHECK(x_in != x_out, "ap_step is not in-place");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int c0 = 0; c0 < B; c0 += net->max_chunk) {
    const int Bc = (B - c0) < net->max_chunk ? (B - c0) : net->max_chunk;
    const size_t off = static_cast<size_t>(c0) * L;
    ap::TailArgs tail = tail_update(x_in + off, x_out + off, ca, cb, cc, z ? z + off : nullptr, seed, stream_id,
                                    (clip_offset + c0) * static_cast<long long>(L));
    if (run_eval(net, x_in + off, Bc, L, t, tail, static_cast<uint8_t*>(ws), st)) return 1;
  }
  return 0;
}

int ap_ddpm_purify(ap_net* net, const float* x_in, float* x_out, int B, int L, int t_star, const float* z,
                   uint64_t seed, int64_t clip_offset, void* ws, size_t ws_bytes, void* stream) {
  if (check_call(net, B, L, ws, ws_bytes)) return 1;
  AP_CHECK(x_in && x_out, "null tensor");
  AP_CHECK(t_star >= 1 && t_star <= net->T, "t_star out of range [1, T]");
  // diffwave_ddpm.py:67 and :159-160, in double from the reference-built fp32 tables
  const double ab = net->alpha_bar[t_star - 1];
  std::vector<StepCoef> steps;
  for (int t = t_star - 1; t >= 0; --t) {
    const double al = net->alpha[t], abt = net->alpha_bar[t];
    StepCoef s;
    s.t = t;
    s.ca = static_cast<float>(1.0 / std::sqrt(al));
    s.cb = static_cast<float>(-(1.0 - al) / std::sqrt(1.0 - abt) / std::sqrt(al));
    s.cc = t > 0 ? net->sigma[t] : 0.f;
    steps.push_back(s);
  }
  return purify_loop(net, x_in, x_out, B, L, static_cast<float>(std::sqrt(ab)), static_cast<float>(std::sqrt(1.0 - ab)),
                     steps, z, seed, clip_offset, static_cast<uint8_t*>(ws), static_cast<cudaStream_t>(stream));
}

int ap_sde_purify(ap_net* net, const float* x_in, float* x_out, int B, int L, int t, const float* z, uint64_t seed,
                  int64_t clip_offset, void* ws, size_t ws_bytes, void* stream) {
  if (check_call(net, B, L, ws, ws_bytes)) return 1;
  AP_CHECK(x_in && x_out, "null tensor");
  AP_CHECK(t >= 1 && t <= net->T, "t out of range [1, T]");
  // diffwave_sde.py:190-191 (diffusion) and :73-134 with dt = 1/N (one Euler-Maruyama step at index k):
  //   x <- (1 + beta_k/2) x - beta_k / sqrt(1 - abar_k) eps + sqrt(beta_k) sqrt((1-abar_{k-1})/(1-abar_k)) z
  const double ab = net->sde_acp[t - 1];
  std::vector<StepCoef> steps;
  for (int k = t - 1; k >= 0; --k) {
    const double bk = net->sde_beta[k], ak = net->sde_acp[k];
    StepCoef s;
    s.t = k;
    s.ca = static_cast<float>(1.0 + 0.5 * bk);
    s.cb = static_cast<float>(-bk / std::sqrt(1.0 - ak));
    s.cc = k > 0 ? static_cast<float>(std::sqrt(bk) * std::sqrt(1.0 - net->sde_acp[k - 1]) / std::sqrt(1.0 - ak)) : 0.f;
    steps.push_back(s);
  }
  return purify_loop(net, x_in, x_out, B, L, static_cast<float>(std::sqrt(ab)), static_cast<float>(std::sqrt(1.0 - ab)),
                     steps, z, seed, clip_offset, static_cast<uint8_t*>(ws), static_cast<cudaStream_t>(stream));
}

int ap_one_shot(ap_net* net, const float* x_in, float* x_out, int B, int L, int reverse_timestep, void* ws,
                size_t ws_bytes, void* stream) {
  AP_CHECK(net, "null handle");
  AP_CHECK(reverse_timestep >= 1 && reverse_timestep <= net->T, "reverse_timestep out of range [1, T]");
  const int t = reverse_timestep - 1;
  const double ab = net->alpha_bar[t];  // diffwave_ddpm.py:195-205
  return ap_step(net, x_in, x_out, B, L, t, static_cast<float>(std::sqrt(1.0 / ab)),
                 static_cast<float>(-std::sqrt(1.0 / ab - 1.0)), 0.f, nullptr, 0, 0, 0, ws, ws_bytes, stream);
}

int ap_logmel(const float* x, int B, int L, float* out, const ap_mel_tables* tabs, void* stream) {
  AP_CHECK(x && out && tabs, "null argument");
  AP_CHECK(B > 0 && L > 0, "B and L must be positive");
  AP_CHECK(tabs->n_mels > 0 && tabs->n_mels <= 128, "n_mels out of range");
  AP_CHECK(tabs->fb_nnz > 0 && tabs->fb_nnz <= ap::kMaxFbNnz, "fb_nnz out of range");
  ap::MelArgs a;
  a.fb_nnz = tabs->fb_nnz;
  a.x = x;
  a.out = out;
  a.tw = static_cast<const float2*>(tabs->twiddles);
  a.fb_start = tabs->fb_start;
  a.fb_len = tabs->fb_len;
  a.fb_off = tabs->fb_off;
  a.fb_w = tabs->fb_w;
  a.B = B;
  a.L = L;
  a.n_frames = 1 + L / ap::kHop;
  a.n_mels = tabs->n_mels;
  const int pairs = (a.n_frames + 1) / 2;
  // persistent over (clip, frame pair) items: up to 7 CTAs of 128 threads per SM (register-resident FFT), so the
  // 1024 items of a 64-clip batch are all resident at once on 148 SMs
  ap::logmel_kernel<<<grid_for(static_cast<long long>(B) * pairs, 1, 7), ap::kFwdThreads, 0,
                      static_cast<cudaStream_t>(stream)>>>(a);
  AP_CUDA(cudaGetLastError());
  return 0;
}

int ap_logmel_backward(const float* x, int B, int L, const float* grad_out, float* grad_x, const ap_mel_tables* tabs,
                       void* stream) {
  AP_CHECK(x && grad_out && grad_x && tabs, "null argument");
  AP_CHECK(B > 0 && L > 0, "B and L must be positive");
  AP_CHECK(L <= 40000, "ap_logmel_backward supports clips of at most 40000 samples");
  AP_CHECK(tabs->n_mels > 0 && tabs->n_mels <= 128, "n_mels out of range");
  AP_CHECK(tabs->fb_nnz > 0 && tabs->fb_nnz <= ap::kMaxFbNnz, "fb_nnz out of range");
  ap::MelBwdArgs a;
  a.x = x;
  a.grad_out = grad_out;
  a.grad_x = grad_x;
  a.tw = static_cast<const float2*>(tabs->twiddles);
  a.fb_start = tabs->fb_start;
  a.fb_len = tabs->fb_len;
  a.fb_off = tabs->fb_off;
  a.fb_w = tabs->fb_w;
  a.B = B;
  a.L = L;
  a.n_frames = 1 + L / ap::kHop;
  a.n_mels = tabs->n_mels;
  const size_t smem = static_cast<size_t>(L) * sizeof(float);
  AP_CUDA(cudaFuncSetAttribute(ap::logmel_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem)));
  ap::logmel_backward_kernel<<<B, ap::kMelThreads, smem, static_cast<cudaStream_t>(stream)>>>(a);
  AP_CUDA(cudaGetLastError());
  return 0;
}

int ap_smooth_inputs_batch(const float* x, int L, int n_rows, int64_t flat0, int64_t per_clip, int64_t first_draw,
                           float sigma, float scale, const float* z, uint64_t seed, uint32_t clip_key0, float* out,
                           void* stream) {
  AP_CHECK(x && out, "null tensor");
  AP_CHECK(L > 0 && n_rows > 0, "L and n_rows must be positive");
  AP_CHECK(flat0 >= 0 && per_clip > 0 && first_draw >= 0, "flat0 / first_draw must be >= 0 and per_clip positive");
  ap::SmoothArgs a;
  a.x = x;
  a.zinj = z;
  a.out = out;
  a.L = L;
  a.n_rows = n_rows;
  a.flat0 = flat0;
  a.per_clip = per_clip;
  a.first_draw = first_draw;
  a.sigma = sigma;
  a.scale = scale;
  a.seed = seed;
  a.clip_key0 = clip_key0;
  const bool vec = (L % 4 == 0) && reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(z) % 16 == 0;
  if (!vec && L % 4 == 0) return fail("ap_smooth_inputs: tensors must be 16-byte aligned when L is a multiple of 4");
  const long long n = static_cast<long long>(n_rows) * (vec ? L / 4 : L);
  ap::smooth_inputs_kernel<<<grid_for(n, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  AP_CUDA(cudaGetLastError());
  return 0;
}

int ap_smooth_inputs(const float* x, int L, int n_draws, float sigma, float scale, const float* z, uint64_t seed,
                     uint32_t clip, int64_t first_draw, float* out, void* stream) {
  // one clip: the work list is just its draws (per_clip larger than any draw count keeps clip == 0)
  return ap_smooth_inputs_batch(x, L, n_draws, 0, int64_t(1) << 62, first_draw, sigma, scale, z, seed, clip, out, stream);
}

int ap_vote_counts_batch(const float* logits, int rows, int K, int64_t flat0, int64_t per_clip, int64_t n_split,
                         int n_clips, int64_t* counts, void* stream) {
  AP_CHECK(logits && counts, "null tensor");
  AP_CHECK(rows > 0 && K > 0 && n_clips > 0, "rows, K and n_clips must be positive");
  AP_CHECK(flat0 >= 0 && per_clip > 0, "flat0 must be >= 0 and per_clip positive");
  AP_CHECK((flat0 + rows - 1) / per_clip < n_clips, "rows reach past the last clip's counters");
  ap::vote_counts_kernel<<<grid_for(rows, 256, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, rows, K, flat0, per_clip, n_split, n_clips, reinterpret_cast<unsigned long long*>(counts));
  AP_CUDA(cudaGetLastError());
  return 0;
}

int ap_vote_counts(const float* logits, int rows, int K, int64_t* counts, void* stream) {
  return ap_vote_counts_batch(logits, rows, K, 0, int64_t(1) << 62, int64_t(1) << 62, 1, counts, stream);
}

namespace {
int nes_args(ap::NesArgs* a, const float* x, int audios, int L, int S, int lead, float sigma, const float* z,
             uint64_t seed, uint32_t audio_key0, int64_t draw0) {
  AP_CHECK(x, "null tensor");
  AP_CHECK(audios > 0 && L > 0, "audios and L must be positive");
  AP_CHECK(S > 0 && S % 2 == 0, "samples per draw batch must be positive and even (antithetic pairs, _NES.py:19-21)");
  AP_CHECK(lead == 0 || lead == 1, "lead must be 0 or 1");
  AP_CHECK(draw0 >= 0, "draw0 must be >= 0");
  *a = ap::NesArgs{};
  a->x = x;
  a->zinj = z;
  a->L = L;
  a->audios = audios;
  a->S = S;
  a->lead = lead;
  a->sigma = sigma;
  a->seed = seed;
  a->audio_key0 = audio_key0;
  a->draw0 = draw0;
  return 0;
}
}  // namespace

int ap_nes_inputs(const float* x, int audios, int L, int S, int lead, float sigma, const float* z, uint64_t seed,
                  uint32_t audio_key0, int64_t draw0, float* out, void* stream) {
  ap::NesArgs a;
  if (nes_args(&a, x, audios, L, S, lead, sigma, z, seed, audio_key0, draw0)) return 1;
  AP_CHECK(out, "null tensor");
  a.out = out;
  const long long n = static_cast<long long>(audios) * (S / 2) * L;
  ap::nes_inputs_kernel<<<grid_for(n, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  AP_CUDA(cudaGetLastError());
  return 0;
}

int ap_nes_grad(const float* loss, int loss_stride, int loss_off, int audios, int L, int S, float grad_scale,
                const float* z, uint64_t seed, uint32_t audio_key0, int64_t draw0, float* grad, void* stream) {
  ap::NesArgs a;
  if (nes_args(&a, loss, audios, L, S, 0, 0.f, z, seed, audio_key0, draw0)) return 1;
  AP_CHECK(grad, "null tensor");
  AP_CHECK(loss_off >= 0 && loss_off + S <= loss_stride, "loss row too short for S samples at loss_off");
  a.loss = loss;
  a.loss_stride = loss_stride;
  a.loss_off = loss_off;
  a.grad = grad;
  a.grad_scale = grad_scale;
  const long long n = static_cast<long long>(audios) * L;
  ap::nes_grad_kernel<<<grid_for(n, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  AP_CUDA(cudaGetLastError());
  return 0;
}

int ap_bias_act_nhwc_bf16(void* y, const float* bias, const void* residual, int64_t rows, int C, int relu,
                          void* stream) {
  AP_CHECK(y && bias, "null tensor");
  AP_CHECK(rows > 0 && C > 0 && C % 8 == 0, "rows must be positive and C a positive multiple of 8");
  AP_CHECK(reinterpret_cast<uintptr_t>(y) % 16 == 0 && reinterpret_cast<uintptr_t>(residual) % 16 == 0 &&
               reinterpret_cast<uintptr_t>(bias) % 16 == 0,
           "tensors must be 16-byte aligned");
  const long long n_vec = static_cast<long long>(rows) * (C / 8);
  ap::bias_act_kernel<<<grid_for(n_vec, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint4*>(y), bias, static_cast<const uint4*>(residual), n_vec, C / 8, relu);
  AP_CUDA(cudaGetLastError());
  return 0;
}

int ap_profile_enable(ap_net* net, int enable) {
  AP_CHECK(net, "null handle");
  net->profile = enable != 0;
  return 0;
}

int ap_profile_read(ap_net* net, double ms_sum[3], int64_t launches[3]) {
  AP_CHECK(net && ms_sum && launches, "null argument");
  for (auto& s : net->spans) {
    AP_CUDA(cudaEventSynchronize(s.b));
    float ms = 0.f;
    AP_CUDA(cudaEventElapsedTime(&ms, s.a, s.b));
    ms_sum[s.kind] += ms;
    launches[s.kind] += 1;
    net->pool.push_back(s.a);
    net->pool.push_back(s.b);
  }
  net->spans.clear();
  return 0;
}

int ap_debug_gemm(const void* a_bf16, const void* b_bf16, float* d, int K, void* stream) {
  AP_CHECK(a_bf16 && b_bf16 && d, "null tensor");
  return run_debug_gemm(false, a_bf16, b_bf16, d, K, static_cast<cudaStream_t>(stream));
}

int ap_debug_gemm_tf32(const float* a_f32, const float* b_f32, float* d, int K, void* stream) {
  AP_CHECK(a_f32 && b_f32 && d, "null tensor");
  return run_debug_gemm(true, a_f32, b_f32, d, K, static_cast<cudaStream_t>(stream));
}

int ap_precision(const ap_net* net, int* tf32, uint32_t* round_bias) {
  AP_CHECK(net && tf32 && round_bias, "null argument");
  *tf32 = net->tf32 ? 1 : 0;
  *round_bias = net->round_bias;
  return 0;
}

// ------------------------------------------------------------------------------------ NCCL (dlopen) --
struct ap_comm {
  void* comm = nullptr;
};

namespace {
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, ...) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
struct NcclId {
  char internal[128];
};
typedef int (*CommInitRankFn)(void**, int, NcclId, int);

NcclApi* nccl() {
  static NcclApi api;
  if (api.lib) return &api;
  const char* env = getenv("AP_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm) continue;
    api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return nullptr;
  api.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<int (*)(void**, int, ...)>(dlsym(api.lib, "ncclCommInitRank"));
  api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(
      dlsym(api.lib, "ncclAllReduce"));
  api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(api.lib, "ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) {
    dlclose(api.lib);
    api.lib = nullptr;
    return nullptr;
  }
  return &api;
}
int nccl_fail(NcclApi* n, const char* what, int rc) {
  return fail(std::string(what) + ": " + (n->GetErrorString ? n->GetErrorString(rc) : "NCCL error") + " (" +
              std::to_string(rc) + ")");
}
}  // namespace

int ap_comm_unique_id(uint8_t id[AP_COMM_ID_BYTES]) {
  NcclApi* n = nccl();
  AP_CHECK(n, "libnccl.so.2 could not be loaded (set AP_NCCL_LIB)");
  NcclId nid;
  int rc = n->GetUniqueId(&nid);
  if (rc) return nccl_fail(n, "ncclGetUniqueId", rc);
  memcpy(id, nid.internal, AP_COMM_ID_BYTES);
  return 0;
}

int ap_comm_init(int rank, int world, const uint8_t id[AP_COMM_ID_BYTES], ap_comm** out) {
  NcclApi* n = nccl();
  AP_CHECK(n, "libnccl.so.2 could not be loaded (set AP_NCCL_LIB)");
  AP_CHECK(out && world > 0 && rank >= 0 && rank < world, "bad rank/world");
  NcclId nid;
  memcpy(nid.internal, id, AP_COMM_ID_BYTES);
  ap_comm* c = new ap_comm();
  int rc = reinterpret_cast<CommInitRankFn>(n->CommInitRank)(&c->comm, world, nid, rank);
  if (rc) {
    delete c;
    return nccl_fail(n, "ncclCommInitRank", rc);
  }
  *out = c;
  return 0;
}

int ap_allreduce_counts(ap_comm* comm, int64_t* counts, size_t cnt, void* stream) {
  NcclApi* n = nccl();
  AP_CHECK(n && comm && comm->comm, "communicator not initialised");
  const int kNcclInt64 = 4, kNcclSum = 0;
  int rc = n->AllReduce(counts, counts, cnt, kNcclInt64, kNcclSum, comm->comm, static_cast<cudaStream_t>(stream));
  if (rc) return nccl_fail(n, "ncclAllReduce", rc);
  return 0;
}

void ap_comm_destroy(ap_comm* comm) {
  if (!comm) return;
  NcclApi* n = nccl();
  if (n && comm->comm) n->CommDestroy(comm->comm);
  delete comm;
}

}  // extern "C"
