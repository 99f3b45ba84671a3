// C ABI of libreni_b200.so -- see include/reni_b200.h.  Host-side glue only: argument
// validation, workspace carving and kernel launches on the caller's stream.
#include "../../include/reni_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include "allreduce_kernel.cuh"
#include "bwd_kernel.cuh"
#include "dw_kernel.cuh"
#include "fwd_kernel.cuh"
#include "layout.cuh"
#include "phase.cuh"
#include "lbwd_kernel.cuh"
#include "shade_kernel.cuh"
#include "small_kernels.cuh"
#include "film_map_fused.cuh"

#ifndef RENI_NO_FORK
#define RENI_NO_FORK 0  // 1: keep every kernel of the step on the caller's stream (A/B switch for the fork/join)
#endif
#ifndef RENI_FWD_TRAIN_ALLHANDS
#define RENI_FWD_TRAIN_ALLHANDS 1  // paired training forward: all-hands epilogue (0: grouped)
#endif
#ifndef RENI_FILM_DW_PERSISTENT
#define RENI_FILM_DW_PERSISTENT 1  // FiLM weight-gradient GEMM: persistent CTAs over one list of all (map, job) blocks
#endif
#ifndef RENI_BWD_TRAIN_PAIR
#define RENI_BWD_TRAIN_PAIR 1  // CTA pairs also for the delta chain with weight gradients (0: one CTA per tile pair)
#endif
#ifndef RENI_DW_BALANCE
#define RENI_DW_BALANCE 0  // weight-gradient GEMM: 1 = CTAs per job in proportion to the job's stash bytes.  Measured
                           // slower (dW 328 vs 256 us at cfg 2): a CTA's throughput is set by its three stages in
                           // flight, not by its bytes, so the output-layer job needs as many CTAs as a hidden job
#endif
#ifndef RENI_LBWD
#define RENI_LBWD 0  // default training backward of the Cond-by-Concat decoder: 0 = tile-major chain + split-K weight-gradient
                     // GEMM, 1 = layer-major (lbwd_kernel.cuh: delta chain and weight gradients in one pass per layer).
                     // Per call RENI_FLAG_LAYER_MAJOR_BWD / RENI_FLAG_TILE_MAJOR_BWD override it.  Measured at cfg 2:
                     // layer-major 448 + 68 us against 293 + 248 us, but its map-level backward cannot hide under a
                     // weight-gradient GEMM (+50 us exposed): no gain for the step, hence not the default.
#endif
#ifndef RENI_FWD_SPLIT_W
#define RENI_FWD_SPLIT_W 1  // forward hidden layers on two-term fp16 weights (W' = W_hi + W_lo, two MMAs per K step): removes
                            // the weight-rounding half of the fp16 operand error (radiance rel-L2 x 0.6..0.8) at twice the
                            // forward tensor-core work; RENI_FLAG_FWD_SINGLE_TERM / RENI_FLAG_FWD_TWO_TERM override per call
#endif
#ifndef RENI_FWD_PAIR
#define RENI_FWD_PAIR 1  // forward on CTA pairs (0: one CTA per tile pair, grouped training / all-hands inference epilogue)
#endif

using namespace reni;

namespace {

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

bool config_ok(const reni_config_t* c) {
  if (c == nullptr) return false;
  if (c->hidden_features != kH) return false;
  if (c->hidden_layers < 1 || c->hidden_layers > kMaxHiddenLayers) return false;
  if (c->out_features < 1 || c->out_features > 3) return false;
  if (c->equivariance < 0 || c->equivariance > 2) return false;
  if (c->ndims < 1 || c->ndims > 128) return false;
  if (c->output_activation != 0 && c->output_activation != 1) return false;
  return true;
}

int64_t tiles_per_map(int64_t P) { return (P + kTileRows - 1) / kTileRows; }

// side length W of the equirectangular grid with P = W * W / 2 directions (get_directions(W), utils.py:46-65); 0 if none
int grid_sidelen(int64_t P) {
  int64_t w = 2;
  while (w * w / 2 < P) w += 2;
  return (w * w / 2 == P) ? (int)w : 0;
}

WorkspaceLayout make_layout(const reni_config_t* c, int64_t B, int64_t P, int32_t flags) {
  WorkspaceLayout w{};
  const int64_t L = c->hidden_layers;
  const int64_t ntiles = B * tiles_per_map(P);
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    int64_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  w.wf = take(L * kWImageBytes);
  w.wb = take(L * kWImageBytes);
  w.wf2 = take(L * kWImageBytes);
  w.wf2lo = take(L * kWImageBytes);
  w.wb2 = take(L * kWImageBytes);
  w.wbias2 = take(L * 2 * kBiasBlockBytes);
  w.w6f = take(kW6ImageBytes);
  w.w6b = take(kW6ImageBytes);
  w.bias = take((L * kH + 16) * 4);
  w.scalars = take(256 * 4);
  w.mc = take(B * 5 * kH * 4);
  w.dmc = take(B * 5 * kH * 4);
  w.map_loss = take(B * 32 * 4);
  w.loss_part = take(ntiles * 4 * kLossPartials * 4);
  const int64_t nin = reni_in_features(c);
  w.xc = take(B * 5 * nin * 4);   // xfull: input columns paired with [dM0..dM3, dc]
  w.dxc = take(B * 5 * nin * 4);  // E: gradient w.r.t. xfull
  w.dip = -1;
  w.wf2m = w.wf2m_lo = w.wb2m = w.wbias2m = -1;
  if ((flags & RENI_FLAG_FILM) && (flags & RENI_FLAG_FILM_PERMAP)) {  // (placed ahead of everything the other flags size)
    w.wf2m = take(B * L * (int64_t)kWImageBytes);
    w.wf2m_lo = take(B * L * (int64_t)kWImageBytes);
    w.wb2m = take(B * L * (int64_t)kWImageBytes);
    w.wbias2m = take(B * L * 2 * (int64_t)kBiasBlockBytes);
  }
  if (flags & RENI_FLAG_SAVE_FOR_BACKWARD) {
    const bool dw = (flags & (RENI_FLAG_NEED_DW | RENI_FLAG_FILM)) != 0;  // (FiLM: dfreq / dphase need the delta stash)
    w.stash_c = take(ntiles * (L + 1) * (int64_t)kPhaseTileBytes);  // phase stash (rebuilds both h and cos), phase.cuh
    w.stash_h = -1;
    w.stash_d = dw ? take(ntiles * (L + 1) * (int64_t)kTileImageBytes) : -1;  // slot 0 unused (delta_0 stays on chip)
    w.stash_gy = take(ntiles * (int64_t)kGyImageBytes);
    w.aout = c->last_layer_linear ? -1 : take(B * P * 3 * 4);  // sine output layer: its pre-activations
    w.ready = dw ? take((ntiles + 16) * 4) : -1;
  } else {
    w.stash_c = w.stash_h = w.stash_d = w.stash_gy = w.aout = w.ready = -1;
  }
  w.film_S = w.film_cs = -1;
  if ((flags & RENI_FLAG_FILM) && (flags & RENI_FLAG_SAVE_FOR_BACKWARD)) {
    w.film_S = take(B * L * (int64_t)kH * kH * 4);  // per-map delta_l^T h_{l-1}
    w.film_cs = take(B * L * (int64_t)kH * 4);      // per-map column sums of delta_l
  }
  w.total = off;
  return w;
}

// RENI_FLAG_FILM_PERMAP needs every unit of four tiles inside one map
bool permap_ok(int64_t P) { return P % kTileRows == 0 && (P / kTileRows) % 4 == 0; }

// RENI_FLAG_GRID_DIRECTIONS / RENI_FLAG_GRID_SINEWEIGHT: directions and sine weights in closed form from the pixel index
// (utils.py:46-78); the sw argument then carries the mask as one bit per pixel (or null)
struct GridArgs {
  int w = 0, dir = 0, sw = 0;
  const uint32_t* mask = nullptr;
};
bool grid_args(int32_t flags, int64_t P, const float* D, const float* sw, GridArgs* g) {
  g->dir = (flags & RENI_FLAG_GRID_DIRECTIONS) ? 1 : 0;
  g->sw = (flags & RENI_FLAG_GRID_SINEWEIGHT) ? 1 : 0;
  if (g->dir || g->sw) {
    g->w = grid_sidelen(P);
    if (g->w == 0) return false;  // P is not W * W / 2
  }
  if (!g->dir && D == nullptr) return false;
  g->mask = g->sw ? reinterpret_cast<const uint32_t*>(sw) : nullptr;
  return true;
}

template <class T>
T* at(void* ws, int64_t off) {
  return off < 0 ? nullptr : reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(ws) + off);
}

// Optional phase timing (debug hook, reni_debug_set_phase_events): CUDA events recorded between the kernels of a
// step so a caller can time each kernel on the launching stream.  Process-wide (autograd runs reni_backward /
// reni_film_backward on its own thread); empty by default; set and cleared by the measuring thread while no call runs.
cudaEvent_t g_phase_events[16];
volatile int g_num_phase_events = 0;
thread_local unsigned long long* g_trace = nullptr;  // reni_debug_set_trace
// reni_debug_set_overlap: SMs given to the co-resident weight-gradient kernel (-1: 36 % share, 0: overlap off) and
// how many of them take the output-layer job (0: round robin over the jobs)
volatile int g_overlap_dw_ctas = 0;  // (measured slower than back-to-back kernels at cfg 2: off unless asked for)
volatile int g_overlap_out_ctas = 0;
inline void mark_phase(int i, cudaStream_t s) {
  if (i < g_num_phase_events && g_phase_events[i] != nullptr) cudaEventRecord(g_phase_events[i], s);
}

// last CUDA runtime error seen by this thread inside the library (debug hook reni_debug_last_cuda_error)
thread_local cudaError_t g_last_cuda = cudaSuccess;
inline cudaError_t note(cudaError_t e) {
  if (e != cudaSuccess) g_last_cuda = e;
  return e;
}
inline cudaError_t last_err() { return note(cudaGetLastError()); }

// Side stream for the map-level backward (tiny latency-bound kernels that only depend on the delta chain): forked
// after the delta-chain kernel, joined at the end of the step, so it runs under the weight-gradient GEMM.  One per
// host thread and device, created on first use; event-based fork/join is legal inside CUDA-graph capture.
struct SideStream {
  int dev = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  cudaEvent_t step[kFilmMapMaxLinears + 1] = {};  // per-layer forks of reni_film_map_backward, created on first use
};
thread_local SideStream g_side;
thread_local cudaEvent_t g_fwd_wait_event = nullptr;  // see reni_forward / RENI_FLAG_PREPARE_WEIGHTS
bool side_stream(SideStream** out) {
  int dev = 0;
  if (note(cudaGetDevice(&dev)) != cudaSuccess) return false;
  if (g_side.dev != dev) {
    SideStream s;
    if (note(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess) return false;
    if (note(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming)) != cudaSuccess) return false;
    if (note(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming)) != cudaSuccess) return false;
    s.dev = dev;
    g_side = s;  // (a thread that hops devices leaks one stream + two events per hop; callers are one-process-per-GPU)
  }
  *out = &g_side;
  return true;
}

// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda): a 2-D uint8 view of `bytes` of
// global memory as rows of 256 B, box = rows_per_box rows -> one TMA load moves rows_per_box * 256 contiguous bytes.
bool encode_rows256(CUtensorMap* map, const void* base, uint64_t bytes, uint32_t rows_per_box) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (note(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q)) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || ptr == nullptr)
      return false;
    fn = reinterpret_cast<EncodeFn>(ptr);
  }
  const cuuint64_t dims[2] = {256, bytes / 256};
  const cuuint64_t strides[1] = {256};
  const cuuint32_t box[2] = {256, rows_per_box};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int num_sms() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}

}  // namespace

extern "C" {

static int32_t launch_forward(const reni_config_t* c, const WorkspaceLayout& w, const float* mc, const float* film,
                              const float* D, int64_t d_bstride, int64_t B, int64_t P, float* out, const float* target,
                              const float* sw, int64_t sw_bstride, void* ws, int32_t flags, cudaStream_t stream, int sms);

int32_t reni_abi_version(void) { return RENI_ABI_VERSION; }

const char* reni_strerror(int32_t code) {
  switch (code) {
    case RENI_OK: return "ok";
    case RENI_ERR_BAD_CONFIG: return "unsupported decoder configuration (hidden_features must be 256, 1..6 hidden layers, out_features <= 3)";
    case RENI_ERR_BAD_ARGUMENT: return "bad argument (null pointer or non-positive size)";
    case RENI_ERR_WORKSPACE: return "workspace too small or not 1024-byte aligned";
    case RENI_ERR_CUDA: return "CUDA runtime error";
    case RENI_ERR_NO_DEVICE: return "no sm_100 device";
    default: return "unknown error";
  }
}

int64_t reni_in_features(const reni_config_t* c) {
  if (c == nullptr) return RENI_ERR_BAD_ARGUMENT;
  const int64_t N = c->ndims;
  switch (c->equivariance) {
    case RENI_EQ_SO2: return 2 * N + N * N + 2;
    case RENI_EQ_SO3: return N + N * N;
    case RENI_EQ_NONE: return 3 * N + N;
    default: return RENI_ERR_BAD_CONFIG;
  }
}

int64_t reni_workspace_bytes(const reni_config_t* c, int64_t B, int64_t P, int32_t flags) {
  if (!config_ok(c)) return RENI_ERR_BAD_CONFIG;
  if (B < 1 || P < 1) return RENI_ERR_BAD_ARGUMENT;
  return make_layout(c, B, P, flags).total;
}

static int32_t launch_prepare_weights(const reni_config_t* c, const float* const* host_weights,
                                      const float* const* host_biases, void* ws, int64_t ws_bytes, cudaStream_t stream);

int32_t reni_prepare_weights(const reni_config_t* c, const float* const* host_weights,
                             const float* const* host_biases, void* ws, int64_t ws_bytes, void* stream) {
  if (!config_ok(c)) return RENI_ERR_BAD_CONFIG;
  if (host_weights == nullptr || host_biases == nullptr || ws == nullptr) return RENI_ERR_BAD_ARGUMENT;
  return launch_prepare_weights(c, host_weights, host_biases, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

static int32_t launch_prepare_weights(const reni_config_t* c, const float* const* host_weights,
                                      const float* const* host_biases, void* ws, int64_t ws_bytes, cudaStream_t stream) {
  const WorkspaceLayout w = make_layout(c, 1, 1, 0);
  if (ws_bytes < w.scalars || (reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return RENI_ERR_WORKSPACE;
  PrepParams p{};
  const int L = c->hidden_layers;
  for (int i = 0; i <= L + 1; ++i) {
    if (host_weights[i] == nullptr || host_biases[i] == nullptr) return RENI_ERR_BAD_ARGUMENT;
    p.w[i] = host_weights[i];
    p.b[i] = host_biases[i];
  }
  p.wf = at<__half>(ws, w.wf);
  p.wb = at<__half>(ws, w.wb);
  p.wf2 = at<__half>(ws, w.wf2);
  p.wf2lo = at<__half>(ws, w.wf2lo);
  p.wb2 = at<__half>(ws, w.wb2);
  p.wbias2 = at<__half>(ws, w.wbias2);
  p.w6f = at<__half>(ws, w.w6f);
  p.w6b = at<__half>(ws, w.w6b);
  p.bias = at<float>(ws, w.bias);
  p.L = L;
  p.out_features = c->out_features;
  p.last_sine = c->last_layer_linear ? 0 : 1;
  p.first_omega = c->first_omega_0;
  p.hidden_omega = c->hidden_omega_0;
  reni_prep_weights_kernel<<<dim3(32, L + 1), 256, 0, stream>>>(p);
  return last_err() == cudaSuccess ? RENI_OK : RENI_ERR_CUDA;
}

int32_t reni_forward(const reni_config_t* c, const float* Z, const float* D, int64_t d_bstride,
                     const float* weight0, const float* bias0, int64_t B, int64_t P, float* out,
                     const float* target, const float* sw, int64_t sw_bstride, void* ws, int64_t ws_bytes,
                     int32_t flags, void* stream_) {
  if (!config_ok(c)) return RENI_ERR_BAD_CONFIG;
  if (Z == nullptr || (D == nullptr && !(flags & RENI_FLAG_GRID_DIRECTIONS)) || weight0 == nullptr || bias0 == nullptr ||
      out == nullptr || ws == nullptr || B < 1 || P < 1)
    return RENI_ERR_BAD_ARGUMENT;
  if ((flags & RENI_FLAG_LOSS) && (target == nullptr || (sw == nullptr && !(flags & RENI_FLAG_GRID_SINEWEIGHT))))
    return RENI_ERR_BAD_ARGUMENT;
  const WorkspaceLayout w = make_layout(c, B, P, flags);
  if (ws_bytes < w.total || (reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return RENI_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int sms = num_sms();
  if (sms <= 0) return RENI_ERR_NO_DEVICE;

  mark_phase(0, stream);
  // ---- per-map prologue: layer 0 hoisted to (M_b, c_b)
  {
    PrologueParams q{};
    q.Z = Z;
    q.W0 = weight0;
    q.b0 = bias0;
    q.mc = at<float>(ws, w.mc);
    q.xfull = at<float>(ws, w.xc);
    q.B = (int)B;
    q.N = c->ndims;
    q.in_features = (int)reni_in_features(c);
    q.equivariance = c->equivariance;
    q.omega0 = c->first_omega_0;
    q.scalars = (flags & RENI_FLAG_LOSS) ? at<float>(ws, w.scalars) : nullptr;
    q.fused_S = 1.5f * (float)P;  // gradient scale of the fused loss: S = 3P/2 (see reni_loss_finish_kernel)
    const size_t smem = (size_t)(q.in_features + 3 * q.N) * sizeof(float);
    if (smem > 48 * 1024)
      note(cudaFuncSetAttribute(reni_prologue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    reni_prologue_kernel<<<dim3(kH / 8, (unsigned)B), 256, smem, stream>>>(q);
    if (last_err() != cudaSuccess) return RENI_ERR_CUDA;
  }
  mark_phase(1, stream);
  if (g_fwd_wait_event != nullptr) {  // (set by reni_loss_forward_backward: weight images being built on the side stream)
    if (note(cudaStreamWaitEvent(stream, g_fwd_wait_event, 0)) != cudaSuccess) return RENI_ERR_CUDA;
  }
  return launch_forward(c, w, at<float>(ws, w.mc), nullptr, D, d_bstride, B, P, out, target, sw, sw_bstride, ws, flags,
                        stream, sms);
}

// Fused decoder forward over hoisted per-map layer-0 operands mc (B, 5, 256); film != null selects the FiLM epilogue.
static int32_t launch_forward(const reni_config_t* c, const WorkspaceLayout& w, const float* mc, const float* film,
                              const float* D, int64_t d_bstride, int64_t B, int64_t P, float* out, const float* target,
                              const float* sw, int64_t sw_bstride, void* ws, int32_t flags, cudaStream_t stream,
                              int sms) {
  FwdParams p{};
  p.D = D;
  p.d_bstride = d_bstride;
  p.mc = mc;
  p.film = film;
  p.wf = at<__half>(ws, w.wf);
  p.wf2 = at<__half>(ws, w.wf2);
  p.w6f = at<__half>(ws, w.w6f);
  p.bias = at<float>(ws, w.bias);
  p.out = out;
  p.stash_u = at<uint16_t>(ws, w.stash_c);
  p.target = (flags & RENI_FLAG_LOSS) ? target : nullptr;
  p.sw = sw;
  p.sw_bstride = sw_bstride;
  GridArgs ga;
  if (!grid_args(flags, P, D, sw, &ga)) return RENI_ERR_BAD_ARGUMENT;
  p.grid_w = ga.w;
  p.dir_grid = ga.dir;
  p.sw_grid = ga.sw;
  p.mask_bits = ga.mask;
  p.loss_part = (flags & RENI_FLAG_LOSS) ? at<float>(ws, w.loss_part) : nullptr;
  p.aout = at<float>(ws, w.aout);
  p.B = (int)B;
  p.P = (int)P;
  p.tiles_per_map = (int)tiles_per_map(P);
  p.ntiles = (int)(B * tiles_per_map(P));
  p.L = c->hidden_layers;
  p.out_tanh = c->output_activation == 1;
  p.last_sine = c->last_layer_linear ? 0 : 1;
  p.so2 = c->equivariance == RENI_EQ_SO2;
  p.trace = RENI_BWD_TRACE ? nullptr : g_trace;
  memset(&p.wmap, 0, sizeof(p.wmap));
  memset(&p.wmap_lo, 0, sizeof(p.wmap_lo));
  memset(&p.bmap, 0, sizeof(p.bmap));
  p.split = (flags & RENI_FLAG_FWD_SINGLE_TERM) ? 0 : ((flags & RENI_FLAG_FWD_TWO_TERM) ? 1 : RENI_FWD_SPLIT_W);
  const int npairs = (p.ntiles + 1) / 2;
  const bool train = (flags & RENI_FLAG_SAVE_FOR_BACKWARD) != 0;
  // CTA pairs (cluster of 2) share the weight stream: half the L2 -> SM weight traffic per SM
  const bool pair_mode = RENI_FWD_PAIR != 0;
  int grid = npairs < sms ? npairs : sms;
  // FiLM on per-map images: the modulation lives in each map's weight / bias images, the kernel is the plain one
  const bool permap = film != nullptr && (flags & RENI_FLAG_FILM_PERMAP) != 0;
  if (permap) {
    if (!pair_mode || w.wf2m < 0) return RENI_ERR_BAD_CONFIG;
    p.film = nullptr;
    p.w_map_rows = p.L * (kWImageBytes / 256);
    p.b_map_rows = p.L * 2 * (kBiasBlockBytes / 256);
    if (!encode_rows256(&p.wmap, at<__half>(ws, w.wf2m), (uint64_t)B * p.L * kWImageBytes, kWChunkBytes / 256))
      return RENI_ERR_CUDA;
    if (!encode_rows256(&p.wmap_lo, at<__half>(ws, w.wf2m_lo), (uint64_t)B * p.L * kWImageBytes, kWChunkBytes / 256))
      return RENI_ERR_CUDA;
    if (!encode_rows256(&p.bmap, at<__half>(ws, w.wbias2m), (uint64_t)B * p.L * 2 * kBiasBlockBytes,
                        kBiasBlockBytes / 256))
      return RENI_ERR_CUDA;
    const int nquads = (p.ntiles + 3) / 4;
    const int nclusters = nquads < sms / 2 ? nquads : sms / 2;
    grid = 2 * nclusters;
  } else if (pair_mode) {  // one cluster of two CTAs per tile quad
    if (!encode_rows256(&p.wmap, p.wf2, (uint64_t)p.L * kWImageBytes, kWChunkBytes / 256)) return RENI_ERR_CUDA;
    if (!encode_rows256(&p.wmap_lo, at<__half>(ws, w.wf2lo), (uint64_t)p.L * kWImageBytes, kWChunkBytes / 256))
      return RENI_ERR_CUDA;
    if (!encode_rows256(&p.bmap, at<__half>(ws, w.wbias2), (uint64_t)p.L * 2 * kBiasBlockBytes, kBiasBlockBytes / 256))
      return RENI_ERR_CUDA;
    const int nquads = (p.ntiles + 3) / 4;
    const int nclusters = nquads < sms / 2 ? nquads : sms / 2;
    grid = 2 * nclusters;
  }
  cudaError_t e;
  auto launch = [&](auto kernel) {
    e = note(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::kTotal));
    if (e != cudaSuccess) return;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kFwdThreads);
    cfg.dynamicSmemBytes = FwdSmem::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pair_mode ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = note(cudaLaunchKernelEx(&cfg, kernel, p));
  };
  if (permap) {
    if (train) launch(reni_fwd_kernel<true, true, true>);
    else launch(reni_fwd_kernel<false, true, true>);
  } else if (film != nullptr) {
    if (!pair_mode) return RENI_ERR_BAD_CONFIG;  // (the unpaired build is an A/B fallback without the FiLM epilogue)
    if (train) launch(reni_fwd_kernel<true, true, true, true>);
    else launch(reni_fwd_kernel<false, true, true, true>);
  } else if (pair_mode) {
    if (train && RENI_FWD_TRAIN_ALLHANDS) launch(reni_fwd_kernel<true, true, true>);
    else if (train) launch(reni_fwd_kernel<true, false, true>);
    else launch(reni_fwd_kernel<false, true, true>);
  } else {
    if (train) launch(reni_fwd_kernel<true, false, false>);
    else launch(reni_fwd_kernel<false, true, false>);
  }
  if (e != cudaSuccess) return RENI_ERR_CUDA;
  mark_phase(2, stream);
  return last_err() == cudaSuccess ? RENI_OK : RENI_ERR_CUDA;
}

// FiLM extras of the shared backward tail (null: Cond-by-Concat decoder)
struct FilmBackwardArgs {
  const float* film;                     // (B, L, 2, 256)
  float* d_mc;                           // (B, 5, 256), written
  float* d_film;                         // (B, L, 2, 256), written
  const float* const* host_weights;      // fp32 parameters [1..L] are read (dfreq needs W_l and b_l)
  const float* const* host_biases;
};

// Shared tail of reni_backward / reni_loss_forward_backward: delta chain, weight-gradient GEMMs, layer-0 and
// map-level reductions.  scalars[0..1] (gradient scale) must already be on the stream.
static int32_t launch_backward(const reni_config_t* c, const WorkspaceLayout& w, const float* Z, const float* D,
                               int64_t d_bstride, const float* weight0, int64_t B, int64_t P, const float* out,
                               const float* grad_out, const float* target, const float* sw, int64_t sw_bstride,
                               float alpha, float* dZ, float* const* host_dW, float* const* host_db, void* ws,
                               int32_t flags, cudaStream_t stream, int use_cos = 1, SideStream* side = nullptr,
                               const FilmBackwardArgs* film = nullptr) {
  const int sms = num_sms();
  if (sms <= 0) return RENI_ERR_NO_DEVICE;
  const bool want_dw = (flags & RENI_FLAG_NEED_DW) != 0;   // the caller wants weight gradients
  const bool need_dw = want_dw || film != nullptr;         // the delta stash + weight-gradient GEMM run
  const int L = c->hidden_layers;
  const int ntiles = (int)(B * tiles_per_map(P));
  float* dmc = film != nullptr ? film->d_mc : at<float>(ws, w.dmc);
  if (cudaMemsetAsync(dmc, 0, (size_t)B * 5 * kH * 4, stream) != cudaSuccess) return RENI_ERR_CUDA;
  mark_phase(3, stream);

  GridArgs ga;
  if (!grid_args(flags, P, D, sw, &ga)) return RENI_ERR_BAD_ARGUMENT;
  // ---- layer-major backward (lbwd_kernel.cuh): head + one launch per hidden layer, delta chain and dW together
  const bool lbwd_wanted = (flags & RENI_FLAG_LAYER_MAJOR_BWD) != 0 || (RENI_LBWD && !(flags & RENI_FLAG_TILE_MAJOR_BWD));
  const bool use_lbwd = lbwd_wanted && want_dw && film == nullptr && g_overlap_dw_ctas == 0;
  if (use_lbwd) {
    for (int i = 1; i <= L + 1; ++i)
      if (host_dW[i] == nullptr || host_db[i] == nullptr) return RENI_ERR_BAD_ARGUMENT;
    LbwdHeadParams hp{};
    hp.out = out;
    hp.grad_out = grad_out;
    hp.aout = at<float>(ws, w.aout);
    hp.target = target;
    hp.sw = sw;
    hp.sw_bstride = sw_bstride;
    hp.map_loss = at<float>(ws, w.map_loss);
    hp.scalars = at<float>(ws, w.scalars);
    hp.w6b = at<__half>(ws, w.w6b);
    hp.stash_u = at<uint16_t>(ws, w.stash_c);
    hp.stash_d = at<__half>(ws, w.stash_d);
    hp.dW_out = host_dW[L + 1];
    hp.db_out = host_db[L + 1];
    hp.out_scale = c->last_layer_linear ? 1.f : c->hidden_omega_0;
    hp.P = (int)P;
    hp.tiles_per_map = (int)tiles_per_map(P);
    hp.ntiles = ntiles;
    hp.L = L;
    hp.out_tanh = c->output_activation == 1;
    hp.use_cos = use_cos;
    hp.out_features = c->out_features;
    hp.grid_w = ga.w;
    hp.sw_grid = ga.sw;
    hp.mask_bits = ga.mask;
    if (note(cudaFuncSetAttribute(reni_lbwd_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  LbwdHeadSmem::kTotal)) != cudaSuccess)
      return RENI_ERR_CUDA;
    reni_lbwd_head_kernel<<<ntiles < sms ? ntiles : sms, kLbwdHeadThreads, LbwdHeadSmem::kTotal, stream>>>(hp);
    if (last_err() != cudaSuccess) return RENI_ERR_CUDA;
    mark_phase(4, stream);
    LbwdParams lp{};
    lp.stash_u = hp.stash_u;
    lp.stash_d = hp.stash_d;
    lp.wb2 = at<__half>(ws, w.wb2);
    lp.scalars = hp.scalars;
    lp.D = D;
    lp.d_bstride = d_bstride;
    lp.dmc = dmc;
    lp.P = (int)P;
    lp.tiles_per_map = hp.tiles_per_map;
    lp.ntiles = ntiles;
    lp.L = L;
    lp.so2 = c->equivariance == RENI_EQ_SO2;
    lp.grid_w = ga.w;
    lp.dir_grid = ga.dir;
    lp.trace = nullptr;
    const int npair = ntiles < sms / 2 ? ntiles : sms / 2;
    if (note(cudaFuncSetAttribute(reni_lbwd_layer_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  LbwdSmem::kTotal)) != cudaSuccess ||
        note(cudaFuncSetAttribute(reni_lbwd_layer_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  LbwdSmem::kTotal)) != cudaSuccess)
      return RENI_ERR_CUDA;
    for (int l = L; l >= 1; --l) {
      lp.l = l;
      lp.dW = host_dW[l];
      lp.db = host_db[l];
      lp.trace = (l == L - 1 || (L == 1 && l == 1)) ? g_trace : nullptr;  // (debug timeline of one mid-chain launch)
      lp.rev = (L - l + 1) & 1;  // each launch walks the tiles against the previous one: the newest deltas are in L2
      if (l == 1) reni_lbwd_layer_kernel<true><<<2 * npair, kLbwdThreads, LbwdSmem::kTotal, stream>>>(lp);
      else reni_lbwd_layer_kernel<false><<<2 * npair, kLbwdThreads, LbwdSmem::kTotal, stream>>>(lp);
    }
    if (last_err() != cudaSuccess) return RENI_ERR_CUDA;
    mark_phase(5, stream);
  }

  BwdParams p{};
  p.trace = RENI_BWD_TRACE ? g_trace : nullptr;  // (debug builds: the timeline hook follows the delta chain instead)
  p.out = out;
  p.grad_out = grad_out;
  p.aout = at<float>(ws, w.aout);
  p.target = target;
  p.sw = sw;
  p.sw_bstride = sw_bstride;
  p.map_loss = at<float>(ws, w.map_loss);
  p.scalars = at<float>(ws, w.scalars);
  p.wb = at<__half>(ws, w.wb);
  p.wb2 = at<__half>(ws, w.wb2);
  p.w6b = at<__half>(ws, w.w6b);
  p.stash_u = at<uint16_t>(ws, w.stash_c);
  p.stash_d = at<__half>(ws, w.stash_d);
  p.stash_gy = at<__half>(ws, w.stash_gy);
  p.D = D;
  p.d_bstride = d_bstride;
  p.grid_w = ga.w;
  p.dir_grid = ga.dir;
  p.sw_grid = ga.sw;
  p.mask_bits = ga.mask;
  p.dmc = dmc;
  const bool permap = film != nullptr && (flags & RENI_FLAG_FILM_PERMAP) != 0;  // modulation folded into per-map images
  p.film = (film != nullptr && !permap) ? film->film : nullptr;
  p.so2 = c->equivariance == RENI_EQ_SO2;
  p.B = (int)B;
  p.P = (int)P;
  p.tiles_per_map = (int)tiles_per_map(P);
  p.ntiles = ntiles;
  p.L = L;
  p.out_tanh = c->output_activation == 1;
  p.d_slots = need_dw ? L + 1 : 1;
  p.use_cos = use_cos;
  // CTA pairs, one cluster of two CTAs per tile quad (tcgen05.mma.cta_group::2).  With weight gradients the pair only
  // pays off because the delta stash is written by bulk copies from shared memory (measured at cfg 2: unpaired
  // st.global 305 us, unpaired bulk 312, paired st.global 335, paired bulk 291).
  const bool pair_mode = !need_dw || RENI_BWD_TRAIN_PAIR || film != nullptr;
  // Overlap mode: the weight-gradient kernel runs BESIDE the delta chain on dw_ctas of the SMs (side stream, forked
  // before the chain) and consumes each tile's stash blocks from L2 as the chain finishes them (per-tile counters),
  // instead of re-reading 1.5 GB from HBM in a kernel of its own afterwards.  Only when there are several waves of
  // tile quads per cluster to pipeline over; off under the per-kernel timing hook (which wants kernels back to back).
  int dw_ctas = 0;
  if (want_dw && film == nullptr && pair_mode && RENI_BWD_BULK_STASH && g_num_phase_events == 0 && !RENI_NO_FORK &&
      w.ready >= 0 && (ntiles + 3) / 4 >= 3 * (sms / 2)) {
    dw_ctas = g_overlap_dw_ctas < 0 ? ((int)(sms * 0.36) & ~1) : (g_overlap_dw_ctas & ~1);
    if (dw_ctas < 2 * (L + 1) || dw_ctas > sms - 4) dw_ctas = 0;
  }
  if (dw_ctas > 0) {
    p.ready = at<uint32_t>(ws, w.ready);
    if (cudaMemsetAsync(p.ready, 0, ((size_t)ntiles + 16) * 4, stream) != cudaSuccess) return RENI_ERR_CUDA;
    if (side == nullptr && !side_stream(&side)) return RENI_ERR_CUDA;
    if (note(cudaEventRecord(side->fork, stream)) != cudaSuccess) return RENI_ERR_CUDA;
    if (note(cudaStreamWaitEvent(side->stream, side->fork, 0)) != cudaSuccess) return RENI_ERR_CUDA;
  }
  memset(&p.wmap, 0, sizeof(p.wmap));
  if (!use_lbwd) {
    cudaLaunchConfig_t cfg{};
    if (pair_mode) {
      const int nquads = (ntiles + 3) / 4;
      const int avail = (sms - dw_ctas) / 2;
      const int nclusters = nquads < avail ? nquads : avail;
      cfg.gridDim = dim3((unsigned)(2 * nclusters));
      if (permap) {
        if (w.wb2m < 0) return RENI_ERR_BAD_CONFIG;
        p.w_map_rows = L * (kWImageBytes / 256);
        if (!encode_rows256(&p.wmap, at<__half>(ws, w.wb2m), (uint64_t)B * L * kWImageBytes, kWChunkBytes / 256))
          return RENI_ERR_CUDA;
      } else if (!encode_rows256(&p.wmap, p.wb2, (uint64_t)L * kWImageBytes, kWChunkBytes / 256)) {
        return RENI_ERR_CUDA;
      }
    } else {
      const int npairs = (ntiles + 1) / 2;
      cfg.gridDim = dim3((unsigned)(npairs < sms ? npairs : sms));
    }
    cfg.blockDim = dim3(kBwdThreads);
    cfg.dynamicSmemBytes = BwdSmem::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pair_mode ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (permap) {
      if (note(cudaFuncSetAttribute(reni_bwd_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    BwdSmem::kTotal)) != cudaSuccess)
        return RENI_ERR_CUDA;
      if (note(cudaLaunchKernelEx(&cfg, reni_bwd_kernel<true, true, false>, p)) != cudaSuccess) return RENI_ERR_CUDA;
    } else if (film != nullptr) {
      if (note(cudaFuncSetAttribute(reni_bwd_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    BwdSmem::kTotal)) != cudaSuccess)
        return RENI_ERR_CUDA;
      if (note(cudaLaunchKernelEx(&cfg, reni_bwd_kernel<true, true, true>, p)) != cudaSuccess) return RENI_ERR_CUDA;
    } else if (need_dw) {
      constexpr bool kTrainPair = RENI_BWD_TRAIN_PAIR != 0;
      if (note(cudaFuncSetAttribute(reni_bwd_kernel<true, kTrainPair>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    BwdSmem::kTotal)) != cudaSuccess)
        return RENI_ERR_CUDA;
      if (note(cudaLaunchKernelEx(&cfg, reni_bwd_kernel<true, kTrainPair>, p)) != cudaSuccess) return RENI_ERR_CUDA;
    } else {
      if (note(cudaFuncSetAttribute(reni_bwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    BwdSmem::kTotal)) != cudaSuccess)
        return RENI_ERR_CUDA;
      if (note(cudaLaunchKernelEx(&cfg, reni_bwd_kernel<false, true>, p)) != cudaSuccess) return RENI_ERR_CUDA;
    }
  }
  if (last_err() != cudaSuccess) return RENI_ERR_CUDA;
  if (!use_lbwd) mark_phase(4, stream);

  // fork: the map-level backward below only needs dmc from the delta chain; with the per-kernel timing hook active
  // everything stays on the caller's stream so the phase events keep their meaning
  // (overlap mode: the fork happened before the chain, the weight-gradient kernel takes the side stream -- launched
  // AFTER the chain so that a tool that serialises kernels in launch order still terminates -- and the map-level
  // kernels follow the chain on the caller's stream while the weight-gradient CTAs drain)
  // (layer-major backward: dmc is complete only behind the last launch, so the two independent halves of the
  // map-level backward -- dZ on the caller's stream, dW_0 / db_0 on the side stream -- run beside each other)
  cudaStream_t mstream = stream, dwstream = stream, w0stream = stream;
  if (dw_ctas > 0) {
    dwstream = side->stream;
  } else if (use_lbwd) {
    if (g_num_phase_events == 0 && !RENI_NO_FORK && dZ != nullptr) {
      if (side == nullptr && !side_stream(&side)) return RENI_ERR_CUDA;
      if (note(cudaEventRecord(side->fork, stream)) != cudaSuccess) return RENI_ERR_CUDA;
      if (note(cudaStreamWaitEvent(side->stream, side->fork, 0)) != cudaSuccess) return RENI_ERR_CUDA;
      w0stream = side->stream;
    }
  } else if (need_dw && film == nullptr && g_num_phase_events == 0 && !RENI_NO_FORK) {
    if (side == nullptr && !side_stream(&side)) return RENI_ERR_CUDA;
    if (note(cudaEventRecord(side->fork, stream)) != cudaSuccess) return RENI_ERR_CUDA;
    if (note(cudaStreamWaitEvent(side->stream, side->fork, 0)) != cudaSuccess) return RENI_ERR_CUDA;
    mstream = side->stream;
    w0stream = mstream;
  }

  if (need_dw && !use_lbwd) {
    DwParams q{};
    q.stash_u = at<uint16_t>(ws, w.stash_c);
    q.stash_d = at<__half>(ws, w.stash_d);
    q.stash_gy = at<__half>(ws, w.stash_gy);
    for (int i = 1; i <= L + 1; ++i) {
      if (want_dw && (host_dW[i] == nullptr || host_db[i] == nullptr)) return RENI_ERR_BAD_ARGUMENT;
      q.dW[i] = want_dw ? host_dW[i] : nullptr;
      q.db[i] = want_dw ? host_db[i] : nullptr;
    }
    q.njobs = want_dw ? L + 1 : L;  // (frozen FiLM decoder: the output-layer job has no consumer)
    q.tiles_per_map = (int)tiles_per_map(P);
    q.film_S = film != nullptr ? at<float>(ws, w.film_S) : nullptr;
    q.film_cs = film != nullptr ? at<float>(ws, w.film_cs) : nullptr;
    q.scalars = at<float>(ws, w.scalars);
    q.out_scale = c->last_layer_linear ? 1.f : c->hidden_omega_0;
    q.ntiles = ntiles;
    q.L = L;
    q.B = (int)B;
    q.out_features = c->out_features;
    if (note(cudaFuncSetAttribute(reni_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DwSmem::kTotal)) !=
        cudaSuccess)
      return RENI_ERR_CUDA;
    if (film != nullptr) {
      // one grid row per map: S[b][l] and cs[b][l] accumulate with atomics, then the map-level FiLM reduction
      if (cudaMemsetAsync(q.film_S, 0, (size_t)B * L * kH * kH * 4, stream) != cudaSuccess) return RENI_ERR_CUDA;
      if (cudaMemsetAsync(q.film_cs, 0, (size_t)B * L * kH * 4, stream) != cudaSuccess) return RENI_ERR_CUDA;
      // slices per (map, job): a CTA costs one pass over its share of the map's 2 * tiles_per_map stash blocks plus
      // a 256 KB flush of atomics (~ 4 blocks' worth of time); the grid runs in ceil(CTAs / SMs) rounds.  Pick the
      // split that minimises rounds * (blocks per CTA + flush).
      const int blocks_per_job = 2 * q.tiles_per_map;
      int slices = 1;
      double best = 1e30;
      for (int s_ = 1; s_ <= 16 && s_ <= blocks_per_job; ++s_) {
        const int64_t ctas = (int64_t)q.njobs * s_ * B;
        const double rounds = (double)((ctas + sms - 1) / sms);
        const double cost = rounds * ((double)blocks_per_job / s_ + 4.0);
        if (cost < best * 0.98) { best = cost; slices = s_; }
      }
      if (RENI_FILM_DW_PERSISTENT && want_dw) {
        // persistent: the SMs share out the list of all (map, job) stash blocks evenly and flush at item boundaries
        // (training step at 32 maps: dW + reductions 0.289 -> 0.254 ms; with a frozen decoder, whose four hidden jobs
        // per map quantise better on the grid below, it measured 3 % slower, hence training only)
        q.film_persistent = 1;
        const int64_t blocks = (int64_t)B * q.njobs * blocks_per_job;
        const int g = blocks < sms ? (int)blocks : sms;
        reni_dw_kernel<<<dim3((unsigned)g), kDwThreads, DwSmem::kTotal, stream>>>(q);
      } else {
        reni_dw_kernel<<<dim3((unsigned)(q.njobs * slices), (unsigned)B), kDwThreads, DwSmem::kTotal, stream>>>(q);
      }
      if (last_err() != cudaSuccess) return RENI_ERR_CUDA;
      FilmReduceParams r{};
      r.S = q.film_S;
      r.cs = q.film_cs;
      r.film = film->film;
      r.dfilm = film->d_film;
      r.B = (int)B;
      r.L = L;
      for (int i = 1; i <= L; ++i) {
        if (film->host_weights[i] == nullptr || film->host_biases[i] == nullptr) return RENI_ERR_BAD_ARGUMENT;
        r.W[i] = film->host_weights[i];
        r.b[i] = film->host_biases[i];
        r.dW[i] = want_dw ? host_dW[i] : nullptr;
        r.db[i] = want_dw ? host_db[i] : nullptr;
      }
      reni_film_dfilm_kernel<<<dim3(kH / 8, (unsigned)L, (unsigned)B), 256, 0, stream>>>(r);
      if (want_dw) reni_film_dw_kernel<<<dim3(kH / 4, (unsigned)L), 256, 0, stream>>>(r);
      if (last_err() != cudaSuccess) return RENI_ERR_CUDA;
      mark_phase(5, stream);
      mark_phase(6, stream);
      return RENI_OK;  // (layer 0 and the mapping network are differentiated by the caller from d_mc / d_film)
    }
    if (dw_ctas > 0) {
      q.ready = at<uint32_t>(ws, w.ready);
      q.stuck = at<uint32_t>(ws, w.ready) + ntiles;
      q.out_ctas = g_overlap_out_ctas;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)dw_ctas);
      cfg.blockDim = dim3(kDwThreads);
      cfg.dynamicSmemBytes = DwSmem::kTotal;
      cfg.stream = dwstream;
      cudaLaunchAttribute attr[1];  // allocated in SM pairs like the chain's clusters, so neither fragments the other
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (note(cudaLaunchKernelEx(&cfg, reni_dw_kernel, q)) != cudaSuccess) return RENI_ERR_CUDA;
    } else {
    int g = sms;
    const int max_useful = ntiles * 2 * (L + 1);
    if (g > max_useful) g = max_useful;
    if (g < L + 1) g = L + 1;
#if RENI_DW_BALANCE
    // bytes per stash block: hidden job 64 KB (delta + phase), output job 34 KB (g_y + phase) -> its share of the CTAs
    if (g >= 4 * (L + 1)) {
      q.out_ctas = (int)((double)g * 34.0 / (64.0 * L + 34.0) + 0.5);
      if (q.out_ctas < 1) q.out_ctas = 1;
    }
#endif
    reni_dw_kernel<<<g, kDwThreads, DwSmem::kTotal, stream>>>(q);
    }
    if (last_err() != cudaSuccess) return RENI_ERR_CUDA;
  }
  if (!use_lbwd) mark_phase(5, stream);

  {
    const int nin = (int)reni_in_features(c);
    const int K5 = 5 * (int)B;
    if (dZ != nullptr) {
      SmallGemmParams g{};  // E[(b,r), i] = sum_j dmc[(b,r), j] W0[j, i]
      g.A = at<float>(ws, w.dmc);
      g.B = weight0;
      g.C = at<float>(ws, w.dxc);
      g.M = K5; g.N = nin; g.K = kH; g.a_sk = 1; g.a_sm = kH; g.ldb = nin; g.ldc = nin;
      if (cudaMemsetAsync(g.C, 0, (size_t)K5 * nin * 4, mstream) != cudaSuccess) return RENI_ERR_CUDA;
      reni_small_gemm_kernel<<<dim3((nin + 63) / 64, (K5 + 63) / 64, kH / kSmallGemmK), 256, 0, mstream>>>(g);
      MapBwdParams m{};
      m.Z = Z;
      m.E = at<float>(ws, w.dxc);
      m.dZ = dZ;
      m.B = (int)B;
      m.N = c->ndims;
      m.in_features = nin;
      m.equivariance = c->equivariance;
      m.alpha2 = 2.f * alpha;
      m.accumulate = 0;
      reni_dz_kernel<<<(unsigned)B, 128, 0, mstream>>>(m);
    }
    if (need_dw) {
      if (host_dW[0] == nullptr || host_db[0] == nullptr) return RENI_ERR_BAD_ARGUMENT;
      SmallGemmParams g{};  // dW0[j, i] += sum_{(b,r)} dmc[(b,r), j] xfull[(b,r), i]
      g.A = at<float>(ws, w.dmc);
      g.B = at<float>(ws, w.xc);
      g.C = host_dW[0];
      g.M = kH; g.N = nin; g.K = K5; g.a_sk = kH; g.a_sm = 1; g.ldb = nin; g.ldc = nin;
      reni_small_gemm_kernel<<<dim3((nin + 63) / 64, kH / 64, (K5 + kSmallGemmK - 1) / kSmallGemmK), 256, 0, w0stream>>>(g);
      reni_db0_kernel<<<1, 256, 0, w0stream>>>(at<float>(ws, w.dmc), host_db[0], (int)B);
    }
    if (last_err() != cudaSuccess) return RENI_ERR_CUDA;
  }
  if (side != nullptr) {  // join
    if (note(cudaEventRecord(side->join, side->stream)) != cudaSuccess) return RENI_ERR_CUDA;
    if (note(cudaStreamWaitEvent(stream, side->join, 0)) != cudaSuccess) return RENI_ERR_CUDA;
  }
  mark_phase(6, stream);
  return RENI_OK;
}

int32_t reni_debug_set_trace(void* device_buffer) {
  g_trace = static_cast<unsigned long long*>(device_buffer);
  return RENI_OK;
}

const char* reni_debug_last_cuda_error(void) { return cudaGetErrorString(g_last_cuda); }

int32_t reni_phase_bits(void) { return RENI_PHASE_BITS; }

int32_t reni_debug_set_overlap(int32_t dw_ctas, int32_t out_ctas) {
  if (dw_ctas < -1 || out_ctas < 0) return RENI_ERR_BAD_ARGUMENT;
  g_overlap_dw_ctas = dw_ctas;
  g_overlap_out_ctas = out_ctas;
  return RENI_OK;
}

int32_t reni_debug_set_phase_events(void* const* events, int32_t n) {
  if (n < 0 || n > 16 || (n > 0 && events == nullptr)) return RENI_ERR_BAD_ARGUMENT;
  for (int i = 0; i < n; ++i) g_phase_events[i] = static_cast<cudaEvent_t>(events[i]);
  g_num_phase_events = n;
  return RENI_OK;
}

int32_t reni_backward(const reni_config_t* c, const float* Z, const float* D, int64_t d_bstride,
                      const float* const* host_weights, int64_t B, int64_t P, const float* out,
                      const float* grad_out, float* dZ, float* const* host_dW, float* const* host_db, void* ws,
                      int64_t ws_bytes, int32_t flags, void* stream_) {
  if (!config_ok(c)) return RENI_ERR_BAD_CONFIG;
  if (Z == nullptr || (D == nullptr && !(flags & RENI_FLAG_GRID_DIRECTIONS)) || host_weights == nullptr ||
      host_weights[0] == nullptr || out == nullptr || grad_out == nullptr || ws == nullptr || B < 1 || P < 1)
    return RENI_ERR_BAD_ARGUMENT;
  flags &= ~RENI_FLAG_GRID_SINEWEIGHT;  // (no sine weights on this path: the gradient comes from the caller)
  if (!(flags & RENI_FLAG_SAVE_FOR_BACKWARD)) return RENI_ERR_BAD_ARGUMENT;
  if ((flags & RENI_FLAG_NEED_DW) && (host_dW == nullptr || host_db == nullptr)) return RENI_ERR_BAD_ARGUMENT;
  const WorkspaceLayout w = make_layout(c, B, P, flags);
  if (ws_bytes < w.total || (reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return RENI_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* scalars = at<float>(ws, w.scalars);
  unsigned int* slot = reinterpret_cast<unsigned int*>(scalars + 2);
  if (cudaMemsetAsync(slot, 0, 4, stream) != cudaSuccess) return RENI_ERR_CUDA;
  const int64_t n = B * P * 3;
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 1184) blocks = 1184;
  reni_absmax_kernel<<<blocks, 256, 0, stream>>>(grad_out, n, slot);
  reni_scale_from_absmax_kernel<<<1, 1, 0, stream>>>(slot, scalars);
  if (last_err() != cudaSuccess) return RENI_ERR_CUDA;
  return launch_backward(c, w, Z, D, d_bstride, host_weights[0], B, P, out, grad_out, nullptr, nullptr, 0, 0.f, dZ,
                         host_dW, host_db, ws, flags, stream);
}

int32_t reni_loss_forward_backward(const reni_config_t* c, const float* Z, const float* D, int64_t d_bstride,
                                   const float* const* host_weights, const float* const* host_biases, int64_t B,
                                   int64_t P, const float* target, const float* sw, int64_t sw_bstride, float alpha,
                                   float beta, int32_t use_cosine, float* out, float* loss_out, float* dZ,
                                   float* const* host_dW, float* const* host_db, void* ws, int64_t ws_bytes,
                                   int32_t flags, void* stream_) {
  if (!config_ok(c)) return RENI_ERR_BAD_CONFIG;
  if (host_weights == nullptr || host_biases == nullptr || host_weights[0] == nullptr || host_biases[0] == nullptr ||
      target == nullptr || (sw == nullptr && !(flags & RENI_FLAG_GRID_SINEWEIGHT)) || out == nullptr ||
      loss_out == nullptr)
    return RENI_ERR_BAD_ARGUMENT;
  flags |= RENI_FLAG_SAVE_FOR_BACKWARD | RENI_FLAG_LOSS;
  if ((flags & RENI_FLAG_NEED_DW) && (host_dW == nullptr || host_db == nullptr)) return RENI_ERR_BAD_ARGUMENT;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int32_t rc;
  if ((flags & RENI_FLAG_PREPARE_WEIGHTS) && g_num_phase_events == 0 && !RENI_NO_FORK) {
    // the fp16 weight images are built on the side stream while the per-map prologue runs on the caller's
    SideStream* ps = nullptr;
    if (!side_stream(&ps)) return RENI_ERR_CUDA;
    if (ps->step[0] == nullptr && note(cudaEventCreateWithFlags(&ps->step[0], cudaEventDisableTiming)) != cudaSuccess)
      return RENI_ERR_CUDA;
    if (note(cudaEventRecord(ps->fork, stream)) != cudaSuccess) return RENI_ERR_CUDA;
    if (note(cudaStreamWaitEvent(ps->stream, ps->fork, 0)) != cudaSuccess) return RENI_ERR_CUDA;
    rc = launch_prepare_weights(c, host_weights, host_biases, ws, ws_bytes, ps->stream);
    if (rc != RENI_OK) return rc;
    if (note(cudaEventRecord(ps->step[0], ps->stream)) != cudaSuccess) return RENI_ERR_CUDA;
    g_fwd_wait_event = ps->step[0];
  } else if (flags & RENI_FLAG_PREPARE_WEIGHTS) {
    rc = launch_prepare_weights(c, host_weights, host_biases, ws, ws_bytes, stream);
    if (rc != RENI_OK) return rc;
  }
  rc = reni_forward(c, Z, D, d_bstride, host_weights[0], host_biases[0], B, P, out, target, sw, sw_bstride, ws,
                    ws_bytes, flags, stream_);
  g_fwd_wait_event = nullptr;
  if (rc != RENI_OK) return rc;
  const WorkspaceLayout w = make_layout(c, B, P, flags);
  // Without the cosine term the backward needs nothing from the loss reduction (S = 3P/2 is written by the prologue,
  // the per-map cosine coefficients are not read): the reduction then runs on the side stream under the delta chain.
  SideStream* side = nullptr;
  cudaStream_t lstream = stream;
  if (!use_cosine && g_num_phase_events == 0 && !RENI_NO_FORK) {
    if (!side_stream(&side)) return RENI_ERR_CUDA;
    if (note(cudaEventRecord(side->fork, stream)) != cudaSuccess) return RENI_ERR_CUDA;
    if (note(cudaStreamWaitEvent(side->stream, side->fork, 0)) != cudaSuccess) return RENI_ERR_CUDA;
    lstream = side->stream;
  }
  if (cudaMemsetAsync(loss_out, 0, 16, lstream) != cudaSuccess) return RENI_ERR_CUDA;
  LossFinishParams f{};
  f.loss_part = at<float>(ws, w.loss_part);
  f.sw = sw;
  f.sw_bstride = sw_bstride;
  {
    GridArgs ga;
    if (!grid_args(flags, P, D, sw, &ga)) return RENI_ERR_BAD_ARGUMENT;
    f.grid_w = ga.w;
    f.sw_grid = ga.sw;
    f.mask_bits = ga.mask;
  }
  f.Z = (alpha != 0.f) ? Z : nullptr;
  f.map_loss = at<float>(ws, w.map_loss);
  f.loss_out = loss_out;
  f.scalars = at<float>(ws, w.scalars);
  f.B = (int)B;
  f.P = (int)P;
  f.tiles_per_map = (int)tiles_per_map(P);
  f.nz = c->ndims * 3;
  f.alpha = alpha;
  f.beta = beta;
  f.use_cos = use_cosine;
  reni_loss_finish_kernel<<<(unsigned)B, 128, 0, lstream>>>(f);
  if (last_err() != cudaSuccess) return RENI_ERR_CUDA;
  rc = launch_backward(c, w, Z, D, d_bstride, host_weights[0], B, P, out, nullptr, target, sw, sw_bstride, alpha, dZ,
                       host_dW, host_db, ws, flags, stream, use_cosine, side);
  return rc;  // (launch_backward joins the side stream before it returns)
}

int32_t reni_film_forward(const reni_config_t* c, const float* mc, const float* film, const float* D,
                          int64_t d_bstride, int64_t B, int64_t P, float* out, void* ws, int64_t ws_bytes,
                          int32_t flags, void* stream_) {
  if (!config_ok(c) || !c->last_layer_linear) return RENI_ERR_BAD_CONFIG;
  if (mc == nullptr || film == nullptr || D == nullptr || out == nullptr || ws == nullptr || B < 1 || P < 1)
    return RENI_ERR_BAD_ARGUMENT;
  flags = (flags & (RENI_FLAG_SAVE_FOR_BACKWARD | RENI_FLAG_FILM_PERMAP)) | RENI_FLAG_FILM;
  if ((flags & RENI_FLAG_FILM_PERMAP) && !permap_ok(P)) return RENI_ERR_BAD_ARGUMENT;
  const WorkspaceLayout w = make_layout(c, B, P, flags);
  if (ws_bytes < w.total || (reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return RENI_ERR_WORKSPACE;
  const int sms = num_sms();
  if (sms <= 0) return RENI_ERR_NO_DEVICE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  mark_phase(0, stream);
  mark_phase(1, stream);
  return launch_forward(c, w, mc, film, D, d_bstride, B, P, out, nullptr, nullptr, 0, ws, flags, stream, sms);
}

int32_t reni_film_backward(const reni_config_t* c, const float* film, const float* D, int64_t d_bstride,
                           const float* const* host_weights, const float* const* host_biases, int64_t B, int64_t P,
                           const float* out, const float* grad_out, float* d_mc, float* d_film,
                           float* const* host_dW, float* const* host_db, void* ws, int64_t ws_bytes, int32_t flags,
                           void* stream_) {
  if (!config_ok(c) || !c->last_layer_linear) return RENI_ERR_BAD_CONFIG;
  if (film == nullptr || D == nullptr || host_weights == nullptr || host_biases == nullptr || out == nullptr ||
      grad_out == nullptr || d_mc == nullptr || d_film == nullptr || ws == nullptr || B < 1 || P < 1)
    return RENI_ERR_BAD_ARGUMENT;
  if ((flags & RENI_FLAG_NEED_DW) && (host_dW == nullptr || host_db == nullptr)) return RENI_ERR_BAD_ARGUMENT;
  flags = (flags & (RENI_FLAG_NEED_DW | RENI_FLAG_FILM_PERMAP)) | RENI_FLAG_SAVE_FOR_BACKWARD | RENI_FLAG_FILM;
  if ((flags & RENI_FLAG_FILM_PERMAP) && !permap_ok(P)) return RENI_ERR_BAD_ARGUMENT;
  const WorkspaceLayout w = make_layout(c, B, P, flags);
  if (ws_bytes < w.total || (reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return RENI_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* scalars = at<float>(ws, w.scalars);
  unsigned int* slot = reinterpret_cast<unsigned int*>(scalars + 2);
  if (cudaMemsetAsync(slot, 0, 4, stream) != cudaSuccess) return RENI_ERR_CUDA;
  const int64_t n = B * P * 3;
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 1184) blocks = 1184;
  reni_absmax_kernel<<<blocks, 256, 0, stream>>>(grad_out, n, slot);
  reni_scale_from_absmax_kernel<<<1, 1, 0, stream>>>(slot, scalars);
  if (last_err() != cudaSuccess) return RENI_ERR_CUDA;
  FilmBackwardArgs fa{film, d_mc, d_film, host_weights, host_biases};
  return launch_backward(c, w, nullptr, D, d_bstride, nullptr, B, P, out, grad_out, nullptr, nullptr, 0, 0.f, nullptr,
                         host_dW, host_db, ws, flags, stream, 1, nullptr, &fa);
}

int32_t reni_adam_step(const reni_adam_segment_t* host_segments, int32_t nseg, int32_t* step, double lr, double beta1,
                       double beta2, double eps, void* stream_) {
  if (host_segments == nullptr || nseg < 1 || step == nullptr) return RENI_ERR_BAD_ARGUMENT;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int sms = num_sms();
  if (sms <= 0) return RENI_ERR_NO_DEVICE;
  for (int base = 0; base < nseg; base += kAdamMaxSegments) {
    AdamParams a{};
    const int n = nseg - base < kAdamMaxSegments ? nseg - base : kAdamMaxSegments;
    int64_t total = 0;
    for (int i = 0; i < n; ++i) {
      const reni_adam_segment_t& sgm = host_segments[base + i];
      if (sgm.param == nullptr || sgm.grad == nullptr || sgm.exp_avg == nullptr || sgm.exp_avg_sq == nullptr ||
          sgm.numel < 0)
        return RENI_ERR_BAD_ARGUMENT;
      a.p[i] = sgm.param;
      a.g[i] = sgm.grad;
      a.m[i] = sgm.exp_avg;
      a.v[i] = sgm.exp_avg_sq;
      total += sgm.numel;
      a.end[i] = total;
    }
    if (total == 0) continue;
    a.nseg = n;
    a.step = step;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
    int64_t blocks = (total + 1023) / 1024;  // ~4 elements per thread
    if (blocks > 8LL * sms) blocks = 8LL * sms;
    reni_adam_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a);
  }
  reni_adam_advance_kernel<<<1, 1, 0, stream>>>(step);
  return last_err() == cudaSuccess ? RENI_OK : RENI_ERR_CUDA;
}

int32_t reni_vad_sample(const float* mu, const float* log_var, const int64_t* idx, const float* eps, int64_t B,
                        int64_t nz, float* Z, void* stream_) {
  if (mu == nullptr || log_var == nullptr || idx == nullptr || eps == nullptr || Z == nullptr || B < 1 || nz < 1)
    return RENI_ERR_BAD_ARGUMENT;
  VadParams p{};
  p.mu = mu; p.log_var = log_var; p.idx = idx; p.eps = eps; p.Z = Z;
  p.B = (int)B; p.nz = (int)nz;
  reni_vad_sample_kernel<<<(unsigned)B, 128, 0, static_cast<cudaStream_t>(stream_)>>>(p);
  return last_err() == cudaSuccess ? RENI_OK : RENI_ERR_CUDA;
}

int32_t reni_vad_backward(const float* mu, const float* log_var, const int64_t* idx, const float* eps, const float* dZ,
                          int64_t B, int64_t nz, float kld_weight_over_zdims, float grad_scale, float* dmu,
                          float* dlog_var, float* kld_out, void* stream_) {
  if (mu == nullptr || log_var == nullptr || idx == nullptr || eps == nullptr || dZ == nullptr || dmu == nullptr ||
      dlog_var == nullptr || kld_out == nullptr || B < 1 || nz < 1)
    return RENI_ERR_BAD_ARGUMENT;
  VadParams p{};
  p.mu = mu; p.log_var = log_var; p.idx = idx; p.eps = eps; p.dZ = dZ;
  p.dmu = dmu; p.dlog_var = dlog_var; p.kld_out = kld_out;
  p.B = (int)B; p.nz = (int)nz; p.kw = kld_weight_over_zdims; p.grad_scale = grad_scale;
  reni_vad_backward_kernel<<<(unsigned)B, 128, 0, static_cast<cudaStream_t>(stream_)>>>(p);
  return last_err() == cudaSuccess ? RENI_OK : RENI_ERR_CUDA;
}

static int32_t launch_shade(bool backward, const float* normals, const float* view, int64_t n_pix, const float* D,
                            int64_t d_bstride, const float* light, const float* grad, int64_t B, int64_t J, float kd,
                            float ks, float shininess, float* colors, float* d_light, cudaStream_t stream) {
  if (normals == nullptr || view == nullptr || D == nullptr || n_pix < 1 || B < 1 || J < 1 || B > 65535 ||
      !(shininess >= 0.f))
    return RENI_ERR_BAD_ARGUMENT;
  ShadeParams p{};
  p.normals = normals; p.view = view; p.D = D; p.d_bstride = d_bstride;
  p.light = light; p.grad = grad; p.colors = colors; p.d_light = d_light;
  p.B = (int)B; p.J = (int)J; p.n_pix = (int)n_pix;
  p.kd = kd;
  // Blinn-Phong normalisation of the specular lobe (pytorch3d_envmap_shader.py:113-115)
  p.cks = ks * (shininess + 2.f) / (4.f * (2.f - expf(-shininess / 2.f)));
  p.shininess = shininess;
  const bool shared = d_bstride == 0;
  const unsigned chunks = (unsigned)(shared ? (B + kShadeMaps - 1) / kShadeMaps : B);
  const int64_t n = backward ? J : n_pix;           // one thread per output row
  const int64_t red = backward ? n_pix : J;         // reduction axis
  const int64_t rows = (n + kShadeThreads - 1) / kShadeThreads * chunks;
  // enough blocks for ~8 per SM, each with at least two tiles of the reduction axis
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  int64_t split = (8LL * sms + rows - 1) / rows;
  const int64_t max_split = (red + 2 * kShadeTile - 1) / (2 * kShadeTile);
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  p.split = (int)split;
  if (split > 1) {
    float* outp = backward ? d_light : colors;
    if (cudaMemsetAsync(outp, 0, (size_t)B * n * 3 * sizeof(float), stream) != cudaSuccess) return RENI_ERR_CUDA;
  }
  const dim3 grid((unsigned)((n + kShadeThreads - 1) / kShadeThreads), chunks, (unsigned)split);
  if (backward) {
    if (shared) reni_shade_bwd_kernel<true><<<grid, kShadeThreads, 0, stream>>>(p);
    else reni_shade_bwd_kernel<false><<<grid, kShadeThreads, 0, stream>>>(p);
  } else {
    if (shared) reni_shade_fwd_kernel<true><<<grid, kShadeThreads, 0, stream>>>(p);
    else reni_shade_fwd_kernel<false><<<grid, kShadeThreads, 0, stream>>>(p);
  }
  return last_err() == cudaSuccess ? RENI_OK : RENI_ERR_CUDA;
}

int32_t reni_envmap_shade_forward(const float* normals, const float* view_dirs, int64_t n_pix, const float* D,
                                  int64_t d_bstride, const float* light, int64_t B, int64_t J, float kd, float ks,
                                  float shininess, float* colors, void* stream_) {
  if (light == nullptr || colors == nullptr) return RENI_ERR_BAD_ARGUMENT;
  return launch_shade(false, normals, view_dirs, n_pix, D, d_bstride, light, nullptr, B, J, kd, ks, shininess, colors,
                      nullptr, static_cast<cudaStream_t>(stream_));
}

int32_t reni_envmap_shade_backward(const float* normals, const float* view_dirs, int64_t n_pix, const float* D,
                                   int64_t d_bstride, const float* grad_colors, int64_t B, int64_t J, float kd, float ks,
                                   float shininess, float* d_light, void* stream_) {
  if (grad_colors == nullptr || d_light == nullptr) return RENI_ERR_BAD_ARGUMENT;
  return launch_shade(true, normals, view_dirs, n_pix, D, d_bstride, nullptr, grad_colors, B, J, kd, ks, shininess,
                      nullptr, d_light, static_cast<cudaStream_t>(stream_));
}

int64_t reni_allreduce_flag_bytes(void) { return (int64_t)kArMaxBlocks * kArMaxWorld * sizeof(uint32_t); }

int32_t reni_allreduce(void* dev_buf_ptrs, void* dev_flag_ptrs, void* multicast_ptr, int64_t numel, int32_t rank,
                       int32_t world, float scale, uint32_t* epoch, uint32_t* status, void* stream_) {
  if (dev_buf_ptrs == nullptr || dev_flag_ptrs == nullptr || epoch == nullptr || status == nullptr || numel < 4 ||
      (numel & 3) != 0 || world < 1 || world > kArMaxWorld || rank < 0 || rank >= world)
    return RENI_ERR_BAD_ARGUMENT;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  AllReduceParams p{};
  p.bufs = static_cast<float* const*>(dev_buf_ptrs);
  p.flags = static_cast<uint32_t* const*>(dev_flag_ptrs);
  p.mc = static_cast<float*>(multicast_ptr);
  p.n4 = numel / 4;
  p.rank = rank;
  p.world = world;
  p.scale = scale;
  p.epoch = epoch;
  p.status = status;
  // one block per SM at most, each thread with kArUnroll 16-byte elements in flight (the grid only depends on numel
  // and world, so every rank -- and every replay of a captured call -- launches the same blocks)
  const int64_t per = (p.n4 + world - 1) / world;
  int blocks = (int)((per + (int64_t)kArThreads * kArUnroll - 1) / ((int64_t)kArThreads * kArUnroll));
  if (blocks < 1) blocks = 1;
  const int sms = num_sms();
  const int cap = sms > 0 && sms < kArMaxBlocks ? sms : kArMaxBlocks;
  if (blocks > cap) blocks = cap;
  reni_allreduce_kernel<<<blocks, kArThreads, 0, stream>>>(p);
  return last_err() == cudaSuccess ? RENI_OK : RENI_ERR_CUDA;
}

#ifndef RENI_FILM_MAP_FUSED
#define RENI_FILM_MAP_FUSED 0  // 1: per-map stage as one cooperative launch per direction (film_map_fused.cuh) instead of
                               // the staged launches -- measured SLOWER at 32 maps (forward 60 vs 47 us, forward +
                               // backward 190 vs 154 us in a replayed graph): its stages are bound by the same L2 latency
                               // per k-step as the staged kernels and the barriers cost what the launches did
#endif
constexpr int64_t kFilmMapSyncBytes = 256;

// cooperative launch of a fused per-map kernel on at most kMapFusedMaxCtas CTAs; `sync` = its two barrier words
extern "C++" {
template <typename P>
static int32_t launch_map_fused(void (*kernel)(P), const P& p, unsigned int* sync, cudaStream_t stream) {
  const int sms = num_sms();
  if (sms <= 0) return RENI_ERR_NO_DEVICE;
  if (cudaMemsetAsync(sync, 0, 8, stream) != cudaSuccess) return RENI_ERR_CUDA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(sms < kMapFusedMaxCtas ? sms : kMapFusedMaxCtas));
  cfg.blockDim = dim3(kMapFusedThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return note(cudaLaunchKernelEx(&cfg, kernel, p)) == cudaSuccess ? RENI_OK : RENI_ERR_CUDA;
}
}  // extern "C++"

int64_t reni_film_map_scratch_bytes(const int32_t* host_map_dims, int32_t n_linears, int64_t B) {
  if (host_map_dims == nullptr || n_linears < 1 || n_linears > kFilmMapMaxLinears || B < 1) return RENI_ERR_BAD_ARGUMENT;
  int64_t widest = 0;
  for (int i = 0; i <= n_linears; ++i) {
    if (host_map_dims[i] < 1) return RENI_ERR_BAD_ARGUMENT;
    if (host_map_dims[i] > widest) widest = host_map_dims[i];
  }
  // two ping-pong activation buffers + the grid-barrier words of the fused launch
  return 2 * align_up(B * widest * (int64_t)sizeof(float), 256) + kFilmMapSyncBytes;
}

// Shared body of the two per-map forwards: act[i] = input of linear i (act[0] = mapping input), act[n] = raw output
static int32_t film_map_forward_impl(const reni_config_t* c, const float* Z, const float* weight0, const float* bias0,
                                     const float* const* host_map_weights, const float* const* host_map_biases,
                                     const int32_t* host_map_dims, int32_t n_linears, int64_t B, float* mc, float* film,
                                     float* const* act, unsigned int* sync, cudaStream_t stream) {
  const int N = c->ndims, Lf = c->hidden_layers + 1;
  const int mn_in = c->equivariance == RENI_EQ_SO2 ? N * N + N : N * N;
  const int so2 = c->equivariance == RENI_EQ_SO2;
  if (RENI_FILM_MAP_FUSED) {
    FilmMapFusedParams q{};
    q.Z = Z; q.W0 = weight0; q.b0 = bias0;
    for (int i = 0; i < n_linears; ++i) {
      if (host_map_weights[i] == nullptr || host_map_biases[i] == nullptr) return RENI_ERR_BAD_ARGUMENT;
      q.W[i] = host_map_weights[i];
      q.bias[i] = host_map_biases[i];
    }
    for (int i = 0; i <= n_linears; ++i) {
      q.act[i] = act[i];
      q.dims[i] = host_map_dims[i];
    }
    q.n_linears = n_linears;
    q.mc = mc; q.film = film;
    q.B = (int)B; q.N = N; q.so2 = so2; q.Lf = Lf;
    q.sync = sync;
    return launch_map_fused(reni_film_map_fused_fwd_kernel, q, sync, stream);
  }
  reni_film_map_input_kernel<<<(unsigned)B, 256, 3 * N * sizeof(float), stream>>>(Z, act[0], N, so2, mn_in);
  for (int i = 0; i < n_linears; ++i) {
    if (host_map_weights[i] == nullptr || host_map_biases[i] == nullptr) return RENI_ERR_BAD_ARGUMENT;
    const int in = host_map_dims[i], out = host_map_dims[i + 1];
    const int leaky = i + 1 < n_linears ? 1 : 0;
    // four maps per block (each weight read once for the four) when there are enough maps and their inputs fit
    const bool quad = B >= 16 && (size_t)4 * in * sizeof(float) <= 200 * 1024;  // (fewer maps: the wider grid wins)
    const size_t smem = (size_t)(quad ? 4 : 1) * in * sizeof(float);
    if (smem > 200 * 1024) return RENI_ERR_BAD_CONFIG;
    auto kernel = quad ? reni_film_map_linear_kernel<4> : reni_film_map_linear_kernel<1>;
    if (smem > 48 * 1024 &&
        note(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
      return RENI_ERR_CUDA;
    kernel<<<dim3((unsigned)((out + 31) / 32), (unsigned)((B + (quad ? 3 : 0)) / (quad ? 4 : 1))), 256, smem, stream>>>(
        act[i], host_map_weights[i], host_map_biases[i], act[i + 1], (int)B, in, out, leaky);
  }
  FilmMapFinishParams f{};
  f.Z = Z;
  f.W0 = weight0;
  f.b0 = bias0;
  f.raw = act[n_linears];
  f.mc = mc;
  f.film = film;
  f.N = N;
  f.so2 = so2;
  f.Lf = Lf;
  reni_film_map_finish_kernel<<<(unsigned)B, 256, 3 * N * sizeof(float), stream>>>(f);
  return last_err() == cudaSuccess ? RENI_OK : RENI_ERR_CUDA;
}

static bool film_map_args_ok(const reni_config_t* c, const int32_t* host_map_dims, int32_t n_linears, int64_t B) {
  if (host_map_dims == nullptr || B < 1 || B > 65535 || n_linears < 1 || n_linears > kFilmMapMaxLinears) return false;
  const int N = c->ndims, Lf = c->hidden_layers + 1;
  const int mn_in = c->equivariance == RENI_EQ_SO2 ? N * N + N : N * N;
  return host_map_dims[0] == mn_in && host_map_dims[n_linears] == 2 * Lf * kH;
}

int32_t reni_film_map_forward(const reni_config_t* c, const float* Z, const float* weight0, const float* bias0,
                              const float* const* host_map_weights, const float* const* host_map_biases,
                              const int32_t* host_map_dims, int32_t n_linears, int64_t B, float* mc, float* film,
                              void* scratch, int64_t scratch_bytes, void* stream_) {
  if (!config_ok(c) || c->equivariance == RENI_EQ_NONE) return RENI_ERR_BAD_CONFIG;
  if (Z == nullptr || weight0 == nullptr || bias0 == nullptr || host_map_weights == nullptr ||
      host_map_biases == nullptr || mc == nullptr || film == nullptr || scratch == nullptr ||
      !film_map_args_ok(c, host_map_dims, n_linears, B))
    return RENI_ERR_BAD_ARGUMENT;
  const int64_t need = reni_film_map_scratch_bytes(host_map_dims, n_linears, B);
  if (need < 0) return (int32_t)need;
  if (scratch_bytes < need) return RENI_ERR_WORKSPACE;
  float* act[kFilmMapMaxLinears + 1];  // two ping-pong buffers
  const int64_t half = (need - kFilmMapSyncBytes) / 2;
  for (int i = 0; i <= n_linears; ++i)
    act[i] = reinterpret_cast<float*>(static_cast<uint8_t*>(scratch) + (i & 1) * half);
  return film_map_forward_impl(c, Z, weight0, bias0, host_map_weights, host_map_biases, host_map_dims, n_linears, B, mc,
                               film, act, reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(scratch) + 2 * half),
                               static_cast<cudaStream_t>(stream_));
}

int64_t reni_film_map_acts_bytes(const int32_t* host_map_dims, int32_t n_linears, int64_t B) {
  if (host_map_dims == nullptr || n_linears < 1 || n_linears > kFilmMapMaxLinears || B < 1) return RENI_ERR_BAD_ARGUMENT;
  int64_t total = 0;
  for (int i = 0; i <= n_linears; ++i) {
    if (host_map_dims[i] < 1) return RENI_ERR_BAD_ARGUMENT;
    total += align_up(B * host_map_dims[i] * (int64_t)sizeof(float), 256);
  }
  return total + kFilmMapSyncBytes;  // (+ the grid-barrier words of the fused launches, behind the last activation)
}

static void film_map_act_ptrs(float* base, const int32_t* dims, int32_t n_linears, int64_t B, float** act) {
  int64_t off = 0;
  for (int i = 0; i <= n_linears; ++i) {
    act[i] = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(base) + off);
    off += align_up(B * dims[i] * (int64_t)sizeof(float), 256);
  }
}

int32_t reni_film_map_forward_train(const reni_config_t* c, const float* Z, const float* weight0, const float* bias0,
                                    const float* const* host_map_weights, const float* const* host_map_biases,
                                    const int32_t* host_map_dims, int32_t n_linears, int64_t B, float* mc, float* film,
                                    void* acts, int64_t acts_bytes, void* stream_) {
  if (!config_ok(c) || c->equivariance == RENI_EQ_NONE) return RENI_ERR_BAD_CONFIG;
  if (Z == nullptr || weight0 == nullptr || bias0 == nullptr || host_map_weights == nullptr ||
      host_map_biases == nullptr || mc == nullptr || film == nullptr || acts == nullptr ||
      !film_map_args_ok(c, host_map_dims, n_linears, B))
    return RENI_ERR_BAD_ARGUMENT;
  const int64_t need = reni_film_map_acts_bytes(host_map_dims, n_linears, B);
  if (need < 0) return (int32_t)need;
  if (acts_bytes < need) return RENI_ERR_WORKSPACE;
  float* act[kFilmMapMaxLinears + 1];
  film_map_act_ptrs(static_cast<float*>(acts), host_map_dims, n_linears, B, act);
  return film_map_forward_impl(c, Z, weight0, bias0, host_map_weights, host_map_biases, host_map_dims, n_linears, B, mc,
                               film, act,
                               reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(acts) + need - kFilmMapSyncBytes),
                               static_cast<cudaStream_t>(stream_));
}

int32_t reni_film_map_backward(const reni_config_t* c, const float* Z, const float* weight0, const float* bias0,
                               const float* const* host_map_weights, const int32_t* host_map_dims, int32_t n_linears,
                               int64_t B, const void* acts, const float* d_mc, const float* d_film, float* dZ,
                               float* dW0, float* db0, float* const* host_map_dW, float* const* host_map_db,
                               void* scratch, int64_t scratch_bytes, void* stream_) {
  if (!config_ok(c) || c->equivariance == RENI_EQ_NONE) return RENI_ERR_BAD_CONFIG;
  if (Z == nullptr || weight0 == nullptr || bias0 == nullptr || host_map_weights == nullptr || acts == nullptr ||
      d_mc == nullptr || d_film == nullptr || dZ == nullptr || scratch == nullptr ||
      !film_map_args_ok(c, host_map_dims, n_linears, B))
    return RENI_ERR_BAD_ARGUMENT;
  const bool want_dw = dW0 != nullptr;  // (a frozen decoder only wants dZ)
  if (want_dw && (db0 == nullptr || host_map_dW == nullptr || host_map_db == nullptr)) return RENI_ERR_BAD_ARGUMENT;
  for (int i = 0; i < n_linears; ++i) {  // (every pointer is checked before anything is forked onto the side stream)
    if (host_map_weights[i] == nullptr) return RENI_ERR_BAD_ARGUMENT;
    if (want_dw && (host_map_dW[i] == nullptr || host_map_db[i] == nullptr)) return RENI_ERR_BAD_ARGUMENT;
  }
  const int64_t acts_bytes = reni_film_map_acts_bytes(host_map_dims, n_linears, B);
  if (acts_bytes < 0) return (int32_t)acts_bytes;
  const int64_t dm_bytes = B * 4 * kH * (int64_t)sizeof(float);
  if (scratch_bytes < acts_bytes + dm_bytes) return RENI_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int N = c->ndims, Lf = c->hidden_layers + 1;
  const int so2 = c->equivariance == RENI_EQ_SO2;
  float* act[kFilmMapMaxLinears + 1];
  float* dact[kFilmMapMaxLinears + 1];  // gradient w.r.t. act[i]
  film_map_act_ptrs(const_cast<float*>(static_cast<const float*>(acts)), host_map_dims, n_linears, B, act);
  film_map_act_ptrs(static_cast<float*>(scratch), host_map_dims, n_linears, B, dact);
  float* dM = reinterpret_cast<float*>(static_cast<uint8_t*>(scratch) + acts_bytes);
  // the dX kernels add their output-row slices with atomics
  if (cudaMemsetAsync(scratch, 0,
                      (size_t)(acts_bytes - kFilmMapSyncBytes - align_up(B * host_map_dims[n_linears] * 4, 256)),
                      stream) != cudaSuccess)
    return RENI_ERR_CUDA;
  if (RENI_FILM_MAP_FUSED) {
    FilmMapFusedBwdParams q{};
    q.Z = Z; q.W0 = weight0; q.b0 = bias0;
    for (int i = 0; i < n_linears; ++i) {
      q.W[i] = host_map_weights[i];
      q.dW[i] = want_dw ? host_map_dW[i] : nullptr;
      q.db[i] = want_dw ? host_map_db[i] : nullptr;
    }
    for (int i = 0; i <= n_linears; ++i) {
      q.act[i] = act[i];
      q.dact[i] = dact[i];
      q.dims[i] = host_map_dims[i];
    }
    q.n_linears = n_linears;
    q.d_mc = d_mc; q.d_film = d_film; q.dM = dM; q.dZ = dZ;
    q.dW0 = want_dw ? dW0 : nullptr; q.db0 = want_dw ? db0 : nullptr;
    q.B = (int)B; q.N = N; q.so2 = so2; q.Lf = Lf;
    q.sync = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(scratch) + acts_bytes - kFilmMapSyncBytes);
    return launch_map_fused(reni_film_map_fused_bwd_kernel, q, q.sync, stream);
  }
  {
    FilmMapBwdHeadParams h{};
    h.Z = Z; h.W0 = weight0; h.b0 = bias0; h.raw = act[n_linears]; h.d_mc = d_mc; h.d_film = d_film;
    h.d_raw = dact[n_linears]; h.dM = dM; h.N = N; h.so2 = so2; h.Lf = Lf;
    reni_film_map_bwd_head_kernel<<<(unsigned)B, 256, 3 * N * sizeof(float), stream>>>(h);
  }
  // The weight-gradient kernels have no consumer inside this call: they run on the side stream, each behind the event
  // that marks its dY ready, while the dX chain (the critical path to dZ) continues on the caller's stream.
  SideStream* side = nullptr;
  cudaStream_t wstream = stream;
  if (want_dw && !RENI_NO_FORK) {
    if (!side_stream(&side)) return RENI_ERR_CUDA;
    wstream = side->stream;
  }
  auto fork_to_side = [&](int slot) -> bool {
    if (side == nullptr) return true;
    if (side->step[slot] == nullptr &&
        note(cudaEventCreateWithFlags(&side->step[slot], cudaEventDisableTiming)) != cudaSuccess)
      return false;
    return note(cudaEventRecord(side->step[slot], stream)) == cudaSuccess &&
           note(cudaStreamWaitEvent(side->stream, side->step[slot], 0)) == cudaSuccess;
  };
  if (want_dw) {
    if (!fork_to_side(n_linears)) return RENI_ERR_CUDA;
    FilmMapBwdW0Params q{};
    q.Z = Z; q.raw = act[n_linears]; q.d_mc = d_mc; q.dM = dM; q.dW0 = dW0; q.db0 = db0;
    q.B = (int)B; q.N = N; q.so2 = so2; q.Lf = Lf;
    reni_film_map_bwd_w0_kernel<<<kH / 8, 256, 0, wstream>>>(q);
  }
  for (int i = n_linears - 1; i >= 0; --i) {
    const int in = host_map_dims[i], out = host_map_dims[i + 1];
    const int leaky = i + 1 < n_linears ? 1 : 0;
    if (want_dw) {
      if (i + 1 < n_linears && !fork_to_side(i + 1)) return RENI_ERR_CUDA;  // dact[i + 1] has just been produced
      reni_film_map_bwd_dw_kernel<<<dim3((unsigned)((in + 255) / 256), (unsigned)((out + kMapBwdRows - 1) / kMapBwdRows)),
                                    256, 0, wstream>>>(dact[i + 1], act[i + 1], act[i], host_map_dW[i], host_map_db[i],
                                                       (int)B, in, out, leaky);
    }
    reni_film_map_bwd_dx_kernel<<<dim3((unsigned)((in + 127) / 128), (unsigned)((out + 31) / 32), (unsigned)((B + 31) / 32)),
                                  128, 0, stream>>>(dact[i + 1], act[i + 1], host_map_weights[i], dact[i], (int)B, in, out,
                                                    leaky);
  }
  {
    FilmMapBwdDzParams z{};
    z.Z = Z; z.W0 = weight0; z.dM = dM; z.dx0 = dact[0]; z.dZ = dZ; z.N = N; z.so2 = so2;
    reni_film_map_bwd_dz_kernel<<<dim3((unsigned)B, (unsigned)((3 * N + 7) / 8)), 256, 3 * N * sizeof(float), stream>>>(z);
  }
  if (side != nullptr) {  // join
    if (note(cudaEventRecord(side->join, side->stream)) != cudaSuccess) return RENI_ERR_CUDA;
    if (note(cudaStreamWaitEvent(stream, side->join, 0)) != cudaSuccess) return RENI_ERR_CUDA;
  }
  return last_err() == cudaSuccess ? RENI_OK : RENI_ERR_CUDA;
}

int32_t reni_film_prepare_maps(const reni_config_t* c, const float* film, const float* const* host_weights,
                               const float* const* host_biases, int64_t B, int64_t P, void* ws, int64_t ws_bytes,
                               int32_t flags, void* stream_) {
  if (!config_ok(c) || !c->last_layer_linear) return RENI_ERR_BAD_CONFIG;
  if (film == nullptr || host_weights == nullptr || host_biases == nullptr || ws == nullptr || B < 1 || B > 65535 ||
      !permap_ok(P))
    return RENI_ERR_BAD_ARGUMENT;
  flags |= RENI_FLAG_FILM | RENI_FLAG_FILM_PERMAP;
  const WorkspaceLayout w = make_layout(c, B, P, flags);
  if (ws_bytes < w.total || (reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return RENI_ERR_WORKSPACE;
  FilmPrepParams q{};
  const int L = c->hidden_layers;
  for (int i = 1; i <= L; ++i) {
    if (host_weights[i] == nullptr || host_biases[i] == nullptr) return RENI_ERR_BAD_ARGUMENT;
    q.w[i] = host_weights[i];
    q.b[i] = host_biases[i];
  }
  q.film = film;
  q.wf2m = at<__half>(ws, w.wf2m);
  q.wf2m_lo = at<__half>(ws, w.wf2m_lo);
  q.wb2m = at<__half>(ws, w.wb2m);
  q.wbias2m = at<__half>(ws, w.wbias2m);
  q.L = L;
  reni_film_prep_maps_kernel<<<dim3(8, (unsigned)L, (unsigned)B), 256, 0, static_cast<cudaStream_t>(stream_)>>>(q);
  return last_err() == cudaSuccess ? RENI_OK : RENI_ERR_CUDA;
}

int32_t reni_film_loss_forward_backward(const reni_config_t* c, const float* mc, const float* film, const float* D,
                                        int64_t d_bstride, const float* const* host_weights,
                                        const float* const* host_biases, int64_t B, int64_t P, const float* target,
                                        const float* sw, int64_t sw_bstride, float beta, int32_t use_cosine, float* out,
                                        float* loss_out, float* d_mc, float* d_film, float* const* host_dW,
                                        float* const* host_db, void* ws, int64_t ws_bytes, int32_t flags,
                                        void* stream_) {
  if (!config_ok(c) || !c->last_layer_linear) return RENI_ERR_BAD_CONFIG;
  if (mc == nullptr || film == nullptr || D == nullptr || host_weights == nullptr || host_biases == nullptr ||
      target == nullptr || sw == nullptr || out == nullptr || loss_out == nullptr || d_mc == nullptr ||
      d_film == nullptr || ws == nullptr || B < 1 || P < 1)
    return RENI_ERR_BAD_ARGUMENT;
  if ((flags & RENI_FLAG_NEED_DW) && (host_dW == nullptr || host_db == nullptr)) return RENI_ERR_BAD_ARGUMENT;
  flags = (flags & (RENI_FLAG_NEED_DW | RENI_FLAG_FILM_PERMAP)) | RENI_FLAG_SAVE_FOR_BACKWARD | RENI_FLAG_FILM | RENI_FLAG_LOSS;
  if ((flags & RENI_FLAG_FILM_PERMAP) && !permap_ok(P)) return RENI_ERR_BAD_ARGUMENT;
  const WorkspaceLayout w = make_layout(c, B, P, flags);
  if (ws_bytes < w.total || (reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return RENI_ERR_WORKSPACE;
  const int sms = num_sms();
  if (sms <= 0) return RENI_ERR_NO_DEVICE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  mark_phase(0, stream);
  reni_set_scale_kernel<<<1, 1, 0, stream>>>(at<float>(ws, w.scalars), 1.5f * (float)P);  // S = 3P/2, see loss finish
  mark_phase(1, stream);
  int32_t rc = launch_forward(c, w, mc, film, D, d_bstride, B, P, out, target, sw, sw_bstride, ws, flags, stream, sms);
  if (rc != RENI_OK) return rc;
  if (cudaMemsetAsync(loss_out, 0, 16, stream) != cudaSuccess) return RENI_ERR_CUDA;
  LossFinishParams f{};
  f.loss_part = at<float>(ws, w.loss_part);
  f.sw = sw;
  f.sw_bstride = sw_bstride;
  f.Z = nullptr;  // the prior term alpha * sum Z^2 belongs to the caller's per-map stage
  f.map_loss = at<float>(ws, w.map_loss);
  f.loss_out = loss_out;
  f.scalars = at<float>(ws, w.scalars);
  f.B = (int)B;
  f.P = (int)P;
  f.tiles_per_map = (int)tiles_per_map(P);
  f.nz = 0;
  f.alpha = 0.f;
  f.beta = beta;
  f.use_cos = use_cosine;
  reni_loss_finish_kernel<<<(unsigned)B, 128, 0, stream>>>(f);
  if (last_err() != cudaSuccess) return RENI_ERR_CUDA;
  FilmBackwardArgs fa{film, d_mc, d_film, host_weights, host_biases};
  return launch_backward(c, w, nullptr, D, d_bstride, nullptr, B, P, out, nullptr, target, sw, sw_bstride, 0.f, nullptr,
                         host_dW, host_db, ws, flags, stream, use_cosine, nullptr, &fa);
}

int32_t reni_selftest_umma(const void* a_img, uint32_t a_bytes, const void* b_img, uint32_t b_bytes, uint32_t a_lbo,
                           uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, uint32_t a_kstep, uint32_t b_kstep,
                           uint32_t a_mn, uint32_t b_mn, uint32_t n, uint32_t ksteps, float* d_out, void* stream) {
  if (a_img == nullptr || b_img == nullptr || d_out == nullptr) return RENI_ERR_BAD_ARGUMENT;
  if (n < 16 || n > 256 || (n % 16) != 0 || (a_bytes % 16) != 0 || (b_bytes % 16) != 0) return RENI_ERR_BAD_ARGUMENT;
  SelfTestParams p{};
  p.a_img = static_cast<const uint8_t*>(a_img);
  p.b_img = static_cast<const uint8_t*>(b_img);
  p.d_out = d_out;
  p.a_bytes = a_bytes;
  p.b_bytes = b_bytes;
  p.a_lbo = a_lbo;
  p.a_sbo = a_sbo;
  p.b_lbo = b_lbo;
  p.b_sbo = b_sbo;
  p.a_kstep = a_kstep;
  p.b_kstep = b_kstep;
  p.a_mn = a_mn;
  p.b_mn = b_mn;
  p.N = n;
  p.ksteps = ksteps;
  const size_t smem = ((a_bytes + 1023) & ~1023u) + b_bytes + 1024;
  if (smem > 220 * 1024) return RENI_ERR_BAD_ARGUMENT;
  if (note(cudaFuncSetAttribute(reni_selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) !=
      cudaSuccess)
    return RENI_ERR_CUDA;
  reni_selftest_umma_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return last_err() == cudaSuccess ? RENI_OK : RENI_ERR_CUDA;
}

int32_t reni_selftest_umma2(const void* a_img, uint32_t a_bytes, const void* b_img, uint32_t b_bytes, uint32_t a_lbo,
                            uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, uint32_t a_kstep, uint32_t b_kstep,
                            uint32_t a_mn, uint32_t b_mn, uint32_t n, uint32_t ksteps, float* d_out, void* stream) {
  if (a_img == nullptr || b_img == nullptr || d_out == nullptr) return RENI_ERR_BAD_ARGUMENT;
  if (n < 16 || n > 256 || (n % 16) != 0 || (a_bytes % 16) != 0 || (b_bytes % 16) != 0) return RENI_ERR_BAD_ARGUMENT;
  SelfTestParams p{};
  p.a_img = static_cast<const uint8_t*>(a_img);
  p.b_img = static_cast<const uint8_t*>(b_img);
  p.d_out = d_out;
  p.a_bytes = a_bytes;
  p.b_bytes = b_bytes;
  p.a_lbo = a_lbo;
  p.a_sbo = a_sbo;
  p.b_lbo = b_lbo;
  p.b_sbo = b_sbo;
  p.a_kstep = a_kstep;
  p.b_kstep = b_kstep;
  p.a_mn = a_mn;
  p.b_mn = b_mn;
  p.N = n;
  p.ksteps = ksteps;
  const size_t smem = ((a_bytes + 1023) & ~1023u) + b_bytes + 1024;
  if (smem > 220 * 1024) return RENI_ERR_BAD_ARGUMENT;
  if (note(cudaFuncSetAttribute(reni_selftest_umma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) !=
      cudaSuccess)
    return RENI_ERR_CUDA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (note(cudaLaunchKernelEx(&cfg, reni_selftest_umma2_kernel, p)) != cudaSuccess) return RENI_ERR_CUDA;
  return last_err() == cudaSuccess ? RENI_OK : RENI_ERR_CUDA;
}

int32_t reni_probe_remote_tx(const void* src, uint32_t bytes, uint32_t* result, int32_t use_tma, void* stream) {
  if (src == nullptr || result == nullptr || bytes == 0 || (bytes % 256) != 0 || bytes > 65536) return RENI_ERR_BAD_ARGUMENT;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (use_tma && !encode_rows256(&tmap, src, 2ull * bytes, bytes / 256)) return RENI_ERR_CUDA;
  if (note(cudaFuncSetAttribute(reni_probe_remote_tx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560)) !=
      cudaSuccess)
    return RENI_ERR_CUDA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = 66560;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (note(cudaLaunchKernelEx(&cfg, reni_probe_remote_tx_kernel, static_cast<const uint8_t*>(src), bytes, result, tmap,
                              (int)use_tma)) !=
      cudaSuccess)
    return RENI_ERR_CUDA;
  return last_err() == cudaSuccess ? RENI_OK : RENI_ERR_CUDA;
}

}  // extern "C"
