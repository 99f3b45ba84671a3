/* reni_b200.h -- C ABI of the B200-native RENI decoder hot path (libreni_b200.so).
 *
 * The reference (JADGardner/RENI) is pure PyTorch: it has no FFI of its own.  Each entry point
 * below replaces a span of reference Python that the drop-in module `reni_b200.models`
 * (same class names / ctor args / parameter names as src/models/RENI.py) calls instead:
 *
 *   reni_prepare_weights        <- SineLayer / nn.Linear parameter use      src/models/RENI.py:63-87,132-178
 *   reni_forward                <- InvariantRepresentation + self.net(x)    src/models/RENI.py:31-53,211-233
 *                                  (+ WeightedMSE / cosine partial sums     src/utils/loss_functions.py:6-13,25-32)
 *   reni_backward               <- autograd of the above (loss.backward())  src/lightning/RENI_module.py:105-118
 *   reni_loss_forward_backward  <- training_step: model + RENITrainLoss /   src/lightning/RENI_module.py:80-146
 *                                  RENITestLoss + backward                  src/utils/loss_functions.py:39-71
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - the caller owns every buffer including the workspace (size from reni_workspace_bytes);
 *     the library allocates no device memory and keeps no per-call state.  What it does keep: one side stream and a
 *     few events per host thread and device (created on first use, for the fork/join of independent kernels inside
 *     a step; legal under CUDA-graph capture), and the process-wide reni_debug_* hooks below, which exist for
 *     measurement and must only be changed while no call is in flight.  Compute calls are re-entrant across
 *     threads on different workspaces;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), no host synchronisation;
 *   - return value: 0 on success, negative reni_status_t otherwise (reni_strerror for text);
 *   - fp32 tensors are dense row-major exactly as the reference's torch tensors:
 *       Z (B, N, 3), D (B or 1, P, 3), out (B, P, 3), weights as nn.Linear (out, in).
 */
#ifndef RENI_B200_H_
#define RENI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RENI_ABI_VERSION 2

typedef enum {
  RENI_OK = 0,
  RENI_ERR_BAD_CONFIG = -1,    /* unsupported hidden size / layer count / equivariance */
  RENI_ERR_BAD_ARGUMENT = -2,  /* null pointer, non-positive size */
  RENI_ERR_WORKSPACE = -3,     /* workspace too small or misaligned */
  RENI_ERR_CUDA = -4,          /* a CUDA runtime call failed (launch configuration, ...) */
  RENI_ERR_NO_DEVICE = -5      /* not running on an sm_100 device */
} reni_status_t;

typedef enum { RENI_EQ_NONE = 0, RENI_EQ_SO2 = 1, RENI_EQ_SO3 = 2 } reni_equivariance_t;

/* Mirrors the constructor arguments of RENIAutoDecoder (src/models/RENI.py:91-104). */
typedef struct {
  int32_t ndims;             /* N: latent code is (N, 3)                                   */
  int32_t equivariance;      /* reni_equivariance_t                                        */
  int32_t hidden_features;   /* must be 256 on this path                                   */
  int32_t hidden_layers;     /* number of 256->256 SineLayers after the first (1..6)       */
  int32_t out_features;      /* <= 16 (3 for RGB)                                          */
  int32_t last_layer_linear; /* 1: nn.Linear output layer, 0: SineLayer                    */
  int32_t output_activation; /* 0: none, 1: tanh                                           */
  float first_omega_0;
  float hidden_omega_0;
} reni_config_t;

/* flags */
#define RENI_FLAG_SAVE_FOR_BACKWARD 1 /* forward stashes cos(a_l) (and h_l if NEED_DW) for reni_backward */
#define RENI_FLAG_NEED_DW 2           /* weight gradients wanted (otherwise latent gradients only)     */
#define RENI_FLAG_LOSS 4              /* forward also produces per-map loss sums (target / sw given)   */
#define RENI_FLAG_FILM 8              /* workspace query for the FiLM core (reni_film_forward/backward) */
#define RENI_FLAG_PREPARE_WEIGHTS 32   /* reni_loss_forward_backward builds the weight images itself (what
                                        * reni_prepare_weights does), on its side stream beside the per-map prologue */
#define RENI_FLAG_FILM_PERMAP 16      /* FiLM core on per-map weight images (reni_film_prepare_maps), see below */
#define RENI_FLAG_TILE_MAJOR_BWD 64   /* training backward of the Cond-by-Concat decoder: force the tile-major delta chain +
                                       * split-K weight-gradient GEMM (the default schedule) */
#define RENI_FLAG_FWD_SINGLE_TERM 256  /* forward hidden layers on one-term fp16 weights (fastest; radiance rel-L2 3e-4..1.3e-3) */
#define RENI_FLAG_FWD_TWO_TERM 512     /* ... on two-term weights W_hi + W_lo (the library default; x 0.6..0.8 of that error) */
#define RENI_FLAG_GRID_DIRECTIONS 1024  /* the directions are the equirectangular grid get_directions(W) (src/utils/utils.py:46-65)
                                        * with P = W * W / 2: computed in the kernels from the pixel index, D may be NULL */
#define RENI_FLAG_GRID_SINEWEIGHT 2048  /* the sine weights are get_sineweight(W) (utils.py:68-78) times a binary mask
                                        * (RENI_module.py:92-94): computed in the kernels; the `sw` argument is then NOT a float
                                        * tensor but NULL (no mask) or a `const uint32_t*` of P bits, bit p = pixel p is kept */
#define RENI_FLAG_LAYER_MAJOR_BWD 128 /* ... force the layer-major schedule (one launch per hidden layer computes the delta chain
                                       * and the weight gradients in one pass over the stash; same results up to summation order) */

int32_t reni_abi_version(void);
const char* reni_strerror(int32_t code);

/* in_features of the first layer for a config (src/models/RENI.py:118-126). */
int64_t reni_in_features(const reni_config_t* cfg);

/* Bytes of workspace needed for a batch of B maps x P directions under `flags`. */
int64_t reni_workspace_bytes(const reni_config_t* cfg, int64_t B, int64_t P, int32_t flags);

/* Width of one entry of the training forward's phase stash in this build (12, or 16 with -DRENI_PHASE_BITS=16): the
 * stash holds one such number per hidden pre-activation and is what reni_workspace_bytes is dominated by; callers that
 * account for the step's HBM traffic (bench.py) ask here. */
int32_t reni_phase_bits(void);

/* Convert the decoder parameters into the fp16 operand images the kernels stream.
 * weights[i] / biases[i], i = 0 .. hidden_layers + 1, are HOST arrays of DEVICE pointers in
 * the order of net.0.linear, ..., net.L.linear, net.(L+1) (state_dict order of the reference).
 * Must be called again whenever the parameters change (i.e. once per optimiser step). */
int32_t reni_prepare_weights(const reni_config_t* cfg, const float* const* host_weights,
                             const float* const* host_biases, void* workspace, int64_t workspace_bytes,
                             void* stream);

/* Decoder forward: out[b,p,:] = net(encoding(Z[b], D[b,p])).
 *   d_batch_stride : elements between consecutive maps in D (0 = one direction grid for all maps)
 *   target, sw     : only with RENI_FLAG_LOSS; sw_batch_stride like d_batch_stride
 *   weights0/bias0 : first-layer parameters (fp32, used by the per-map prologue)
 * With RENI_FLAG_LOSS the per-map sums needed by the losses are left in the workspace for
 * reni_loss_finish / reni_backward. */
int32_t reni_forward(const reni_config_t* cfg, const float* Z, const float* D, int64_t d_batch_stride,
                     const float* weight0, const float* bias0, int64_t B, int64_t P, float* out,
                     const float* target, const float* sw, int64_t sw_batch_stride, void* workspace,
                     int64_t workspace_bytes, int32_t flags, void* stream);

/* Backward of reni_forward for an external gradient (autograd contract, RENI_module.py:105 +
 * loss.backward()).  Must follow a reni_forward call made with RENI_FLAG_SAVE_FOR_BACKWARD (and
 * RENI_FLAG_NEED_DW if weight gradients are wanted) on the same workspace, B, P and stream order.
 *   out, grad_out : (B, P, 3)   forward result and dLoss/dout
 *   dZ            : (B, N, 3)   written
 *   host_dW/db    : HOST arrays of hidden_layers + 2 DEVICE pointers shaped like the parameters; the
 *                   gradients are ACCUMULATED (+=) into them (caller zero-fills); NULL without NEED_DW
 *   host_weights  : as in reni_prepare_weights (only weights[0] is read, fp32) */
int32_t reni_backward(const reni_config_t* cfg, const float* Z, const float* D, int64_t d_batch_stride,
                      const float* const* host_weights, int64_t B, int64_t P, const float* out,
                      const float* grad_out, float* dZ, float* const* host_dW, float* const* host_db,
                      void* workspace, int64_t workspace_bytes, int32_t flags, void* stream);

/* One fused training / latent-optimisation step (training_step of the reference without the
 * optimiser): forward, loss, backward.
 *   loss = sum_b mean_{p,c}((o-t)^2 sw) + alpha * sum Z^2 + beta * WeightedCosineSimilarity(o,t,sw)
 *   (RENITrainLoss: alpha = beta = 0, use_cosine = 0;  RENITestLoss: loss_functions.py:60-71)
 *   loss_out : 4 floats {loss, mse, prior, cosine} (device), overwritten
 *   dZ       : (B, N, 3) written, includes the 2*alpha*Z prior gradient
 *   host_dW/db as in reni_backward (NULL for latent-only optimisation, flags without NEED_DW)
 * reni_prepare_weights must have been called on this workspace for the current parameters. */
int32_t reni_loss_forward_backward(const reni_config_t* cfg, const float* Z, const float* D,
                                   int64_t d_batch_stride, const float* const* host_weights,
                                   const float* const* host_biases, int64_t B, int64_t P, const float* target,
                                   const float* sw, int64_t sw_batch_stride, float alpha, float beta,
                                   int32_t use_cosine, float* out, float* loss_out, float* dZ,
                                   float* const* host_dW, float* const* host_db, void* workspace,
                                   int64_t workspace_bytes, int32_t flags, void* stream);

/* FiLM-conditioned decoder core (RENIAutoDecoderFiLM, src/models/RENI.py:515-524,527-678):
 *     h_0 = sin([f | 1] . mc[b]),   h_l = sin(freq_l[b] * (W_l h_{l-1} + b_l) + phase_l[b]),  l = 1..L,
 *     out = act(W_out h_L + b_out)
 * replaces FiLMLayer.forward x L + final_layer (forward_with_frequencies_phase_shifts, RENI.py:666-678).  Everything
 * that is constant per map stays with the caller, which can differentiate it with autograd on (B, .) tensors:
 *   mc   (B, 5, 256) : rows 0..3 = freq_0 * M_b (first FiLM layer hoisted onto the direction features
 *                      f = [dx, dz, |d_xz|, dy] for SO2 (RENI.py:418-436) or f = [dx, dy, dz, 0] for SO3 (:405-415)),
 *                      row 4 = freq_0 * b_0 + phase_0
 *   film (B, L, 2, 256) : freq_l = 15 * raw + 30 and phase_l of the mapping network (RENI.py:481-512,667), l = 1..L
 * cfg: hidden_layers = L = siren_hidden_layers - 1, hidden_omega_0 = first_omega_0 = 1, last_layer_linear = 1,
 * equivariance SO2 / SO3, output_activation none / tanh ("exp" is applied by the caller).  reni_prepare_weights must
 * have been called with that cfg (weights[0] / biases[0] are ignored by it); workspace size = reni_workspace_bytes
 * with RENI_FLAG_FILM [| RENI_FLAG_SAVE_FOR_BACKWARD]. */
int32_t reni_film_forward(const reni_config_t* cfg, const float* mc, const float* film, const float* D,
                          int64_t d_batch_stride, int64_t B, int64_t P, float* out, void* workspace,
                          int64_t workspace_bytes, int32_t flags, void* stream);

/* Backward of reni_film_forward (made with RENI_FLAG_SAVE_FOR_BACKWARD on the same workspace):
 *   d_mc (B, 5, 256), d_film (B, L, 2, 256) : written -- gradients w.r.t. mc and film
 *   host_weights / host_biases : fp32 parameters as in reni_prepare_weights ([1..L] are read)
 *   host_dW / host_db          : with RENI_FLAG_NEED_DW, entries [1..L+1] are ACCUMULATED into ([0] is unused: the
 *                                first layer is differentiated by the caller through d_mc); NULL otherwise. */
int32_t reni_film_backward(const reni_config_t* cfg, const float* film, const float* D, int64_t d_batch_stride,
                           const float* const* host_weights, const float* const* host_biases, int64_t B, int64_t P,
                           const float* out, const float* grad_out, float* d_mc, float* d_film,
                           float* const* host_dW, float* const* host_db, void* workspace, int64_t workspace_bytes,
                           int32_t flags, void* stream);

/* FiLM core on PER-MAP weight images (RENI_FLAG_FILM_PERMAP in the flags of the workspace query and of every
 * reni_film_* compute call of the step).  sin(freq_l[b] (W_l h + b_l) + phase_l[b]) = sin((diag(freq_l[b]) W_l) h +
 * (freq_l[b] b_l + phase_l[b])): with the modulation folded into fp16 weight / bias images of each map, built here from
 * the fp32 parameters (one rounding, as for the shared images), the FiLM layers run on the kernels of the
 * Cond-by-Concat decoder -- plain sin epilogue with the bias on the tensor core, delta stash written by bulk copies --
 * and the backward's delta_l * freq_l becomes part of the backward image.  Costs B * L * 264 KB of workspace and one
 * launch; requires (P / 128) % 4 == 0 with P % 128 == 0 (the four tiles of a CTA pair's unit share one map), else
 * RENI_ERR_BAD_ARGUMENT.  Call after reni_prepare_weights and before reni_film_forward /
 * reni_film_loss_forward_backward, on the same stream. */
int32_t reni_film_prepare_maps(const reni_config_t* cfg, const float* film, const float* const* host_weights,
                               const float* const* host_biases, int64_t B, int64_t P, void* workspace,
                               int64_t workspace_bytes, int32_t flags, void* stream);

/* FiLM per-map stage, forward only (no-grad decoding of a few latents): mapping-network input (RENI.py:405-452),
 * mapping network (Linear / LeakyReLU(0.2) stack, RENI.py:481-512), freq = 15 raw + 30 (:667) and the hoisted first FiLM
 * layer -> mc (B, 5, 256) and film (B, L, 2, 256) for reni_film_forward, in 2 + n_linears launches.
 *   weight0 / bias0            : net[0].layer parameters (256, 2 + N) for SO2, (256, N) for SO3
 *   host_map_weights / biases  : HOST arrays of n_linears (<= 8) DEVICE pointers, mapping_network.network[2 i]
 *   host_map_dims              : HOST array of n_linears + 1 sizes: mapping input, then each linear's out features
 *   scratch                    : caller-owned device buffer of reni_film_map_scratch_bytes(...) bytes
 * Differentiated calls keep this stage with the caller (batched GEMMs + autograd). */
int64_t reni_film_map_scratch_bytes(const int32_t* host_map_dims, int32_t n_linears, int64_t B);
int32_t reni_film_map_forward(const reni_config_t* cfg, const float* Z, const float* weight0, const float* bias0,
                              const float* const* host_map_weights, const float* const* host_map_biases,
                              const int32_t* host_map_dims, int32_t n_linears, int64_t B, float* mc, float* film,
                              void* scratch, int64_t scratch_bytes, void* stream);

/* FiLM per-map stage for TRAINING / latent fitting: the same forward keeping every activation of the mapping network
 * (`acts`, reni_film_map_acts_bytes bytes, caller-owned, must stay untouched until the backward), and its hand-derived
 * backward from the core's d_mc / d_film (replaces ~120 autograd launches of (B, .) torch ops by 6 + 3 + 2 n_linears):
 *   dZ (B, N, 3)                 : written
 *   dW0, db0, host_map_dW/db[i]  : ACCUMULATED (+=) gradients of net[0].layer and mapping_network.network[2 i];
 *                                  all NULL for a frozen decoder (latent fitting: only dZ is produced)
 *   scratch                      : reni_film_map_acts_bytes(...) + B * 4 * 256 * 4 bytes
 * Reference: RENI.py:405-452 (mapping input), :481-512 (mapping network), :515-524 + :666-678 (first FiLM layer). */
int64_t reni_film_map_acts_bytes(const int32_t* host_map_dims, int32_t n_linears, int64_t B);
int32_t reni_film_map_forward_train(const reni_config_t* cfg, const float* Z, const float* weight0, const float* bias0,
                                    const float* const* host_map_weights, const float* const* host_map_biases,
                                    const int32_t* host_map_dims, int32_t n_linears, int64_t B, float* mc, float* film,
                                    void* acts, int64_t acts_bytes, void* stream);
int32_t reni_film_map_backward(const reni_config_t* cfg, const float* Z, const float* weight0, const float* bias0,
                               const float* const* host_map_weights, const int32_t* host_map_dims, int32_t n_linears,
                               int64_t B, const void* acts, const float* d_mc, const float* d_film, float* dZ,
                               float* dW0, float* db0, float* const* host_map_dW, float* const* host_map_db,
                               void* scratch, int64_t scratch_bytes, void* stream);

/* One fused FiLM training / latent-fit step around the core: reni_film_forward with the loss sums, the loss reduction
 * and reni_film_backward with the loss gradient formed in the kernel (replaces the criterion + loss.backward() of
 * training_step, src/lightning/RENI_module.py:105-134, for a FiLM decoder).
 *   loss = sum_b mean_{p,c}((o-t)^2 sw) + beta * WeightedCosineSimilarity(o,t,sw)   (loss_functions.py:6-32)
 *   loss_out : 4 floats {loss, mse, 0, cosine}; the prior alpha * sum Z^2 and the KLD are per-map terms of the caller
 *   d_mc, d_film, host_dW/db as in reni_film_backward.  output_activation none / tanh only. */
int32_t reni_film_loss_forward_backward(const reni_config_t* cfg, const float* mc, const float* film, const float* D,
                                        int64_t d_batch_stride, const float* const* host_weights,
                                        const float* const* host_biases, int64_t B, int64_t P, const float* target,
                                        const float* sw, int64_t sw_batch_stride, float beta, int32_t use_cosine,
                                        float* out, float* loss_out, float* d_mc, float* d_film,
                                        float* const* host_dW, float* const* host_db, void* workspace,
                                        int64_t workspace_bytes, int32_t flags, void* stream);

/* Fused Adam step over a list of fp32 segments: replaces torch.optim.Adam(params, lr).step() as the reference's
 * configure_optimizers builds it (src/lightning/RENI_module.py:185-192: default betas (0.9, 0.999), eps 1e-8, no
 * weight decay, dense over every parameter including the whole latent table).  One launch per 24 segments.
 *   host_segments : HOST array of nseg descriptors of DEVICE buffers (param and both moments are updated in place)
 *   step          : DEVICE int32, number of completed steps; read as t = *step + 1 and incremented on the stream */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
} reni_adam_segment_t;
int32_t reni_adam_step(const reni_adam_segment_t* host_segments, int32_t nseg, int32_t* step, double lr, double beta1,
                       double beta2, double eps, void* stream);

/* Variational auto-decoder latents (RENIVariationalAutoDecoder.sample_latent, src/models/RENI.py:329-335; KLD,
 * src/utils/loss_functions.py:16-22; RENIVADTrainLoss as training_step builds it, src/lightning/RENI_module.py:312-315).
 *   reni_vad_sample   : Z[b] = mu[idx[b]] + eps[b] * exp(0.5 * log_var[idx[b]]); eps (B, N, 3) ~ N(0,1) from the caller
 *   reni_vad_backward : dmu / dlog_var (whole tables, ACCUMULATED) from dLoss/dZ plus the KLD term's own gradient, all
 *                       times grad_scale (1 / world size under data parallelism: DDP averages the replicated tables'
 *                       gradients); kld_out (1 float, ACCUMULATED) += kld_weight / Z_dims * sum_b KLD_b
 *   idx : (B) int64 rows of the tables; nz = 3 N floats per row */
int32_t reni_vad_sample(const float* mu, const float* log_var, const int64_t* idx, const float* eps, int64_t B,
                        int64_t nz, float* Z, void* stream);
int32_t reni_vad_backward(const float* mu, const float* log_var, const int64_t* idx, const float* eps, const float* dZ,
                          int64_t B, int64_t nz, float kld_weight_over_zdims, float grad_scale, float* dmu,
                          float* dlog_var, float* kld_out, void* stream);

/* Blinn-Phong shading of a surface by every texel of the environment map -- what the FIT_INVERSE task does with the
 * decoder output (blinn_phong_shading_env_map, src/utils/pytorch3d_envmap_shader.py:85-119, reached from
 * RENI_module.get_render, :386-396): with unit surface normals n_p and unit view directions v_p per render pixel,
 *   colors[b, p, :] = sum_j ( kd * clamp(n_p . l_j) + c * ks * clamp(n_p . normalize(v_p + l_j))^shininess ) * light[b, j, :]
 *   c = (shininess + 2) / (4 (2 - exp(-shininess / 2)));   light = environment map x sine weight (EnvironmentMap, :33-44)
 * and its adjoint with respect to the light colours, d_light[b, j, :] = sum_p (same weight) * grad_colors[b, p, :].
 * The (B, H, W, J[, 3]) tensors the reference materialises are never formed.
 *   normals, view_dirs : (n_pix, 3);  D : (B or 1, J, 3), d_batch_stride 0 = one grid for all maps
 *   light / d_light : (B, J, 3);  colors / grad_colors : (B, n_pix, 3), all overwritten where written */
int32_t reni_envmap_shade_forward(const float* normals, const float* view_dirs, int64_t n_pix, const float* D,
                                  int64_t d_batch_stride, const float* light, int64_t B, int64_t J, float kd, float ks,
                                  float shininess, float* colors, void* stream);
int32_t reni_envmap_shade_backward(const float* normals, const float* view_dirs, int64_t n_pix, const float* D,
                                   int64_t d_batch_stride, const float* grad_colors, int64_t B, int64_t J, float kd,
                                   float ks, float shininess, float* d_light, void* stream);

/* In-place mean (or scaled sum) of one fp32 buffer over the W ranks of a node, through NVLink peer memory: the one
 * exchange step of data-parallel training (Lightning's DDPStrategy averages every gradient, run.py:97), as a plain
 * kernel that can be captured in the step's CUDA graph directly behind the gradient kernels.
 *   dev_buf_ptrs  : DEVICE array of W pointers to the ranks' buffers (the same symmetric allocation on every rank,
 *                   e.g. torch.distributed._symmetric_memory's buffer_ptrs_dev)
 *   dev_flag_ptrs : DEVICE array of W pointers to the ranks' flag blocks, reni_allreduce_flag_bytes() each, zeroed once
 *   multicast_ptr : multicast address aliasing the W buffers (NVSwitch, multimem.ld_reduce / multimem.st), or NULL for
 *                   the two-shot exchange over the peer pointers
 *   numel         : fp32 elements, a multiple of 4;  scale: 1 / W for the mean
 *   epoch         : DEVICE block of reni_allreduce_flag_bytes() / 64 words, zeroed once: per-block call counters kept by
 *                   the kernel itself (order the flag values across calls and graph replays)
 *   status        : DEVICE word, set to 1 if a rank gave up waiting for a peer (bounded spin; the result is then wrong)
 * Every rank must make the same sequence of calls. */
int64_t reni_allreduce_flag_bytes(void);
int32_t reni_allreduce(void* dev_buf_ptrs, void* dev_flag_ptrs, void* multicast_ptr, int64_t numel, int32_t rank,
                       int32_t world, float scale, uint32_t* epoch, uint32_t* status, void* stream);

/* Debug / measurement hook: register up to 16 CUDA events (cudaEvent_t handles, HOST array) that the
 * calling thread's subsequent reni_forward / reni_backward / reni_loss_forward_backward calls record on
 * their stream between kernels: [0] start, [1] after the per-map prologue, [2] after the forward kernel,
 * [3] after the loss reduction, [4] after the delta-chain kernel, [5] after the weight-gradient GEMM,
 * [6] end of the step.  n = 0 clears.  Process-wide (autograd calls the backward on its own thread): set and
 * clear it while no library call is in flight; besides per-thread side streams this is the only state kept. */
int32_t reni_debug_set_phase_events(void* const* host_events, int32_t n);

/* Debug hook: device buffer of 3 x 4096 uint64 into which CTA 0 of the calling thread's next forward kernels
 * writes a clock64 timeline (event code << 48 | clock) per role: [0] MMA issuer, [1], [2] epilogue groups.
 * NULL clears.  Used by tools/trace_fwd.py to read the pipeline's critical path. */
int32_t reni_debug_set_trace(void* device_buffer);

/* Measurement / tuning hook for the training backward's OVERLAP mode, in which the weight-gradient kernel runs beside
 * the delta-chain kernel on `dw_ctas` of the SMs (side stream) and takes each tile's stash blocks out of L2 as the
 * chain finishes them (per-tile ready counters in the workspace) instead of re-reading them from HBM afterwards.
 * dw_ctas = 0 (the default): overlap off, kernels back to back; -1: a 36 % share of the SMs; else the SM count
 * (rounded down to even); out_ctas = how many of them take the output-layer job (0: round robin over jobs).
 * Process-wide; results are identical up to the order of the fp32 gradient reductions.  Measured slower than the
 * back-to-back kernels at cfg 2, hence off by default. */
int32_t reni_debug_set_overlap(int32_t dw_ctas, int32_t out_ctas);

/* Debug hook: cudaGetErrorString of the last CUDA runtime error this thread saw inside the library
 * ("no error" if none); lets a caller turn RENI_ERR_CUDA into a readable message. */
const char* reni_debug_last_cuda_error(void);

/* Debug / test hook: one 128 x N x (16*ksteps) tcgen05.mma with caller-supplied operand images
 * and descriptor fields; writes the fp32 accumulator tile (128 x N, row-major) to d_out. */
int32_t reni_selftest_umma(const void* a_img, uint32_t a_bytes, const void* b_img, uint32_t b_bytes,
                           uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, uint32_t a_kstep,
                           uint32_t b_kstep, uint32_t a_mn_major, uint32_t b_mn_major, uint32_t n,
                           uint32_t ksteps, float* d_out, void* stream);

/* Same, for a CTA pair: one cluster of two CTAs, tcgen05.mma.cta_group::2 with M = 256.  a_img holds two
 * operand images back to back (rows 0..127, rows 128..255), b_img the two N/2-row halves of B; the descriptor
 * fields apply to each CTA's own image.  d_out is 256 x N, row-major. */
int32_t reni_selftest_umma2(const void* a_img, uint32_t a_bytes, const void* b_img, uint32_t b_bytes,
                           uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, uint32_t a_kstep,
                           uint32_t b_kstep, uint32_t a_mn_major, uint32_t b_mn_major, uint32_t n,
                           uint32_t ksteps, float* d_out, void* stream);

/* Debug probe: does a bulk copy into a CTA's own shared memory complete on an mbarrier of its cluster peer?
 * result[0] = 1 if the leader's barrier completed, result[1], result[2] = byte sums each CTA read back.
 * use_tma = 0: plain cp.async.bulk (measured: completes on the destination CTA only); 1: TMA tile load with
 * .cta_group::2 through a tensor map (completes on the leader's barrier). */
int32_t reni_probe_remote_tx(const void* src, uint32_t bytes, uint32_t* result, int32_t use_tma, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RENI_B200_H_ */
