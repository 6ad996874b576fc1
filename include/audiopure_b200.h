/* audiopure_b200 -- C ABI of the B200 (sm_100a) purification hot path.
 *
 * The reference (cychomatica/AudioPure) is pure Python: its "plugin interface" for this path is the
 * duck-typed defender / transform / certifier objects of acoustic_system.py:5-9,35-51 and
 * robustness_eval/certified_robust.py:8-14.  There is no native FFI to replace, so this header declares
 * the entry points the Python host side (audiopure_b200/*.py, ctypes) binds, one per reference call site,
 * each citing the reference function it stands in for.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; ap_last_error() gives the message of the
 *     calling thread's last failure;
 *   - all tensor arguments are raw DEVICE pointers (fp32 unless stated), caller-owned; nothing is
 *     allocated behind the caller's back except the handle's own small tables; work is enqueued on
 *     `stream` (a cudaStream_t passed as void*) with no host synchronisation;
 *   - waveforms are [B][L] row-major (the reference's (B,1,L) with the unit channel squeezed);
 *   - one handle per device; a handle is not thread-safe.
 */
#ifndef AUDIOPURE_B200_H
#define AUDIOPURE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AP_ABI_VERSION 6

typedef struct ap_net ap_net;   /* the packed DiffWave epsilon-network + its diffusion schedule */
typedef struct ap_comm ap_comm; /* an NCCL communicator for the vote-count all-reduce           */

/* Precision of the residual-stack GEMMs (north_star: "BF16 tensor-core mode" / "TF32 mode").  Default: bf16
 * operands, fp32 accumulation, residual stream / gates / skip sum stored as bf16.  AP_FLAG_TF32: the same kernels
 * instantiated for tcgen05 kind::tf32 -- the residual stream, gates and skip sum stay fp32 in HBM and shared
 * memory, operands are rounded to tf32 (10 mantissa bits) only on their way into the tensor core, tanh/sigmoid are
 * fp32-accurate.  About 2.2x slower and 2x the workspace; for parity work against the fp32 reference.
 * With this flag the GEMM weights (w1, w2, ws, wf) are fp32 words already rounded to tf32.                    */
#define AP_FLAG_TF32 1u

/* configs/config.json "wavenet_config"/"diffusion_config" (diffwave_ddpm.py:395-402) + schedule tables.
 * The kernels support res_channels == skip_channels == 256, in/out channels == 1.               */
typedef struct ap_config {
  int32_t num_res_layers; /* 36 */
  int32_t dilation_cycle; /* 12 -> dilation 2^(n mod 12), WaveNet.py:116 */
  int32_t T;              /* 200 */
  int32_t max_chunk;      /* clips processed per pass (bounds the workspace); 0 -> 64 (bf16) / 32 (tf32) */
  uint32_t flags;         /* 0 or AP_FLAG_TF32; any other bit is rejected */
  /* HOST pointers to fp32[T] tables, copied at create.  Built by the caller with the reference's own
   * expressions so they are bit-equal to it: util.py:111-123 and diffwave_sde.py:57-58.          */
  const float* alpha;
  const float* alpha_bar;
  const float* sigma;
  const float* sde_beta;          /* RevVPSDE.discrete_betas  */
  const float* sde_alphas_cumprod; /* RevVPSDE.alphas_cumprod */
} ap_config;

/* Packed weights, DEVICE pointers that must outlive the handle (WaveNet_Speech_Commands.pack_weights in audiopure_b200/wavenet.py builds them from
 * a reference-layout state dict: weight-norm folded once, WaveNet.py:28,67,72).                   */
typedef struct ap_weights {
  const void* w1;     /* bf16 (fp32 with AP_FLAG_TF32, also w2/ws/wf) [layers][512][768]: dilated conv, rows gate-interleaved (sigmoid rows x 1/2), K = tap*256 + cin */
  const float* b1;    /* f32  [layers][512]: its bias, same row order and 1/2 factors                       */
  const void* w2;     /* bf16 [layers][256][256]: 1/2 sqrt(.5) * res_conv (the kernels keep 2 x gate)        */
  const float* c2;    /* f32  [T][layers][256]: sqrt(.5)*b_res[n] + fc_t[n+1](emb(t))  (0 shift for last)  */
  const float* part0; /* f32  [T][256]: fc_t[0](emb(t))                                                    */
  const float* w0;    /* f32  [256]: init conv weight (1 -> 256)                                           */
  const float* b0;    /* f32  [256]                                                                        */
  const void* ws;     /* bf16 [256][layers*256]: 1/2 sqrt(1/layers) * skip_conv of every layer, concatenated K */
  const float* bs;    /* f32  [256]: sqrt(1/layers) * sum_n b_skip[n]                                      */
  const void* wf;     /* bf16 [256][256]: final_conv[0]                                                    */
  const float* bf;    /* f32  [256]                                                                        */
  const float* wo;    /* f32  [256]: final_conv[2] (ZeroConv1d) weight                                     */
  float bo;           /*            its bias                                                               */
} ap_weights;

const char* ap_last_error(void);
int ap_abi_version(void);

/* create_diffwave_model (diffwave_ddpm.py:395-411): bind packed weights + schedule to the current device. */
int ap_create(const ap_config* cfg, const ap_weights* w, ap_net** out);
void ap_destroy(ap_net* net);

/* Bytes of caller-provided workspace any call below needs for a batch of B clips of L samples. */
size_t ap_workspace_bytes(const ap_net* net, int B, int L);

/* WaveNet_Speech_Commands.forward((x, t*ones)) (WaveNet.py:164-172) == DiffWave.compute_eps_t
 * (diffwave_ddpm.py:166-172): eps_out[B][L] = eps_theta(x, t), t in [0, T).                              */
int ap_eps(ap_net* net, const float* x, int B, int L, int t, float* eps_out, void* ws, size_t ws_bytes,
           void* stream);

/* One network evaluation with the affine update fused into the output head:
 *   x_out = ca * x_in + cb * eps_theta(x_in, t) + cc * z
 * (DiffWave.compute_coefficients + the update at diffwave_ddpm.py:99-102,159 expressed as coefficients).
 * z: injected noise [B][L], or NULL -> in-kernel Philox keyed on (seed, stream_id, clip_offset*L + index). */
int ap_step(ap_net* net, const float* x_in, float* x_out, int B, int L, int t, float ca, float cb, float cc,
            const float* z, uint64_t seed, uint32_t stream_id, int64_t clip_offset, void* ws, size_t ws_bytes,
            void* stream);

/* DiffWave.forward (diffwave_ddpm.py:36-47): diffuse to t_star (:49-73) then t_star reverse steps (:75-104).
 * z: [t_star][B][L] in the reference's draw order (z[0] diffusion, z[1+i] reverse step i), or NULL -> Philox.
 * clip_offset: index of x_in's first clip in the caller's global batch (keeps Philox noise independent of
 * how the batch is sharded over GPUs).                                                                   */
int ap_ddpm_purify(ap_net* net, const float* x_in, float* x_out, int B, int L, int t_star, const float* z,
                   uint64_t seed, int64_t clip_offset, void* ws, size_t ws_bytes, void* stream);

/* RevDiffWave.audio_editing_sample, one sample_step (diffwave_sde.py:183-205): diffuse with the cumprod
 * alpha-bar (:190-191), then t Euler-Maruyama steps of the reverse VP-SDE at indices k = t-1..0
 * (RevVPSDE.f/.g, :73-134).  z: [t+1][B][L] (z[0] diffusion noise e, z[1+i] increment of step i; the last
 * one has zero weight), or NULL -> Philox.                                                               */
int ap_sde_purify(ap_net* net, const float* x_in, float* x_out, int B, int L, int t, const float* z,
                  uint64_t seed, int64_t clip_offset, void* ws, size_t ws_bytes, void* stream);

/* DiffWave.one_shot_denoise (diffwave_ddpm.py:174-182,195-205) at t = reverse_timestep - 1. */
int ap_one_shot(ap_net* net, const float* x_in, float* x_out, int B, int L, int reverse_timestep, void* ws,
                size_t ws_bytes, void* stream);

/* torchaudio MelSpectrogram(n_fft=2048, hop=512, n_mels, slaney/slaney, pad 'constant') + AmplitudeToDB
 * ('power'), the transform built at adaptive_attack_eval.py:83-85.  out: [B][n_mels][1 + L/512].
 * Tables are DEVICE pointers owned by the caller (audiopure_b200/transforms.py builds them).             */
typedef struct ap_mel_tables {
  const void* twiddles;    /* float2[1024]: exp(-2 pi i k / 2048) */
  const int32_t* fb_start; /* [n_mels] */
  const int32_t* fb_len;   /* [n_mels] */
  const int32_t* fb_off;   /* [n_mels] */
  const float* fb_w;       /* concatenated non-zero filterbank weights */
  int32_t n_mels;
  int32_t fb_nnz;          /* number of floats in fb_w (<= 2304: every bin belongs to at most two triangular filters) */
} ap_mel_tables;
int ap_logmel(const float* x, int B, int L, float* out, const ap_mel_tables* tabs, void* stream);

/* Gradient of ap_logmel with respect to the waveform, for attacks that back-propagate through AcousticSystem
 * (white_box_attack.py:392,437-439): grad_x[B][L] = d<grad_out, logmel(x)>/dx, grad_out: [B][n_mels][1 + L/512].
 * L is limited by the per-clip shared-memory accumulator (L <= 40000).                                     */
int ap_logmel_backward(const float* x, int B, int L, const float* grad_out, float* grad_x,
                       const ap_mel_tables* tabs, void* stream);

/* RobustCertificate.smooth_predict's input construction (certified_robust.py:46-54):
 *   out[j][l] = scale * (x[l] + sigma * z_j[l]),  j in [0, n_draws)
 * z: injected [n_draws][L], or NULL -> Philox keyed on (seed, clip, first_draw + j, l).                  */
int ap_smooth_inputs(const float* x, int L, int n_draws, float sigma, float scale, const float* z,
                     uint64_t seed, uint32_t clip, int64_t first_draw, float* out, void* stream);

/* certified_robust.py:58-67: counts[c] += #{rows : argmax_k logits[row][k] == c}; counts is int64[K] and
 * is accumulated into (zero it first).                                                                   */
int ap_vote_counts(const float* logits, int rows, int K, int64_t* counts, void* stream);

/* The same two steps for a certify call over SEVERAL clips (certified_robust.py:81-96), so that every launch is a
 * full batch whatever n_0, n and the number of ranks are.  The call's work list is the clip-major flattening of
 * (clip, draw), draw in [0, per_clip): row r of a launch is item flat = flat0 + r, clip = flat / per_clip,
 * draw = first_draw + flat % per_clip.
 *   ap_smooth_inputs_batch: out[r][l] = scale * (x[clip][l] + sigma * z); z injected ([clips][per_clip][L], indexed
 *     by flat) or NULL -> Philox keyed on (seed, clip_key0 + clip, draw, l) -- the draws do not depend on how the
 *     list is batched or sharded.
 *   ap_vote_counts_batch: counts[sel][clip][argmax] += 1 with sel = (flat % per_clip >= n_split); counts is
 *     int64[2][n_clips][K] (selection pass n_0 | estimation pass n of certify), accumulated into.            */
int ap_smooth_inputs_batch(const float* x, int L, int n_rows, int64_t flat0, int64_t per_clip, int64_t first_draw,
                           float sigma, float scale, const float* z, uint64_t seed, uint32_t clip_key0, float* out,
                           void* stream);
int ap_vote_counts_batch(const float* logits, int rows, int K, int64_t flat0, int64_t per_clip, int64_t n_split,
                         int n_clips, int64_t* counts, void* stream);

/* NES black-box gradient estimation with antithetic sampling (robustness_eval/_NES.py:15-55), one draw batch of S
 * samples (S even) per audio:
 *   ap_nes_inputs: out[a][lead + j][l] = x[a][l] + sigma * noise_j[l], noise_j = +z_j (j < S/2), -z_{j-S/2} otherwise
 *     (_NES.py:19-24); lead = 1 prepends the clean audio (first draw batch, _NES.py:22-23).  out: [audios][lead+S][L].
 *     z: injected [audios][S/2][L] or NULL -> Philox keyed on (seed, audio_key0 + a, draw0 + j, l).
 *   ap_nes_grad: grad[a][l] += grad_scale * sum_j loss[a][loss_off + j] * noise_j[l]   (_NES.py:44-48,52), the noise
 *     re-generated from the same keys (or read from the same z) instead of being kept in HBM.                   */
int ap_nes_inputs(const float* x, int audios, int L, int S, int lead, float sigma, const float* z, uint64_t seed,
                  uint32_t audio_key0, int64_t draw0, float* out, void* stream);
int ap_nes_grad(const float* loss, int loss_stride, int loss_off, int audios, int L, int S, float grad_scale,
                const float* z, uint64_t seed, uint32_t audio_key0, int64_t draw0, float* grad, void* stream);

/* Consumer-side epilogue of the ResNeXt bottleneck after batch-norm folding (resnext.py:56-64):
 *   y[r][c] = relu?(y[r][c] + bias[c] (+ residual[r][c]))      in place, y/residual channels-last bf16 [rows][C],
 * bias f32[C], C a multiple of 8.  One pass instead of separate bias-add, residual-add and clamp launches.      */
int ap_bias_act_nhwc_bf16(void* y, const float* bias, const void* residual, int64_t rows, int C, int relu,
                          void* stream);

/* Sharded certification: sum the int64 vote counts of all ranks (NCCL, loaded with dlopen).  The unique id
 * is created on rank 0 and handed to the other ranks by the caller (torch.distributed broadcast).        */
#define AP_COMM_ID_BYTES 128
int ap_comm_unique_id(uint8_t id[AP_COMM_ID_BYTES]);
int ap_comm_init(int rank, int world, const uint8_t id[AP_COMM_ID_BYTES], ap_comm** out);
int ap_allreduce_counts(ap_comm* comm, int64_t* counts, size_t n, void* stream);
void ap_comm_destroy(ap_comm* comm);

/* Measurement hook (bench.py's roofline line): while enabled, every prologue / residual-layer / tail kernel
 * launch is bracketed by CUDA events recorded on the launching stream.  ap_profile_read waits for them, adds
 * the per-kind device time (ms) and launch counts accumulated since the last read into ms_sum[3] /
 * launches[3] (index 0: residual-layer kernel, 1: tail kernel, 2: prologue kernel), and resets.            */
int ap_profile_enable(ap_net* net, int enable);
int ap_profile_read(ap_net* net, double ms_sum[3], int64_t launches[3]);

/* Bring-up check of the tcgen05/TMA conventions: D[128][256] (f32) = A[128][K] * B[256][K]^T, bf16 inputs
 * (K a multiple of 64), or fp32 inputs through kind::tf32 (K a multiple of 32).                            */
int ap_debug_gemm(const void* a_bf16, const void* b_bf16, float* d, int K, void* stream);
int ap_debug_gemm_tf32(const float* a_f32, const float* b_f32, float* d, int K, void* stream);

/* Which precision a handle was created with, and (tf32) what ap_create's probe found: round_bias is 0x1000 if the
 * tensor core drops the low 13 mantissa bits of an fp32 operand (the kernels then pre-add half a tf32 ulp), 0 if
 * it rounds to nearest itself.                                                                               */
int ap_precision(const ap_net* net, int* tf32, uint32_t* round_bias);

#ifdef __cplusplus
}
#endif
#endif /* AUDIOPURE_B200_H */
