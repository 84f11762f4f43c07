/*
 * nst_b200.h - C ABI of the B200-native neural-style-transfer hot path.
 *
 * The reference (msmink01/text-based-image-style-transfer) has no FFI of its own: its boundary is the
 * Python import `from multi_style_transfer.run_style_transfer import run_multi_style_transfer`
 * (app.py:36).  This header is what our Python mirror of that package binds through ctypes; every
 * entry point names the reference code it replaces.  Plain C: device pointers are `void*`/`float*`
 * into CUDA memory owned by the caller (torch tensors), `stream` is a `cudaStream_t` passed as
 * `void*` (NULL = default stream).  All functions return 0 on success and a negative code on failure;
 * `nst_last_error()` returns the message of the calling thread's last failure.
 *
 * There is no CPU path: every compute entry fails with NST_ERR_DEVICE unless the current device is
 * compute capability 10.x.
 */
#ifndef NST_B200_H
#define NST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NST_ABI_VERSION 1

#define NST_MAX_CONV 16  /* conv1_1 .. conv5_4 of torchvision VGG-19 `features` */
#define NST_OK 0
#define NST_ERR_ARG -1
#define NST_ERR_CUDA -2
#define NST_ERR_DEVICE -3
#define NST_ERR_STATE -4
#define NST_ERR_UNSUPPORTED -5

typedef struct nst_net nst_net;   /* VGG-19 trunk: repacked weights (fp16 forward, bf16 data-gradient) */
typedef struct nst_plan nst_plan; /* per-resolution activations, tensor maps, targets, optimizer state */

/* loss vector written by nst_plan_eval / read back by nst_lbfgs_status (floats) */
enum {
  NST_LOSS_TOTAL = 0, /* run_style_transfer.py:139 */
  NST_LOSS_CONTENT_W = 1,
  NST_LOSS_STYLE_W = 2,
  NST_LOSS_TV_W = 3,
  NST_LOSS_EDGE_W = 4,
  NST_LOSS_GRAM0 = 5, /* 5..9: un-weighted per-layer Gram MSE (style_transfer_losses.py:138-144) */
  NST_LOSS_CONTENT = 10,
  NST_LOSS_STYLE = 11,
  NST_LOSS_TV = 12,
  NST_LOSS_EDGE = 13,
  NST_LOSS_COUNT = 16
};

typedef struct nst_status {
  int n_iter;        /* torch LBFGS state['n_iter'] */
  int func_evals;    /* state['func_evals'] */
  int closure_calls; /* iter[0] of run_style_transfer.py:99 */
  int stop;          /* 0 = the step ran to max_iter; else the reason it ended early (lbfgs_ctl.h) */
  int hist_len;
  int reserved;
  double loss, prev_loss, t, H_diag, gtd, gmax, max_td;
  float losses[NST_LOSS_COUNT];
} nst_status;

int nst_abi_version(void);
const char* nst_last_error(void);
/* 0 when the current CUDA device can run this library (sm_100), NST_ERR_DEVICE otherwise */
int nst_device_check(void);

/* ---- VGG-19 trunk ------------------------------------------------------------------------------
 * Replaces Vgg19.__init__ (helper_functions.py:44-92): `weights[i]` / `biases[i]` are device pointers
 * to conv i's torch-layout fp32 parameters ([Cout,Cin,3,3], [Cout]) for the first `n_conv` convs of
 * torchvision's vgg19().features.  Parameters are copied and repacked; the caller may free them. */
int nst_net_create(nst_net** out, const float* const* weights, const float* const* biases, int n_conv, void* stream);
void nst_net_destroy(nst_net* net);

/* ---- plan --------------------------------------------------------------------------------------
 * tap_mask: bit i set = conv i's pre-ReLU output is kept (Vgg19.forward's dict, helper_functions.py:94-101)
 * style_mask / content_mask: subsets of tap_mask that carry a Gram / content target.
 * with_grad != 0 also allocates the backward buffers and the L-BFGS state (history pairs). */
int nst_plan_create(nst_plan** out, const nst_net* net, int H, int W, uint32_t tap_mask, uint32_t style_mask,
                    uint32_t content_mask, int with_grad);
void nst_plan_destroy(nst_plan* plan);
/* bytes of device memory the plan holds */
size_t nst_plan_bytes(const nst_plan* plan);

/* normalize() constants (style_transfer_losses.py:9-28) and loss weights (run_style_transfer.py:115-139) */
int nst_plan_set_norm(nst_plan* plan, const float mean[3], const float std[3]);
int nst_plan_set_weights(nst_plan* plan, float w_style, float w_content, float w_tv, float w_edge);

/* Vgg19.forward(normalize(x)): x = [3,H,W] fp32 device; taps stay resident as NHWC fp16 */
int nst_plan_features(nst_plan* plan, const float* x, void* stream);
int nst_plan_tap_shape(const nst_plan* plan, int conv, int* C, int* H, int* W);
/* pre-ReLU feature of conv `conv` as [C,H,W] fp32 (torch layout) */
int nst_plan_get_tap(nst_plan* plan, int conv, float* out_chw, void* stream);
/* gram_matrix (style_transfer_losses.py:70-95) of a resident tap -> [C,C] fp32 */
int nst_plan_tap_gram(nst_plan* plan, int conv, float* out, void* stream);
/* gram_matrix of an arbitrary [C,H,W] fp32 device tensor (b = 1) -> [C,C] fp32 */
int nst_gram_chw(const float* x, int C, int H, int W, float* out, void* stream);
/* StyleMixer(...).mix() (StyleMixer.py:25-38) of the same tap of two plans, then gram_matrix -> [C,C] fp32 */
int nst_style_mix_gram(nst_plan* a, nst_plan* b, int conv, float weight_b, float* out, void* stream);
/* StyleMixer(...).mix() alone -> [C,Ho,Wo] fp32; Ho = Ha + Hb/2, Wo = Wa + Wb/2 */
int nst_style_mix_chw(nst_plan* a, nst_plan* b, int conv, float weight_b, float* out_chw, void* stream);

/* targets of the optimisation */
int nst_plan_set_style_target(nst_plan* plan, int conv, const float* gram /* [C,C] device */, void* stream);
/* content target = resident tap of `src` (same resolution) times an optional per-channel gate
 * (channel_att_per_chosen_layers, run_style_transfer.py:13-25) */
int nst_plan_set_content_target(nst_plan* plan, int conv, nst_plan* src, const float* gate /* [C] or NULL */,
                                void* stream);
/* ChannelAttention.forward's gate (ChannelAttention.py:30-37) from the resident tap: w1 [C/r,C], w2 [C,C/r] */
int nst_plan_channel_gate(nst_plan* plan, int conv, const float* w1, const float* w2, int reduction,
                          float* gate_out /* [C] device */, void* stream);
/* get_gradient_imgs(to_grayscale(normalize(content))) (run_style_transfer.py:73-74) */
int nst_plan_set_edge_target(nst_plan* plan, const float* content /* [3,H,W] device */, void* stream);

/* the closure (run_style_transfer.py:102-148) at x: losses -> NST_LOSS_COUNT floats, gradient -> [3,H,W]
 * (either may be NULL; both device pointers) */
int nst_plan_eval(nst_plan* plan, const float* x, float* losses, float* grad, void* stream);

/* individual loss terms on device tensors (function-level mirror of style_transfer_losses.py) */
int nst_tv_edge(const float* x, const float* edge_target /* [2,H,W] or NULL */, int H, int W, const float mean[3],
                const float std[3], float* out2 /* device: tv, edge */, void* stream);
int nst_edge_images(const float* gray_or_rgb, int channels, int H, int W, float* out /* [2,H,W] */, void* stream);

/* normalize (style_transfer_losses.py:9-28) of `planes` = b*3 planes of H*W */
int nst_normalize(const float* x, float* y, int planes, int C, int H, int W, const float mean[3], const float std[3],
                  void* stream);
/* to_grayscale (helper_functions.py:104-113): [C,H,W] -> [H,W] */
int nst_grayscale(const float* x, float* y, int C, int H, int W, void* stream);
/* nn.MSELoss(reduction='mean') of two equally shaped tensors -> device scalar */
int nst_mse(const float* a, const float* b, size_t n, float* out, void* stream);
/* total_variation_loss (style_transfer_losses.py:149-174) of `planes` = b*c planes -> device scalar */
int nst_total_variation(const float* y, int planes, int H, int W, float* out, void* stream);
/* ChannelAttention.forward (ChannelAttention.py:23-40) on a [C,H,W] tensor */
int nst_channel_attention_chw(const float* x, int C, int H, int W, const float* w1, const float* w2, int reduction,
                              float* y, void* stream);
/* StyleMixer.mix (StyleMixer.py:25-38) on two [C,H,W] tensors -> [C, Ha + Hb/2, Wa + Wb/2] */
int nst_style_mix_tensors(const float* a, int Ha, int Wa, const float* b, int Hb, int Wb, int C, float weight_b,
                          float* out, void* stream);

/* ---- L-BFGS (torch.optim.LBFGS defaults; run_style_transfer.py:90,99-151) ------------------------ */
int nst_lbfgs_init(nst_plan* plan, const float* x0 /* [3,H,W] device */, int trace_capacity, void* stream);
/* one optimizer.step(closure): up to 20 evaluations enqueued without host synchronisation */
int nst_lbfgs_step(nst_plan* plan, void* stream);
/* Captures and instantiates the CUDA graph nst_lbfgs_step launches (about 800 kernel nodes, milliseconds of host time)
 * without running it; nst_lbfgs_step does this lazily on its first call otherwise.  Call it after the last change of
 * weights / normalisation / trace capacity (each drops the graph) when the first step must not pay for the capture,
 * e.g. before a timed region.  No-op when steps are enqueued directly (NST_NO_GRAPH, or the NULL stream). */
int nst_lbfgs_prepare_graph(nst_plan* plan, void* stream);
/* synchronises the stream and copies the control block */
int nst_lbfgs_status(nst_plan* plan, nst_status* out, void* stream);
/* current iterate, clamped to [0,1] (run_style_transfer.py:153-155) -> [3,H,W] fp32 device */
int nst_lbfgs_get_x(nst_plan* plan, float* out, void* stream);
/* per-evaluation loss rows {total, content, style, tv, edge} (weighted); returns rows copied */
int nst_lbfgs_trace(nst_plan* plan, float* host_out, int max_rows, void* stream);
/* number of kernels one nst_lbfgs_step enqueues (for launch accounting) */
int nst_lbfgs_launches_per_step(const nst_plan* plan);

/* the first n_evals (1..20) evaluations of an optimizer.step(), enqueued directly (no graph); returns the number
 * of kernels launched.  Lets a benchmark time an exact number of evaluations. */
int nst_lbfgs_partial_step(nst_plan* plan, int n_evals, void* stream);

/* ---- per-launch timing with CUDA events (measurement only) ----------------------------------------- */
enum {
  NST_K_START = 0,
  NST_K_PIXEL = 1,       /* TV + edge losses and gradient */
  NST_K_CONV1_FWD = 2,   /* conv1_1 forward (CUDA cores) */
  NST_K_CONV_FWD = 3,    /* tcgen05 implicit-GEMM forward, `layer` = conv index */
  NST_K_GRAM = 4,        /* Gram split-K + both finalize kernels (3 launches) */
  NST_K_CONTENT = 5,
  NST_K_ASSEMBLE = 6,
  NST_K_GRAM_BWD = 7,    /* Gram backward as a tcgen05 1x1 convolution */
  NST_K_CONV_DGRAD = 8,  /* tcgen05 implicit-GEMM data gradient */
  NST_K_CONV1_DGRAD = 9,
  NST_K_LBFGS_PASS1 = 10,
  NST_K_LBFGS_REDUCE = 11,
  NST_K_LBFGS_CONTROL = 12,
  NST_K_LBFGS_PASS2 = 13
};
typedef struct nst_launch_time {
  int kind;
  int layer;
  float ms;
} nst_launch_time;
/* one closure evaluation with an event after every launch; returns the number of rows written */
int nst_plan_eval_timed(nst_plan* plan, const float* x, float* grad, nst_launch_time* out, int max_out, void* stream);
/* one whole optimizer.step() (20 evaluations + 20 L-BFGS iterations, single stream, no graph) with an event after every
 * launch, i.e. every kernel timed in the middle of a long busy stream; it advances the optimizer like nst_lbfgs_step.
 * Returns the number of rows written (about 900). */
int nst_lbfgs_step_timed(nst_plan* plan, nst_launch_time* out, int max_out, void* stream);
/* The same step with one event wherever the KIND of launch changes: a row is a run of same-kind launches (e.g. the
 * twelve forward convolutions) timed as a whole with programmatic dependent launch between them intact; `layer` holds
 * the number of launches in the run.  Side launches that normally sit between forward convolutions are issued after
 * the last one so that the runs are uninterrupted. */
int nst_lbfgs_step_timed_grouped(nst_plan* plan, nst_launch_time* out, int max_out, void* stream);

/* ---- tuning aids: exported ONLY by the instrumented build (tools/build.py --instrument -> libnst_b200_instr.so, compiled
 * with -DNST_INSTRUMENT and loaded by tools/ alone).  The product library contains neither these entry points nor the
 * stamps / wait counters / wrong-result timing experiments behind them. */
#ifdef NST_INSTRUMENT
/* phase timestamps (SM clock) of CTA 0 of one convolution launch (mode 0 forward, 1 data gradient; conv 0 = conv1_1's
   data gradient) -> out[0..6], the SM cycles its roles spent waiting -> out[8..13] (slots: csrc/conv_tc.cu) and the
   lifetime in SM cycles of every CTA -> out[16 + cta]; out holds 176 values: tuning aid, see tools/conv_phases.py */
int nst_plan_conv_phases(nst_plan* plan, int conv, int mode, long long* out14, void* stream);
/* launch spans {earliest CTA start, latest CTA end} (%globaltimer ns) of the convolution launches inside the captured
   step: slot = conv (forward), 16 + conv (data gradient), 32 + conv (Gram backward); enable re-captures the step with
   the slots armed, out96 (8 x 48 values, may be NULL) reads them back and re-arms: tuning aid, tools/conv_timeline.py */
int nst_plan_timeline(nst_plan* plan, int enable, unsigned long long* out96, void* stream);
/* SM clock at the phase boundaries of the most recent L-BFGS controller launch (slots: csrc/lbfgs_ctl.h), 8 values:
   tuning aid, tools/ctl_phases.py */
int nst_lbfgs_ctl_clocks(nst_plan* plan, long long* out8, void* stream);
#endif /* NST_INSTRUMENT */

/* ---- host-buffer convenience (the e2e path: copies inside) ---------------------------------------
 * content_u8: [H,W,3] uint8 host; out_u8: [H,W,3] uint8 host (truncating, like ToPILImage).
 * Requires style / content / edge targets to be refreshed from the content image, which this call
 * does (content target at the plan's content taps, edge target), then runs the whole loop. */
int nst_run_frame_host(nst_plan* plan, const uint8_t* content_u8, uint8_t* out_u8, int num_steps, int channel_attention,
                       const float* ca_w1, const float* ca_w2, void* stream);
/* shared != 0: other plans step on the same GPU at the same time as this one (several host threads / streams).  Their CTA-pair
 * convolution launches are then issued without programmatic dependent launch - two such chains side by side can hang the GPU -
 * at ~5 % of a single chain's speed.  nst_run_frames_host sets it itself for count > 1; nst_batch_create needs none (one chain). */
int nst_plan_set_shared_gpu(nst_plan* plan, int shared);
/* the same for `count` (<= 16) independent frames at once: plans[k] / streams[k] (all distinct, same weights and targets as a
 * frame-by-frame run would use) process content_u8[k] -> out_u8[k]; every frame's optimizer.step() is enqueued before the host
 * waits for the first one, so small frames (bound by per-launch latency) overlap: the per-frame body of apply_video_process,
 * app.py:784-815, several frames at a time.  Results are bit-identical to nst_run_frame_host.
 * closure_calls (optional): evaluations each frame ran. */
int nst_run_frames_host(nst_plan* const* plans, int count, const uint8_t* const* content_u8, uint8_t* const* out_u8, int num_steps,
                        int channel_attention, const float* ca_w1, const float* ca_w2, void* const* streams, int* closure_calls);

/* ---- several images per launch (SURVEY 8 f row 2: "batch dimension b > 1 in conv / Gram") -------------------------------------
 * The reference evaluates one image per call; its gram_matrix already normalises by the batch size b
 * (multi_style_transfer/style_transfer_losses.py:84-93) and apply_video_process (app.py:784-815) runs the same loop on
 * independent frames.  nst_batch_create makes a HEAD plan for `batch` (2..8) images of H x W (multiples of 16) plus `batch`
 * MEMBER plans: the head owns every tensor the tcgen05 convolution / Gram / Gram-backward launches touch as [batch][...] arrays
 * and processes all images in ONE launch each; the members are ordinary single-image plans over their slices and carry what is
 * per image (targets, pixel-space losses, conv1_1, the L-BFGS state).  Set targets / initialise / read status, trace and result
 * through the members (nst_batch_member); set norm / weights on the head (forwarded); step with nst_lbfgs_step(head): one CUDA
 * graph per optimizer.step() of all members.  nst_plan_destroy(head) destroys the members. */
int nst_batch_create(nst_plan** head_out, const nst_net* net, int H, int W, uint32_t tap_mask, uint32_t style_mask,
                     uint32_t content_mask, int batch);
int nst_batch_size(const nst_plan* plan);                 /* images per launch: `batch` for a head, 1 otherwise */
nst_plan* nst_batch_member(nst_plan* head, int index);    /* NULL + nst_last_error() when out of range */
/* frozen != 0: the member sits out the following nst_lbfgs_step(head) calls (its own loop has ended, run_style_transfer.py:100,
 * while other members still run); cleared by nst_lbfgs_init.  Asynchronous on `stream`. */
int nst_lbfgs_freeze(nst_plan* member, int frozen, void* stream);
/* nst_run_frames_host for a head: frame k (k < count <= batch) on member k, host uint8 in / out, copies inside. */
int nst_run_batch_host(nst_plan* head, int count, const uint8_t* const* content_u8, uint8_t* const* out_u8, int num_steps,
                       int channel_attention, const float* ca_w1, const float* ca_w2, void* stream, int* closure_calls);

/* ---- mask compositing: text/segmentation_style_transfer.py:5-94 (segmentation_style_transfer + _edge_smoothing), the step that
 * follows run_multi_style_transfer in app.py:203,318,407,512.  content, style, out: [H][W][C] uint8 device buffers (C <= 4),
 * mask: [H][W] bytes (non-zero = stylised pixel).  edge_smoothing = 0 selects (np.where, :52); otherwise the mask is blurred
 * with cv2.GaussianBlur's 8-bit fixed-point arithmetic (even sizes + 1, :76-77; at most 127) and the images are blended in
 * fp64 and truncated (:91).  Bit-exact with the reference. */
int nst_mask_composite(const uint8_t* content, const uint8_t* style, const uint8_t* mask, int H, int W, int C, int edge_smoothing,
                       uint8_t* out, void* stream);
/* the k integer weights (sum 256) of that blur, host side; k odd, 1..127 */
int nst_mask_gaussian_weights(int k, int* w);

/* ---- video back end: the frame list apply_video_process hands to the encoder, app.py:800-806 (cv2.cvtColor RGB2BGR of every
 * stylised frame) and app.py:820-840 (n_interp cross-dissolved frames cv2.addWeighted(prev, 1 - a, frame, a, 0),
 * a = (i + 1) / (n_interp + 1), between consecutive frames).  frames_rgb: [F][H][W][3] uint8 device; out_bgr:
 * [(F - 1) * (n_interp + 1) + 1][H][W][3] uint8 device.  n_interp <= 15 (the reference's slider: 0..5).  Bit-exact with
 * OpenCV's 8-bit addWeighted (fp32 fma(a, alpha, b * beta), round to nearest even). */
int nst_video_assemble(const uint8_t* frames_rgb, int F, int H, int W, int n_interp, uint8_t* out_bgr, void* stream);

/* ---- multi-plane depth style transfer: components/style_transfer_depth/util.py:9-35,53-66 (mask_image_depth over the bins of
 * create_bins = generate_mip_layers) and :69-88 (reconstruct_mip_image).  depth: [H][W] uint8 (depth_f64 = 0; dmin / dmax = the
 * map's minimum / maximum, normalised in the kernel as numpy does: (d - min) / (max - min) in fp64) or the already normalised
 * [H][W] fp64 map (depth_f64 = 1).  lo / hi: the n inclusive bin bounds (host arrays).  n <= 16 (the reference's slider: 2..10).
 * split: image [H][W][C] -> out [n][H][W][C], pixels outside bin i zeroed in plane i.
 * merge: planes [n][H][W][3] -> out [H][W][3], the masked planes added up in uint8 (wraps where bins overlap). */
int nst_mip_split(const uint8_t* image, const void* depth, int depth_f64, int dmin, int dmax, int H, int W, int C, int n,
                  const double* lo, const double* hi, uint8_t* out, void* stream);
int nst_mip_merge(const uint8_t* planes, const void* depth, int depth_f64, int dmin, int dmax, int H, int W, int n,
                  const double* lo, const double* hi, uint8_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NST_B200_H */
