// C ABI of the library (include/nst_b200.h): VGG-19 trunk, per-resolution plan, the fused closure
// evaluation and the device-resident L-BFGS loop.  Host-side orchestration only; every kernel lives in
// conv_tc.cu / gram.cu / pixel.cu / lbfgs.cu.
//
// Reference call stack this file re-hosts (paths under /root/reference/multi_style_transfer/):
//   run_multi_style_transfer   run_style_transfer.py:27-159
//   closure                    run_style_transfer.py:102-148
//   Vgg19.forward              helper_functions.py:94-101
#include "../../include/nst_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "conv_tc.cuh"
#include "gram.cuh"
#include "lbfgs.cuh"
#include "mask.cuh"
#include "video.cuh"
#include "depth.cuh"
#include "pixel.cuh"

using namespace nst;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(expr)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess) return fail(NST_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                                        __FILE__, __LINE__);                                             \
  } while (0)
#define CKI(expr)                  \
  do {                             \
    int r__ = (expr);              \
    if (r__ != NST_OK) return r__; \
  } while (0)

extern "C" int nst_abi_version(void) { return NST_ABI_VERSION; }
extern "C" const char* nst_last_error(void) { return g_err; }

static int g_num_sms = 0;
// SMs the convolution launches may occupy (NST_CONV_SMS; default all): leaves the rest to work of another frame that runs beside them
static int g_conv_sms = 0;
extern "C" int nst_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(NST_ERR_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  // cudaGetDeviceProperties costs milliseconds: a device that passed once is not queried again
  static bool checked[64] = {};
  if (dev >= 0 && dev < 64 && checked[dev] && g_num_sms != 0) return NST_OK;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return fail(NST_ERR_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(NST_ERR_DEVICE, "device %d is sm_%d%d; this library contains sm_100a code only (no fallback)", dev,
                prop.major, prop.minor);
  if (g_num_sms == 0) {
    e = conv_tc_init();
    if (e == cudaSuccess) e = conv1_tc_init();
    if (e == cudaSuccess) e = gram_init();
    if (e == cudaSuccess) e = lbfgs_init();
    if (e != cudaSuccess) return fail(NST_ERR_CUDA, "kernel attribute setup: %s", cudaGetErrorString(e));
    g_num_sms = prop.multiProcessorCount;
    g_conv_sms = g_num_sms;
    if (getenv("NST_CONV_SMS") != nullptr) {
      const int v = atoi(getenv("NST_CONV_SMS"));
      if (v >= 2 && v <= g_num_sms) g_conv_sms = v & ~1;
    }
  }
  if (dev >= 0 && dev < 64) checked[dev] = true;
  return NST_OK;
}

// ------------------------------------------------------------------------------------------------
// VGG-19 `features` configuration E (torchvision/models/vgg.py:94), first 16 convolutions
// ------------------------------------------------------------------------------------------------
static const int kCin[NST_MAX_CONV] = {3, 64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512, 512, 512, 512};
static const int kCout[NST_MAX_CONV] = {64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512, 512, 512, 512, 512};
static const int kLevel[NST_MAX_CONV] = {0, 0, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
static const int kPoolAfter[NST_MAX_CONV] = {0, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1};

struct nst_net {
  int n_conv = 0;
  float* w32[NST_MAX_CONV] = {};
  float* b32[NST_MAX_CONV] = {};
  __half* wf[NST_MAX_CONV] = {};         // [9][Cout][Cin]
  __nv_bfloat16* wb[NST_MAX_CONV] = {};  // [9][Cin][Cout], taps flipped
};

extern "C" void nst_net_destroy(nst_net* net) {
  if (!net) return;
  for (int i = 0; i < NST_MAX_CONV; ++i) {
    cudaFree(net->w32[i]);
    cudaFree(net->b32[i]);
    cudaFree(net->wf[i]);
    cudaFree(net->wb[i]);
  }
  delete net;
}

extern "C" int nst_net_create(nst_net** out, const float* const* weights, const float* const* biases, int n_conv,
                              void* stream) {
  if (!out || !weights || !biases || n_conv < 1 || n_conv > NST_MAX_CONV) return fail(NST_ERR_ARG, "nst_net_create: bad arguments");
  CKI(nst_device_check());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  nst_net* net = new nst_net();
  net->n_conv = n_conv;
  for (int i = 0; i < n_conv; ++i) {
    const size_t nw = static_cast<size_t>(kCout[i]) * kCin[i] * 9;
    cudaError_t e = cudaMalloc(&net->w32[i], nw * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&net->b32[i], kCout[i] * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpyAsync(net->w32[i], weights[i], nw * sizeof(float), cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(net->b32[i], biases[i], kCout[i] * sizeof(float), cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && i == 0) {
      e = cudaMalloc(&net->wb[0], static_cast<size_t>(9) * 16 * 64 * sizeof(__nv_bfloat16));
      if (e == cudaSuccess) e = launch_pack_weights_conv1_bwd(net->w32[0], net->wb[0], s);
    }
    if (e == cudaSuccess && i >= 1) {
      e = cudaMalloc(&net->wf[i], nw * sizeof(__half));
      if (e == cudaSuccess) e = cudaMalloc(&net->wb[i], nw * sizeof(__nv_bfloat16));
      if (e == cudaSuccess) e = launch_pack_weights(net->w32[i], net->wf[i], net->wb[i], kCout[i], kCin[i], s);
    }
    if (e != cudaSuccess) {
      nst_net_destroy(net);
      return fail(NST_ERR_CUDA, "nst_net_create: conv %d: %s", i, cudaGetErrorString(e));
    }
  }
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    nst_net_destroy(net);
    return fail(NST_ERR_CUDA, "nst_net_create: %s", cudaGetErrorString(e));
  }
  *out = net;
  return NST_OK;
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct nst_plan {
  const nst_net* net = nullptr;
  int H = 0, W = 0;
  int n_layers = 0;
  uint32_t tap_mask = 0, style_mask = 0, content_mask = 0;
  int with_grad = 0;
  int lh[5] = {}, lw[5] = {};  // resolution per pooling level
  PixelConsts pc = {{0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}};
  float w_style = 0.f, w_content = 0.f, w_tv = 0.f, w_edge = 0.f;
  size_t bytes = 0;
  std::vector<void*> allocs;

  // forward
  __half* act[NST_MAX_CONV] = {};
  __half* tap[NST_MAX_CONV] = {};
  uint8_t* route[NST_MAX_CONV] = {};
  ConvParams fwd[NST_MAX_CONV];
  // backward
  __nv_bfloat16* gpre[NST_MAX_CONV] = {};
  __nv_bfloat16* gadd[NST_MAX_CONV] = {};
  ConvParams dgrad[NST_MAX_CONV];
  ConvParams scale[NST_MAX_CONV];
  // style
  int n_style = 0;
  int style_conv[GRAM_MAX_LAYERS] = {};
  float* style_target[GRAM_MAX_LAYERS] = {};
  __half* dh[GRAM_MAX_LAYERS] = {};
  float* alpha = nullptr;       // [GRAM_MAX_LAYERS]
  float* dh_scale = nullptr;    // [GRAM_MAX_LAYERS] power-of-two scale of the fp16 Gram-backward operand (from the target)
  const float* style_fin[GRAM_MAX_LAYERS] = {};   // per style layer: per-block sums of (G - T)^2 inside a set's fin_part
  int style_fin_n[GRAM_MAX_LAYERS] = {};
  GramParams gram_shallow;  // every style layer except a style layer on the deepest conv (side stream)
  // The same layers planned for the SMs the convolution chain leaves idle while this launch runs beside it (conv4_2 ..
  // conv5_1 and the first data gradients have 128 work items on 148 SMs at 512^2): a persistent launch on that many CTAs
  // never takes an SM a convolution tile is waiting for.  r01's 148-CTA launch pushed conv4_4 into a second wave (+15 us
  // on the critical path).  num_layers == 0: no idle SMs at this resolution, the wide plan is used.
  GramParams gram_shallow_narrow;
  const float* style_fin_narrow[GRAM_MAX_LAYERS] = {};
  int style_fin_n_narrow[GRAM_MAX_LAYERS] = {};
  GramParams gram_deep;     // the style layer on the deepest conv, if any (critical path)
  float* gram_ws = nullptr;
  cudaStream_t side = nullptr;
  cudaStream_t side2 = nullptr;  // the shallow layers' Gram launch when it runs on the idle SMs: ~100 us long, so the short side work
                                 // (content loss and its seed, loss assembly) must not queue behind it
  cudaEvent_t ev[8] = {};
  // content
  int n_content = 0;
  int content_conv[NST_MAX_CONV] = {};
  float* content_target[NST_MAX_CONV] = {};
  double* content_part = nullptr;
  int content_part_off[NST_MAX_CONV + 1] = {};
  // pixel
  float* tedge = nullptr;
  float* grad_pix = nullptr;
  double* tv_part = nullptr;
  double* edge_part = nullptr;
  float* losses = nullptr;  // [NST_LOSS_COUNT]
  // optimizer
  LbfgsBuffers lb = {};
  float* trace = nullptr;
  int trace_cap = 0;
  cudaGraphExec_t step_graph = nullptr;
  int launches_per_step = 0;
  // staging for nst_run_frame_host
  uint8_t* u8_dev = nullptr;
  float* img_dev = nullptr;
  float* gate_dev = nullptr;
  float* pooled_dev = nullptr;
  // seed_folded[i]: the Gram backward of style layer conv i runs inside the data gradient of conv i+1 (no launch of its own)
  bool seed_folded[NST_MAX_CONV] = {};
  // ---- several images per launch (nst_batch_create).  The HEAD (batch > 1) owns the tensors the convolution and
  // Gram launches touch - taps, activations, routing bytes, gradients, seeds, Gram targets / backward operands / scalars - as
  // [batch][...] arrays and the parameter sets that process all images in one launch each.  Its MEMBERS (head != nullptr) are
  // ordinary single-image plans whose pointers to those tensors are slices of the head's: image `index`.  Everything per
  // image - pixel-space losses, content loss, loss assembly, conv1_1 (fp32 planar image on either side), the optimizer and
  // its history - runs per member; a member is also usable on its own (targets from the content frame, status, trace).
  bool shared_gpu = false;   // nst_plan_set_shared_gpu: other plans step on this GPU at the same time
  int batch = 1;
  nst_plan* head = nullptr;
  int index = 0;
  std::vector<nst_plan*> members;
  int style_fin_stride[GRAM_MAX_LAYERS] = {};         // head: per-image distance inside fin_part of the layer's Gram set
  int style_fin_stride_narrow[GRAM_MAX_LAYERS] = {};
#ifdef NST_INSTRUMENT
  unsigned long long* timeline = nullptr;  // [48][8] launch spans of the conv kernels (nst_plan_timeline)
  bool timeline_on = false;
#endif
};

static int plan_alloc(nst_plan* p, void** ptr, size_t bytes, bool zero) {
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) return fail(NST_ERR_CUDA, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
  if (zero) {
    e = cudaMemset(*ptr, 0, bytes);
    if (e != cudaSuccess) return fail(NST_ERR_CUDA, "cudaMemset: %s", cudaGetErrorString(e));
  }
  p->allocs.push_back(*ptr);
  p->bytes += bytes;
  return NST_OK;
}
template <typename T>
static int plan_alloc_t(nst_plan* p, T** ptr, size_t count, bool zero = false) {
  return plan_alloc(p, reinterpret_cast<void**>(ptr), count * sizeof(T), zero);
}

static void drop_graph(nst_plan* p) {
  if (p->step_graph) {
    cudaGraphExecDestroy(p->step_graph);
    p->step_graph = nullptr;
  }
}

extern "C" void nst_plan_destroy(nst_plan* p) {
  if (!p) return;
  for (nst_plan* m : p->members) nst_plan_destroy(m);
  p->members.clear();
  drop_graph(p);
  if (p->side) cudaStreamDestroy(p->side);
  if (p->side2) cudaStreamDestroy(p->side2);
  for (int k = 0; k < 8; ++k)
    if (p->ev[k]) cudaEventDestroy(p->ev[k]);
  for (void* a : p->allocs) cudaFree(a);
  delete p;
}
extern "C" size_t nst_plan_bytes(const nst_plan* p) { return p ? p->bytes : 0; }

static int style_index(const nst_plan* p, int conv) {
  for (int l = 0; l < p->n_style; ++l)
    if (p->style_conv[l] == conv) return l;
  return -1;
}
static int content_index(const nst_plan* p, int conv) {
  for (int l = 0; l < p->n_content; ++l)
    if (p->content_conv[l] == conv) return l;
  return -1;
}

static int build_conv_params(nst_plan* p) {
  const nst_net* net = p->net;
  // epilogue outputs through shared memory + TMA stores (conv_epilogue.cuh); NST_DIRECT_STORES=1 keeps the per-thread stores
  const bool tma_out = getenv("NST_DIRECT_STORES") == nullptr;
  {
    // ---- conv1_1 forward on tensor cores (conv1_tc.cu): only the epilogue half of ConvParams is used
    ConvParams& f = p->fwd[0];
    memset(&f, 0, sizeof(f));
    f.H = p->H;
    f.W = p->W;
    f.K = 64;
    f.N = 64;
    f.taps = 9;
    f.block_n = 64;
    f.bias = net->b32[0];
    f.out_tap = p->tap[0];
    f.out_act = p->n_layers > 1 ? p->act[0] : nullptr;
    conv_finalize_params(f, CONV_FWD);
#ifdef NST_INSTRUMENT
    f.dbg_flags = getenv("NST_DBG_CONV1") ? atoi(getenv("NST_DBG_CONV1")) : 0;  // timing experiments only (instrumented build)
#endif
    if (tma_out && getenv("NST_CONV1_CUDA_CORES") == nullptr) {
      if (f.out_tap && make_tmap_out(&f.tmO0, f.out_tap, p->H, p->W, 64, 64, CONV_TILE_W, 4) != 0) return fail(NST_ERR_CUDA, "tensor map (tap out 0)");
      if (f.out_act && make_tmap_out(&f.tmO1, f.out_act, p->H, p->W, 64, 64, CONV_TILE_W, 4) != 0) return fail(NST_ERR_CUDA, "tensor map (act out 0)");
      f.tma_out = 1;
    }
  }
  const int B = p->batch;   // head: every launch below covers B images ([B][H][W][C] tensors, stacked view B * H rows for outputs)
  for (int i = 1; i < p->n_layers; ++i) {
    const int lv = kLevel[i];
    const int H = p->lh[lv], W = p->lw[lv];
    const bool pooled = kPoolAfter[i] && i < p->n_layers - 1;
    // ---- forward
    ConvParams& f = p->fwd[i];
    memset(&f, 0, sizeof(f));
    f.H = H;
    f.W = W;
    f.K = kCin[i];
    f.N = kCout[i];
    f.taps = 9;
    f.batch = B;
    if (make_tmap_act(&f.tmA, p->act[i - 1], H, W, kCin[i], 64, CONV_TILE_W + 2, CONV_TILE_H + 2, B) != 0) return fail(NST_ERR_CUDA, "tensor map (act %d)", i);
    f.block_n = conv_block_n(kCout[i], H * B, W, 9 * kCin[i], g_num_sms);
    f.pair = conv_use_pair(CONV_FWD, f.block_n, 9, kCin[i], kCout[i]) ? 1 : 0;
    if (make_tmap_wgt(&f.tmB, net->wf[i], 9, kCout[i], kCin[i], f.block_n, f.pair != 0) != 0)
      return fail(NST_ERR_CUDA, "tensor map (weights %d)", i);
    f.bias = net->b32[i];
    f.out_tap = p->tap[i];
    f.out_act = p->act[i];
    f.out_route = p->route[i];
    f.pool = pooled ? 1 : 0;
    conv_finalize_params(f, CONV_FWD);
    if (tma_out) {
      if (f.out_tap && make_tmap_out(&f.tmO0, f.out_tap, H, W, kCout[i], 32, CONV_TILE_W, 4, B) != 0) return fail(NST_ERR_CUDA, "tensor map (tap out %d)", i);
      if (!pooled && f.out_act && make_tmap_out(&f.tmO1, f.out_act, H, W, kCout[i], 32, CONV_TILE_W, 4, B) != 0) return fail(NST_ERR_CUDA, "tensor map (act out %d)", i);
      f.tma_out = 1;
    }
    if (!p->with_grad) continue;
    // ---- data-gradient of conv i: consumes gpre[i], produces gpre[i-1]
    ConvParams& d = p->dgrad[i];
    memset(&d, 0, sizeof(d));
    d.H = H;
    d.W = W;
    d.K = kCout[i];
    d.N = kCin[i];
    d.taps = 9;
    d.batch = B;
    if (make_tmap_act(&d.tmA, p->gpre[i], H, W, kCout[i], 64, CONV_TILE_W + 2, CONV_TILE_H + 2, B) != 0) return fail(NST_ERR_CUDA, "tensor map (grad %d)", i);
    d.block_n = conv_block_n(kCin[i], H * B, W, 9 * kCout[i], g_num_sms);
    d.pair = conv_use_pair(CONV_DGRAD, d.block_n, 9, kCout[i], kCin[i]) ? 1 : 0;
    if (make_tmap_wgt(&d.tmB, net->wb[i], 9, kCin[i], kCout[i], d.block_n, d.pair != 0) != 0)
      return fail(NST_ERR_CUDA, "tensor map (weights^T %d)", i);
    const bool prev_pooled = kPoolAfter[i - 1] != 0;  // conv i reads the pooled output of conv i-1
    d.out_grad = p->gpre[i - 1];
    if (prev_pooled) {
      d.route = p->route[i - 1];
      d.Hup = p->lh[kLevel[i - 1]];
      d.Wup = p->lw[kLevel[i - 1]];
    } else {
      // ReLU mask: act = relu(tap) rounded the same way, so (tap > 0) == (act > 0).  A style layer's data gradient already
      // pulls that tap through L2 for the folded Gram backward; masking from it saves the DRAM read of the activation
      // (33.5 MB at conv1_1 / 512^2, the slowest launch of the step).
      d.mask_act = p->tap[i - 1] != nullptr ? p->tap[i - 1] : p->act[i - 1];
      d.addend = p->gadd[i - 1];
    }
    conv_finalize_params(d, CONV_DGRAD);
    if (tma_out) {
      const int rc = prev_pooled ? make_tmap_out(&d.tmO0, d.out_grad, d.Hup, d.Wup, kCin[i], 16, 2 * CONV_TILE_W, 8, B)
                                 : make_tmap_out(&d.tmO0, d.out_grad, H, W, kCin[i], 32, CONV_TILE_W, 4, B);
      if (rc != 0) return fail(NST_ERR_CUDA, "tensor map (grad out %d)", i);
      // Measured (profiles/r01_conv_phases_tma_store.log, r01_timeline_512_tma_store.log): data gradients keep their direct
      // 16-byte stores.  The plain masked gradient gains nothing from the TMA store (conv1_2: 40 -> 50 us).  The pool-routing
      // scatter gains 2-6 us per layer in isolation, but inside the step the NEXT layer, which reads that gradient, loses
      // more (conv1_2: +15 us, conv2_2: +7 us): 32-byte box rows reach L2 as single-sector writes.  NST_TMA_DGRAD=1 opts in.
      static const bool all_dgrad = getenv("NST_TMA_DGRAD") != nullptr;
      d.tma_out = all_dgrad ? 1 : 0;
    }
  }
  if (!p->with_grad) return NST_OK;
  {
    // ---- conv1_1's data gradient on tensor cores: N = 16 (3 image channels + zero padding), K = 9 x 64
    ConvParams& d = p->dgrad[0];
    memset(&d, 0, sizeof(d));
    d.H = p->H;
    d.W = p->W;
    d.K = 64;
    d.N = 16;
    d.taps = 9;
    d.block_n = 16;
    if (make_tmap_act(&d.tmA, p->gpre[0], p->H, p->W, 64, 64, CONV_TILE_W + 2, CONV_TILE_H + 2) != 0) return fail(NST_ERR_CUDA, "tensor map (grad 0)");
    if (make_tmap_wgt(&d.tmB, net->wb[0], 9, 16, 64, 16) != 0) return fail(NST_ERR_CUDA, "tensor map (weights^T 0)");
    d.grad_pix = p->grad_pix;
    d.out_pix = nullptr;  // set per call
    conv_finalize_params(d, CONV_DGRAD_PIX);
  }
  // ---- Gram backward as a 1x1 convolution: seed[l] = alpha_l * F_l * (G_l - T_l)/max|.|
  for (int l = 0; l < p->n_style; ++l) {
    const int i = p->style_conv[l];
    const int lv = kLevel[i];
    const int C = kCout[i];
    ConvParams& c = p->scale[i];
    memset(&c, 0, sizeof(c));
    c.H = p->lh[lv];
    c.W = p->lw[lv];
    c.K = C;
    c.N = C;
    c.taps = 1;
    c.batch = B;
    c.b_per_image = 1;                 // dh: [B][C][C], the image in the map's third dimension
    c.alpha_stride = GRAM_MAX_LAYERS;  // alpha: [B][GRAM_MAX_LAYERS]
    if (make_tmap_act(&c.tmA, p->tap[i], c.H, c.W, C, 64, CONV_TILE_W, CONV_TILE_H, B) != 0) return fail(NST_ERR_CUDA, "tensor map (tap %d)", i);
    c.block_n = conv_block_n(C, c.H * B, c.W, C, g_num_sms);
    if (make_tmap_wgt(&c.tmB, p->dh[l], B, C, C, c.block_n) != 0) return fail(NST_ERR_CUDA, "tensor map (dh %d)", i);
    c.alpha = p->alpha + l;
    c.out_grad = i == p->n_layers - 1 ? p->gpre[i] : p->gadd[i];
    conv_finalize_params(c, CONV_SCALE);
    if (tma_out) {
      if (make_tmap_out(&c.tmO0, c.out_grad, c.H, c.W, C, 32, CONV_TILE_W, 4, B) != 0) return fail(NST_ERR_CUDA, "tensor map (seed out %d)", i);
      c.tma_out = 1;
    }
  }
  // ---- fold the Gram backward of a style layer into the data gradient that produces that layer's gradient
  // (conv_tc.cuh: seed_k).  Not for the deepest conv (its seed IS the first gradient), not when the layer is also a
  // content layer (that seed is accumulated by the content kernel), not with N tiles above 128 (tensor memory).
  for (int i = 0; i < NST_MAX_CONV; ++i) p->seed_folded[i] = false;
  if (getenv("NST_NO_SEED_FOLD") == nullptr) {
    for (int l = 0; l < p->n_style; ++l) {
      const int j = p->style_conv[l];      // style layer conv j; its gradient gpre[j] is produced by dgrad[j + 1]
      const int i = j + 1;
      if (j == p->n_layers - 1 || i >= p->n_layers || kPoolAfter[j] || content_index(p, j) >= 0) continue;
      ConvParams& d = p->dgrad[i];
      if (d.block_n > 128 || d.route != nullptr) continue;
      const int C = kCout[j];
      if (make_tmap_act(&d.tmA2, p->tap[j], d.H, d.W, C, 64, CONV_TILE_W, CONV_TILE_H, B) != 0) return fail(NST_ERR_CUDA, "tensor map (tap %d, folded)", j);
      if (make_tmap_wgt(&d.tmB2, p->dh[l], B, C, C, d.block_n, d.pair != 0) != 0) return fail(NST_ERR_CUDA, "tensor map (dh %d, folded)", j);
      d.b_per_image = 1;
      d.alpha_stride = GRAM_MAX_LAYERS;
      d.seed_k = C;
      d.idesc2 = umma_idesc_f16(d.pair ? 256 : 128, d.block_n, 0, 0, 0);
      d.alpha = p->alpha + l;
      d.addend = nullptr;
      p->seed_folded[j] = true;
    }
  }
  return NST_OK;
}

// Two parameter sets: the style layer on the deepest conv (critical path, main stream) and all the others (side stream).
static int build_gram_params(nst_plan* p) {
  GramParams* sets[2] = {&p->gram_shallow, &p->gram_deep};
  memset(sets[0], 0, sizeof(GramParams));
  memset(sets[1], 0, sizeof(GramParams));
  for (int l = 0; l < p->n_style; ++l) {
    const int i = p->style_conv[l];
    const int lv = kLevel[i];
    GramParams& g = *sets[i == p->n_layers - 1 ? 1 : 0];
    const int k = g.num_layers++;
    GramLayer& L = g.L[k];
    L.C = kCout[i];
    L.HW = p->lh[lv] * p->lw[lv];
    L.inv_norm = 1.f / (static_cast<float>(L.C) * static_cast<float>(L.HW));
    L.target = p->style_target[l];
    L.target_stride = static_cast<size_t>(L.C) * L.C;   // head: [batch][C][C], one target per image
    L.gram_out = nullptr;   // only G - T's fp16 image (dh) and its sum of squares are consumed
    L.dh = p->dh[l];
    L.dh_scale = p->dh_scale + l;
    L.alpha = p->alpha + l;
    // d/dF of w_s/n_style * mean((G-T)^2), G = F F^T / (C HW): 4 w_s (G-T) F / (n_style C^3 HW)
    L.grad_coef = 4.f * p->w_style / (static_cast<float>(p->n_style) * static_cast<float>(L.C) *
                                      static_cast<float>(L.C) * static_cast<float>(L.C) * static_cast<float>(L.HW));
    if (make_tmap_feat(&g.tm[k], p->tap[i], L.HW, L.C, p->batch) != 0) return fail(NST_ERR_CUDA, "tensor map (gram %d)", i);
  }
  return NST_OK;
}

static int conv_grid(const ConvParams& c) { return conv_grid_ctas(c, g_num_sms); }

// Plans the shallow style layers' Gram launch for the SMs that stay idle beside the deeper convolutions (nst_plan).
static int build_gram_narrow(nst_plan* p) {
  memset(&p->gram_shallow_narrow, 0, sizeof(GramParams));
  for (int l = 0; l < GRAM_MAX_LAYERS; ++l) {
    p->style_fin_narrow[l] = p->style_fin[l];
    p->style_fin_n_narrow[l] = p->style_fin_n[l];
    p->style_fin_stride_narrow[l] = p->style_fin_stride[l];
  }
  const GramParams& wide = p->gram_shallow;
  if (wide.num_layers == 0 || p->side == nullptr || getenv("NST_GRAM_WIDE") != nullptr) return NST_OK;
  const int last = p->n_layers - 1;
  int max_shallow = -1;
  for (int l = 0; l < p->n_style; ++l)
    if (p->style_conv[l] != last) max_shallow = p->style_conv[l];
  // launches that run while the shallow Gram work is in flight: forward convolutions behind its last tap, the deepest
  // layer's Gram + Gram backward, and the data gradients down to the one that consumes its results
  int busiest = 0;
  for (int i = max_shallow + 1; i <= last; ++i) {
    if (conv_grid(p->fwd[i]) > busiest) busiest = conv_grid(p->fwd[i]);
    if (p->with_grad && i >= max_shallow + 2 && conv_grid(p->dgrad[i]) > busiest) busiest = conv_grid(p->dgrad[i]);
  }
  if (max_shallow + 1 > last) return NST_OK;   // nothing runs beside it
  const int idle = g_num_sms - busiest;
  if (idle < 8) return NST_OK;
  GramParams& g = p->gram_shallow_narrow;
  g = wide;
  const size_t ws = gram_plan(g, idle / p->batch > 0 ? idle / p->batch : 1);
  g.max_ctas = idle;
  float* wsp = nullptr;
  float* fin = nullptr;
  CKI(plan_alloc_t(p, &wsp, ws * p->batch));
  CKI(plan_alloc_t(p, &fin, static_cast<size_t>(g.num_fin_blocks) * p->batch, true));
  g.ws = wsp;
  g.fin_part = fin;
  g.ws_img = ws;
  for (int j = 0; j < g.num_layers; ++j) {
    const GramLayer& L = g.L[j];
    for (int q = 0; q < p->n_style; ++q)
      if (p->dh[q] == L.dh) {
        p->style_fin_narrow[q] = fin + L.fin_blk0;
        p->style_fin_n_narrow[q] = L.fused ? L.pairs : L.fin_blocks;
        p->style_fin_stride_narrow[q] = g.num_fin_blocks;
      }
  }
  return NST_OK;
}

static int plan_create_impl(nst_plan** out, const nst_net* net, int H, int W, uint32_t tap_mask, uint32_t style_mask,
                            uint32_t content_mask, int with_grad, int batch, nst_plan* head, int index);

extern "C" int nst_plan_create(nst_plan** out, const nst_net* net, int H, int W, uint32_t tap_mask, uint32_t style_mask,
                               uint32_t content_mask, int with_grad) {
  return plan_create_impl(out, net, H, W, tap_mask, style_mask, content_mask, with_grad, 1, nullptr, 0);
}

// batch > 1: a head (shared tensors x batch, batched parameter sets, no optimizer state of its own);
// head != nullptr: member `index` of that head (shared tensors are slices of the head's)
static int plan_create_impl(nst_plan** out, const nst_net* net, int H, int W, uint32_t tap_mask, uint32_t style_mask,
                            uint32_t content_mask, int with_grad, int batch, nst_plan* head, int index) {
  if (!out || !net || H < 1 || W < 1) return fail(NST_ERR_ARG, "nst_plan_create: bad arguments");
  CKI(nst_device_check());
  tap_mask |= style_mask | content_mask;
  if (tap_mask == 0 || (tap_mask >> net->n_conv) != 0)
    return fail(NST_ERR_ARG, "nst_plan_create: tap mask 0x%x outside the %d convolutions of the net", tap_mask, net->n_conv);
  nst_plan* p = new nst_plan();
  p->net = net;
  p->H = H;
  p->W = W;
  p->tap_mask = tap_mask;
  p->style_mask = style_mask;
  p->content_mask = content_mask;
  p->with_grad = with_grad ? 1 : 0;
  p->batch = batch;
  p->head = head;
  p->index = index;
  // tensor shared between a head and its members: the head allocates it for all images, a member takes its slice
  auto shared = [&](auto** ptr, auto* head_field, size_t count, bool zero) -> int {
    if (head != nullptr) {
      *ptr = head_field + static_cast<size_t>(index) * count;
      return NST_OK;
    }
    return plan_alloc_t(p, ptr, count * static_cast<size_t>(batch), zero);
  };
  int n_layers = 0;
  for (int i = 0; i < NST_MAX_CONV; ++i)
    if (tap_mask & (1u << i)) n_layers = i + 1;
  p->n_layers = n_layers;
  p->lh[0] = H;
  p->lw[0] = W;
  for (int l = 1; l < 5; ++l) {
    p->lh[l] = p->lh[l - 1] / 2;  // MaxPool2d(2, 2), ceil_mode=False
    p->lw[l] = p->lw[l - 1] / 2;
  }
  if (p->lh[kLevel[n_layers - 1]] < 1 || p->lw[kLevel[n_layers - 1]] < 1) {
    delete p;
    return fail(NST_ERR_ARG, "nst_plan_create: %dx%d is too small for conv index %d", H, W, n_layers - 1);
  }
#define PA(call)            \
  do {                      \
    int r__ = (call);       \
    if (r__ != NST_OK) {    \
      nst_plan_destroy(p);  \
      return r__;           \
    }                       \
  } while (0)
  for (int i = 0; i < n_layers; ++i) {
    if (style_mask & (1u << i)) {
      if (p->n_style == GRAM_MAX_LAYERS) {
        nst_plan_destroy(p);
        return fail(NST_ERR_UNSUPPORTED, "at most %d style layers", GRAM_MAX_LAYERS);
      }
      p->style_conv[p->n_style++] = i;
    }
    if (content_mask & (1u << i)) p->content_conv[p->n_content++] = i;
  }
  if (with_grad) {
    const uint32_t tgt = style_mask | content_mask;
    if (!(tgt & (1u << (n_layers - 1)))) {
      nst_plan_destroy(p);
      return fail(NST_ERR_ARG, "with_grad: the deepest tapped conv must carry a style or content target");
    }
    for (int i = 0; i < n_layers - 1; ++i)
      if ((tgt & (1u << i)) && kPoolAfter[i]) {
        nst_plan_destroy(p);
        return fail(NST_ERR_UNSUPPORTED, "targets on a conv that is followed by a max-pool (conv index %d)", i);
      }
  }
  // ---- activations
  for (int i = 0; i < n_layers; ++i) {
    const int lv = kLevel[i];
    const size_t pix = static_cast<size_t>(p->lh[lv]) * p->lw[lv];
    const int C = kCout[i];
    const bool last = i == n_layers - 1;
    const bool pooled = kPoolAfter[i] && !last;
    if (tap_mask & (1u << i)) PA(shared(&p->tap[i], head ? head->tap[i] : nullptr, pix * C, false));
    if (!last) {
      const size_t opix = pooled ? static_cast<size_t>(p->lh[lv + 1]) * p->lw[lv + 1] : pix;
      PA(shared(&p->act[i], head ? head->act[i] : nullptr, opix * C, false));
      if (pooled) PA(shared(&p->route[i], head ? head->route[i] : nullptr, opix * C, false));
    }
    if (with_grad) {
      PA(shared(&p->gpre[i], head ? head->gpre[i] : nullptr, pix * C, true));
      if (((style_mask | content_mask) & (1u << i)) && !last) PA(shared(&p->gadd[i], head ? head->gadd[i] : nullptr, pix * C, true));
    }
  }
  // ---- style / content / pixel state
  PA(shared(&p->alpha, head ? head->alpha : nullptr, GRAM_MAX_LAYERS, true));
  PA(shared(&p->dh_scale, head ? head->dh_scale : nullptr, GRAM_MAX_LAYERS, true));
  {
    const float ones[GRAM_MAX_LAYERS] = {1.f, 1.f, 1.f, 1.f, 1.f};   // until nst_plan_set_style_target derives it from the target
    for (int b = 0; b < (head ? 1 : batch); ++b)
      if (cudaMemcpy(p->dh_scale + b * GRAM_MAX_LAYERS, ones, sizeof(ones), cudaMemcpyHostToDevice) != cudaSuccess) {
        nst_plan_destroy(p);
        return fail(NST_ERR_CUDA, "cudaMemcpy (operand scales)");
      }
  }
  PA(plan_alloc_t(p, &p->losses, NST_LOSS_COUNT, true));
  for (int l = 0; l < p->n_style; ++l) {
    const size_t cc = static_cast<size_t>(kCout[p->style_conv[l]]) * kCout[p->style_conv[l]];
    PA(shared(&p->style_target[l], head ? head->style_target[l] : nullptr, cc, true));
    PA(shared(&p->dh[l], head ? head->dh[l] : nullptr, cc, true));
  }
  int cpart = 0;
  for (int l = 0; l < p->n_content; ++l) {
    const int i = p->content_conv[l];
    const size_t numel = static_cast<size_t>(p->lh[kLevel[i]]) * p->lw[kLevel[i]] * kCout[i];
    PA(plan_alloc_t(p, &p->content_target[l], numel, true));
    p->content_part_off[l] = cpart;
    cpart += content_blocks(numel);
  }
  p->content_part_off[p->n_content] = cpart;
  PA(plan_alloc_t(p, &p->content_part, cpart > 0 ? cpart : 1, true));
  const size_t n = static_cast<size_t>(3) * H * W;
  PA(plan_alloc_t(p, &p->tedge, static_cast<size_t>(2) * H * W, true));
  PA(plan_alloc_t(p, &p->grad_pix, n, true));
  PA(plan_alloc_t(p, &p->tv_part, pixel_blocks(H, W), true));
  PA(plan_alloc_t(p, &p->edge_part, pixel_blocks(H, W), true));
  PA(build_gram_params(p));
  {
    GramParams* sets[2] = {&p->gram_shallow, &p->gram_deep};
    for (int k = 0; k < 2; ++k) {
      if (sets[k]->num_layers == 0) continue;
      // the side-stream set shares the SMs with the forward convolutions of the deeper layers: plan it for all
      // SMs anyway (its CTAs run wherever a conv tile is not resident); a head's item list runs once per image
      const int share = g_num_sms / batch > 0 ? g_num_sms / batch : 1;
      const size_t ws = gram_plan(*sets[k], share);
      float* wsp = nullptr;
      float* fin = nullptr;
      PA(plan_alloc_t(p, &wsp, ws * batch));
      PA(plan_alloc_t(p, &fin, static_cast<size_t>(sets[k]->num_fin_blocks) * batch, true));
      sets[k]->ws = wsp;
      sets[k]->fin_part = fin;
      sets[k]->batch = batch;
      sets[k]->ws_img = ws;
      sets[k]->scal_stride = GRAM_MAX_LAYERS;
      // where the loss assembly finds each style layer's per-block sums
      for (int j = 0; j < sets[k]->num_layers; ++j) {
        const GramLayer& L = sets[k]->L[j];
        int l = 0;
        for (int q = 0; q < p->n_style; ++q)
          if (p->dh[q] == L.dh) l = q;
        p->style_fin[l] = fin + L.fin_blk0;
        p->style_fin_n[l] = L.fused ? L.pairs : L.fin_blocks;
        p->style_fin_stride[l] = sets[k]->num_fin_blocks;
      }
    }
  }
  {
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    cudaError_t e = cudaStreamCreateWithPriority(&p->side, cudaStreamNonBlocking, prio_least);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&p->side2, cudaStreamNonBlocking, prio_least);
    for (int k = 0; k < 8 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&p->ev[k], cudaEventDisableTiming);
    if (e != cudaSuccess) {
      nst_plan_destroy(p);
      return fail(NST_ERR_CUDA, "side stream / events: %s", cudaGetErrorString(e));
    }
    if (getenv("NST_NO_SIDE_STREAM") != nullptr) {
      cudaStreamDestroy(p->side);
      p->side = nullptr;
      if (p->side2) cudaStreamDestroy(p->side2);
      p->side2 = nullptr;
    }
  }
  // ---- optimizer (a head has none: its members optimise)
  if (with_grad && batch == 1) {
    LbfgsBuffers& b = p->lb;
    b.n_pad = static_cast<int>((n + 3) / 4 * 4);
    lbfgs_plan(b, g_num_sms);
    PA(plan_alloc_t(p, &b.x, b.n_pad, true));
    PA(plan_alloc_t(p, &b.g, b.n_pad, true));
    PA(plan_alloc_t(p, &b.g_prev, b.n_pad, true));
    PA(plan_alloc_t(p, &b.d, b.n_pad, true));
    PA(plan_alloc_t(p, &b.hist, lbfgs_hist_floats(b), true));
    PA(plan_alloc_t(p, &b.part, static_cast<size_t>(b.nblocks) * LB_PART_STRIDE, true));
    PA(plan_alloc_t(p, &b.td_part, b.nblocks, true));
    PA(plan_alloc_t(p, &b.dots, NST_LBFGS_SLOTS * NST_LBFGS_NDOT, true));
    PA(plan_alloc_t(p, &b.scal, NST_LBFGS_NSCAL, true));
    PA(plan_alloc_t(p, &b.R, NST_CTL_MAT_DOUBLES, true));
    PA(plan_alloc_t(p, &b.YY, NST_CTL_MAT_DOUBLES, true));
    PA(plan_alloc_t(p, &b.ctl, 1, true));
    b.eval_loss = p->losses;
  }
  PA(build_conv_params(p));
  PA(build_gram_narrow(p));
#undef PA
  *out = p;
  return NST_OK;
}

extern "C" int nst_plan_set_norm(nst_plan* p, const float mean[3], const float std[3]) {
  if (!p || !mean || !std) return fail(NST_ERR_ARG, "nst_plan_set_norm: bad arguments");
  for (int c = 0; c < 3; ++c) {
    p->pc.mean[c] = mean[c];
    p->pc.stdv[c] = std[c];
  }
  drop_graph(p);
  for (nst_plan* m : p->members) CKI(nst_plan_set_norm(m, mean, std));
  return NST_OK;
}

// Plans that step at the same time on one GPU (several frames in flight, the planes of the depth variant): their CTA-pair
// launches must not be chained by programmatic dependent launch (conv_tc.cu, launch_one_t).  nst_run_frames_host sets this
// itself; a caller that steps several plans from several host threads / streams sets it on each of them.
extern "C" int nst_plan_set_shared_gpu(nst_plan* p, int shared) {
  if (!p) return fail(NST_ERR_ARG, "nst_plan_set_shared_gpu: null plan");
  const bool want = shared != 0;
  if (p->shared_gpu == want) return NST_OK;
  p->shared_gpu = want;
  for (int i = 0; i < NST_MAX_CONV; ++i) p->fwd[i].no_pdl_pair = p->dgrad[i].no_pdl_pair = p->scale[i].no_pdl_pair = want ? 1 : 0;
  drop_graph(p);   // the launch attributes are baked into the captured step
  for (nst_plan* m : p->members) CKI(nst_plan_set_shared_gpu(m, shared));
  return NST_OK;
}

extern "C" int nst_plan_set_weights(nst_plan* p, float w_style, float w_content, float w_tv, float w_edge) {
  if (!p) return fail(NST_ERR_ARG, "nst_plan_set_weights: null plan");
  for (nst_plan* m : p->members) CKI(nst_plan_set_weights(m, w_style, w_content, w_tv, w_edge));
  p->w_style = w_style;
  p->w_content = w_content;
  p->w_tv = w_tv;
  p->w_edge = w_edge;
  GramParams* sets[3] = {&p->gram_shallow, &p->gram_deep, &p->gram_shallow_narrow};
  for (int k = 0; k < 3; ++k)
    for (int l = 0; l < sets[k]->num_layers; ++l) {
      GramLayer& L = sets[k]->L[l];
      L.grad_coef = 4.f * w_style / (static_cast<float>(p->n_style) * static_cast<float>(L.C) * static_cast<float>(L.C) *
                                     static_cast<float>(L.C) * static_cast<float>(L.HW));
    }
  drop_graph(p);
  return NST_OK;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// conv1_1 forward: tcgen05 with fp16 high/low operand splitting (conv1_tc.cu), or the fp32 CUDA-core kernel (pixel.cu)
// when the TMA output maps are switched off (NST_DIRECT_STORES / NST_CONV1_CUDA_CORES)
static cudaError_t conv1_forward(nst_plan* p, const float* x, cudaStream_t s) {
  if (p->fwd[0].tma_out) return launch_conv1_tc(p->fwd[0], x, p->net->w32[0], p->pc, g_conv_sms, s);
  return launch_conv1_fwd(x, p->net->w32[0], p->net->b32[0], p->tap[0], p->act[0], p->H, p->W, p->pc, s);
}

static int forward_enqueue(nst_plan* p, const float* x, cudaStream_t s) {
  const nst_net* net = p->net;
  CK(conv1_forward(p, x, s));
  for (int i = 1; i < p->n_layers; ++i) CK(launch_conv_tc(p->fwd[i], CONV_FWD, g_conv_sms, s));
  return NST_OK;
}

extern "C" int nst_plan_features(nst_plan* p, const float* x, void* stream) {
  if (!p || !x) return fail(NST_ERR_ARG, "nst_plan_features: bad arguments");
  return forward_enqueue(p, x, static_cast<cudaStream_t>(stream));
}

extern "C" int nst_plan_tap_shape(const nst_plan* p, int conv, int* C, int* H, int* W) {
  if (!p || conv < 0 || conv >= p->n_layers || !(p->tap_mask & (1u << conv))) return fail(NST_ERR_ARG, "conv %d is not tapped", conv);
  if (C) *C = kCout[conv];
  if (H) *H = p->lh[kLevel[conv]];
  if (W) *W = p->lw[kLevel[conv]];
  return NST_OK;
}

extern "C" int nst_plan_get_tap(nst_plan* p, int conv, float* out, void* stream) {
  int C, H, W;
  CKI(nst_plan_tap_shape(p, conv, &C, &H, &W));
  if (!out) return fail(NST_ERR_ARG, "nst_plan_get_tap: null output");
  CK(launch_nhwc_half_to_nchw_float(p->tap[conv], out, H, W, C, static_cast<cudaStream_t>(stream)));
  return NST_OK;
}

// Gram of an NHWC fp16 [HW, C] matrix (C multiple of 64) -> [C, C] fp32, normalised by 1/(c_true * HW)
static int gram_of(const __half* feat, int HW, int C, int c_true, float* out, cudaStream_t s) {
  GramParams g;
  memset(&g, 0, sizeof(g));
  g.num_layers = 1;
  GramLayer& L = g.L[0];
  L.C = C;
  L.HW = HW;
  L.inv_norm = 1.f / (static_cast<float>(c_true) * static_cast<float>(HW));
  L.gram_out = out;
  if (make_tmap_feat(&g.tm[0], feat, HW, C) != 0) return fail(NST_ERR_CUDA, "tensor map (gram)");
  const size_t ws = gram_plan(g, g_num_sms);
  float* wsp = nullptr;
  float* fin = nullptr;
  CK(cudaMalloc(&wsp, ws * sizeof(float)));
  cudaError_t e = cudaMalloc(&fin, static_cast<size_t>(g.num_fin_blocks) * sizeof(float));
  if (e != cudaSuccess) {
    cudaFree(wsp);
    return fail(NST_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
  }
  g.ws = wsp;
  g.fin_part = fin;
  e = launch_gram(g, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(wsp);
  cudaFree(fin);
  if (e != cudaSuccess) return fail(NST_ERR_CUDA, "gram: %s", cudaGetErrorString(e));
  return NST_OK;
}

extern "C" int nst_plan_tap_gram(nst_plan* p, int conv, float* out, void* stream) {
  int C, H, W;
  CKI(nst_plan_tap_shape(p, conv, &C, &H, &W));
  if (!out) return fail(NST_ERR_ARG, "nst_plan_tap_gram: null output");
  return gram_of(p->tap[conv], H * W, C, C, out, static_cast<cudaStream_t>(stream));
}

extern "C" int nst_gram_chw(const float* x, int C, int H, int W, float* out, void* stream) {
  if (!x || !out || C < 1 || H < 1 || W < 1) return fail(NST_ERR_ARG, "nst_gram_chw: bad arguments");
  CKI(nst_device_check());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int Cp = (C + 63) / 64 * 64;
  const size_t HW = static_cast<size_t>(H) * W;
  __half* feat = nullptr;
  float* tmp = nullptr;
  CK(cudaMalloc(&feat, HW * Cp * sizeof(__half)));
  cudaError_t e = cudaMemsetAsync(feat, 0, HW * Cp * sizeof(__half), s);
  // rows of the padded matrix have stride Cp: convert plane by plane through the generic transpose
  if (e == cudaSuccess && Cp != C) e = cudaMalloc(&tmp, static_cast<size_t>(Cp) * Cp * sizeof(float));
  int rc = NST_OK;
  if (e == cudaSuccess) {
    if (Cp == C) {
      e = launch_nchw_float_to_nhwc_half(x, feat, H, W, C, s);
      if (e == cudaSuccess) rc = gram_of(feat, static_cast<int>(HW), Cp, C, out, s);
    } else {
      e = launch_nchw_float_to_nhwc_half_padded(x, feat, H, W, C, Cp, s);
      if (e == cudaSuccess) rc = gram_of(feat, static_cast<int>(HW), Cp, C, tmp, s);
      if (e == cudaSuccess && rc == NST_OK)
        e = cudaMemcpy2DAsync(out, static_cast<size_t>(C) * sizeof(float), tmp, static_cast<size_t>(Cp) * sizeof(float),
                              static_cast<size_t>(C) * sizeof(float), C, cudaMemcpyDeviceToDevice, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
  }
  cudaFree(feat);
  cudaFree(tmp);
  if (e != cudaSuccess) return fail(NST_ERR_CUDA, "nst_gram_chw: %s", cudaGetErrorString(e));
  return rc;
}

static int mix_common(nst_plan* a, nst_plan* b, int conv, float wb, __half** mixed, int* Ho, int* Wo, int* Cc,
                      cudaStream_t s) {
  int Ca, Ha, Wa, Cb, Hb, Wb;
  CKI(nst_plan_tap_shape(a, conv, &Ca, &Ha, &Wa));
  CKI(nst_plan_tap_shape(b, conv, &Cb, &Hb, &Wb));
  // StyleMixer.py:31-32: np.array(shape_a) + np.array(shape_b) // 2  (operator precedence: Ha + Hb // 2)
  *Ho = Ha + Hb / 2;
  *Wo = Wa + Wb / 2;
  *Cc = Ca;
  CK(cudaMalloc(mixed, static_cast<size_t>(*Ho) * *Wo * Ca * sizeof(__half)));
  cudaError_t e = launch_style_mix(a->tap[conv], Ha, Wa, b->tap[conv], Hb, Wb, *mixed, *Ho, *Wo, Ca, wb, s);
  if (e != cudaSuccess) {
    cudaFree(*mixed);
    *mixed = nullptr;
    return fail(NST_ERR_CUDA, "style mix: %s", cudaGetErrorString(e));
  }
  return NST_OK;
}

extern "C" int nst_style_mix_gram(nst_plan* a, nst_plan* b, int conv, float weight_b, float* out, void* stream) {
  if (!a || !b || !out) return fail(NST_ERR_ARG, "nst_style_mix_gram: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __half* mixed = nullptr;
  int Ho, Wo, C;
  CKI(mix_common(a, b, conv, weight_b, &mixed, &Ho, &Wo, &C, s));
  const int rc = gram_of(mixed, Ho * Wo, C, C, out, s);
  cudaFree(mixed);
  return rc;
}

extern "C" int nst_style_mix_chw(nst_plan* a, nst_plan* b, int conv, float weight_b, float* out_chw, void* stream) {
  if (!a || !b || !out_chw) return fail(NST_ERR_ARG, "nst_style_mix_chw: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __half* mixed = nullptr;
  int Ho, Wo, C;
  CKI(mix_common(a, b, conv, weight_b, &mixed, &Ho, &Wo, &C, s));
  cudaError_t e = launch_nhwc_half_to_nchw_float(mixed, out_chw, Ho, Wo, C, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(mixed);
  if (e != cudaSuccess) return fail(NST_ERR_CUDA, "nst_style_mix_chw: %s", cudaGetErrorString(e));
  return NST_OK;
}

// ------------------------------------------------------------------------------------------------
// targets
// ------------------------------------------------------------------------------------------------
extern "C" int nst_plan_set_style_target(nst_plan* p, int conv, const float* gram, void* stream) {
  if (!p || !gram) return fail(NST_ERR_ARG, "nst_plan_set_style_target: bad arguments");
  const int l = style_index(p, conv);
  if (l < 0) return fail(NST_ERR_ARG, "conv %d is not a style layer of this plan", conv);
  const size_t cc = static_cast<size_t>(kCout[conv]) * kCout[conv];
  CK(cudaMemcpyAsync(p->style_target[l], gram, cc * sizeof(float), cudaMemcpyDeviceToDevice,
                     static_cast<cudaStream_t>(stream)));
  // scale of the fp16 backward operand (G - T) * s: a function of the target alone (gram.cuh)
  CK(launch_gram_target_scale(p->style_target[l], static_cast<int>(cc), p->dh_scale + l, static_cast<cudaStream_t>(stream)));
  return NST_OK;
}

extern "C" int nst_plan_set_content_target(nst_plan* p, int conv, nst_plan* src, const float* gate, void* stream) {
  if (!p || !src) return fail(NST_ERR_ARG, "nst_plan_set_content_target: bad arguments");
  const int l = content_index(p, conv);
  if (l < 0) return fail(NST_ERR_ARG, "conv %d is not a content layer of this plan", conv);
  int C, H, W, C2, H2, W2;
  CKI(nst_plan_tap_shape(p, conv, &C, &H, &W));
  CKI(nst_plan_tap_shape(src, conv, &C2, &H2, &W2));
  if (H != H2 || W != W2) return fail(NST_ERR_ARG, "content target resolution %dx%d != plan %dx%d", H2, W2, H, W);
  CK(launch_make_content_target(src->tap[conv], gate, p->content_target[l], static_cast<size_t>(H) * W, C,
                                static_cast<cudaStream_t>(stream)));
  return NST_OK;
}

extern "C" int nst_plan_channel_gate(nst_plan* p, int conv, const float* w1, const float* w2, int reduction,
                                     float* gate_out, void* stream) {
  int C, H, W;
  CKI(nst_plan_tap_shape(p, conv, &C, &H, &W));
  if (!w1 || !w2 || !gate_out || reduction < 1 || C % reduction != 0) return fail(NST_ERR_ARG, "nst_plan_channel_gate: bad arguments");
  if (!p->pooled_dev) CKI(plan_alloc_t(p, &p->pooled_dev, 512, true));
  CK(launch_channel_gate(p->tap[conv], w1, w2, p->pooled_dev, gate_out, static_cast<size_t>(H) * W, C, C / reduction,
                         static_cast<cudaStream_t>(stream)));
  return NST_OK;
}

extern "C" int nst_plan_set_edge_target(nst_plan* p, const float* content, void* stream) {
  if (!p || !content) return fail(NST_ERR_ARG, "nst_plan_set_edge_target: bad arguments");
  CK(launch_edge_target(content, p->tedge, p->H, p->W, p->pc, static_cast<cudaStream_t>(stream)));
  return NST_OK;
}

// ------------------------------------------------------------------------------------------------
// the closure
// ------------------------------------------------------------------------------------------------
// optional per-launch timing (nst_plan_eval_timed): an event is recorded after every launch
struct LaunchTimer {
  std::vector<cudaEvent_t> ev;
  std::vector<int> kind, layer;
  cudaStream_t s = nullptr;
  // grouped: one event where the kind of launch changes (runs of same-kind launches are timed as a whole, programmatic
  // dependent launch between them intact); otherwise one event after every launch
  bool grouped = false;
  int cur = -1000;
  int push(int k, int l) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return -1;
    if (cudaEventRecord(e, s) != cudaSuccess) return -1;
    ev.push_back(e);
    kind.push_back(k);
    layer.push_back(l);
    return 0;
  }
  // after a launch
  int mark(int k, int l) {
    if (grouped) {
      if (k != NST_K_START && !layer.empty()) layer.back() += 1;  // launches in the open group
      return 0;
    }
    return push(k, l);
  }
  // before a launch
  int begin(int k) {
    if (!grouped || k == cur) return 0;
    cur = k;
    return push(k, 0);  // the event opens the group of kind k and closes the previous one
  }
  ~LaunchTimer() {
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
  }
};
#define TM(k, l)                                                                     \
  do {                                                                               \
    if (tm && tm->mark((k), (l)) != 0) return fail(NST_ERR_CUDA, "event record failed"); \
  } while (0)
#define TB(k)                                                                        \
  do {                                                                               \
    if (tm && tm->begin(k) != 0) return fail(NST_ERR_CUDA, "event record failed");    \
  } while (0)

// Enqueues one closure evaluation.  With a side stream (p->side, default) the evaluation is a small DAG: the
// critical path conv forward -> Gram of the deepest style layer -> data gradients runs on `s`; everything that
// only feeds it sideways (pixel-space losses, Gram + finalize + Gram-backward seeds of the shallower style layers,
// content loss, loss assembly) runs on the side stream and fills the SMs that layers with fewer than 148 tiles
// leave idle.  Under stream capture this becomes parallel branches of the CUDA graph.
//
// A head (p->batch > 1, nst_batch_create) enqueues the same DAG for all its members: the convolution, Gram and Gram-backward
// launches once for every image (the head's parameter sets), the per-image kernels once per member on the member's buffers
// - x, grad, counter and stop_flag are then the members' optimizer buffers and the arguments are ignored (grad != nullptr
// still selects the backward pass).
static int eval_enqueue(nst_plan* p, const float* x, float* grad, int* counter, const int* stop_flag, int* launches,
                        cudaStream_t s, LaunchTimer* tm = nullptr) {
  int nl = 0;
  const bool use_vgg = (p->w_style > 0.f && p->n_style > 0) || (p->w_content > 0.f && p->n_content > 0);
  if (grad != nullptr && !p->with_grad) return fail(NST_ERR_STATE, "plan was created without gradient buffers");
  const bool is_head = p->batch > 1;
  if (is_head && (tm != nullptr || static_cast<int>(p->members.size()) != p->batch))
    return fail(NST_ERR_STATE, "a batch head evaluates through its members and is not timed per launch");
  // the plans that own the per-image state: the plan itself, or the members of a head
  nst_plan* const self[1] = {p};
  nst_plan* const* img = is_head ? p->members.data() : self;
  const int n_img = is_head ? p->batch : 1;
  auto img_x = [&](int m) -> const float* { return is_head ? img[m]->lb.x : x; };
  auto img_grad = [&](int m) -> float* { return is_head ? img[m]->lb.g : grad; };
  TB(NST_K_START);
  TM(NST_K_START, -1);
  const bool conc = tm == nullptr && p->side != nullptr && use_vgg;
  cudaStream_t s2 = conc ? p->side : s;
  enum { EV_FORK = 0, EV_TAPS = 1, EV_CONTENT_IN = 2, EV_CONTENT = 3, EV_SEEDS = 4, EV_GRAM = 5, EV_JOIN = 6, EV_GRAM_SHALLOW = 7 };
  // record on `from`, make `to` wait
  auto edge = [&](int ev, cudaStream_t from, cudaStream_t to) -> cudaError_t {
    if (!conc) return cudaSuccess;
    cudaError_t e = cudaEventRecord(p->ev[ev], from);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(to, p->ev[ev], 0);
    return e;
  };
  const int last = p->n_layers - 1;
  const bool deep_style = use_vgg && p->n_style > 0 && p->style_conv[p->n_style - 1] == last;
  const int n_shallow = use_vgg ? p->n_style - (deep_style ? 1 : 0) : 0;
  const int max_shallow = n_shallow > 0 ? p->style_conv[n_shallow - 1] : -1;
  int max_content = -1;
  for (int l = 0; l < p->n_content; ++l)
    if (p->content_conv[l] != last && p->content_conv[l] > max_content) max_content = p->content_conv[l];

  // grouped timing: the side launches that normally sit between forward convolutions are issued after the last one, so
  // that the twelve forward launches form one uninterrupted run
  const bool grouped = tm != nullptr && tm->grouped;
  // When does the shallow layers' Gram launch go out (side stream)?  0 (default): right after its last tap (conv4_1); its
  // ~150 CTAs then push one 128-tile convolution into a second wave (conv4_4: 36 us instead of 18).  1: after the
  // second-to-last forward convolution - it then delays the deepest layer's Gram chain by as much.  2: after the deepest
  // layer's Gram launches - it then collides with a 128-tile data gradient.  Measured at 512^2: 1228 / 1230 / 1172
  // evaluations/s; the work has to overlap something, so the simplest schedule stays.
  static const int gram_when = getenv("NST_GRAM_WHEN") ? atoi(getenv("NST_GRAM_WHEN")) : 0;
  const bool can_delay = max_shallow > 0 && last - 1 > max_shallow && deep_style;
  const int late_at = !can_delay || gram_when == 0 ? max_shallow : (gram_when == 1 ? last - 1 : last + 100);
  const bool shallow_after_deep = can_delay && gram_when == 2 && !grouped;
  const int at_shallow = grouped && max_shallow > 0 ? last : late_at;
  const int at_content = grouped && max_content > 0 ? last : max_content;

  // beside the convolution chain the shallow layers' Gram launch keeps to the idle SMs; alone (timing, single stream) it
  // takes the whole GPU
  const bool narrow = conc && p->gram_shallow_narrow.num_layers > 0 && p->side2 != nullptr;
  const GramParams& gram_shallow = narrow ? p->gram_shallow_narrow : p->gram_shallow;
  cudaStream_t s3 = narrow ? p->side2 : s2;   // stream of the shallow layers' Gram launch
  // its results (fp16 operands, alpha, per-block sums) are consumed on s2 (seed launches, loss assembly) and, through
  // EV_SEEDS, on the main stream
  auto gram_shallow_launch = [&]() -> cudaError_t {
    cudaError_t e = cudaSuccess;
    if (conc) {
      e = cudaEventRecord(p->ev[EV_TAPS], s);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(s3, p->ev[EV_TAPS], 0);
    }
    if (e == cudaSuccess) e = launch_gram(gram_shallow, s3);
    if (e == cudaSuccess && narrow) e = cudaEventRecord(p->ev[EV_GRAM_SHALLOW], s3);
    return e;
  };
  bool shallow_joined = !narrow;   // s2 has waited for the shallow Gram launch
  auto content_launch = [&](int l, int accumulate, cudaStream_t st) -> cudaError_t {
    const int i = p->content_conv[l];
    const size_t numel = static_cast<size_t>(p->lh[kLevel[i]]) * p->lw[kLevel[i]] * kCout[i];
    const float gcoef = 2.f * p->w_content / (static_cast<float>(numel) * static_cast<float>(p->n_content));
    for (int m = 0; m < n_img; ++m) {
      nst_plan* q = img[m];
      // seed of the backward pass: d/dF of w_c/n_content * mean((F - Fc)^2).  A conv that is both a style and a
      // content layer gets its Gram seed first, then this one in accumulate mode.
      __nv_bfloat16* seed = nullptr;
      if (grad != nullptr && (accumulate || style_index(p, i) < 0)) seed = i == last ? q->gpre[i] : q->gadd[i];
      cudaError_t e = launch_content_loss(q->tap[i], q->content_target[l], seed, q->content_part + q->content_part_off[l], numel,
                                          gcoef, accumulate, st);
      if (e != cudaSuccess) return e;
    }
    nl += n_img - 1;   // the callers count one launch
    return cudaSuccess;
  };

  // ---- side: pixel-space terms
  CK(edge(EV_FORK, s, s2));
  TB(NST_K_PIXEL);
  for (int m = 0; m < n_img; ++m) {
    nst_plan* q = img[m];
    CK(launch_pixel_losses(img_x(m), q->tedge, q->grad_pix, q->tv_part, q->edge_part, p->H, p->W, p->pc, p->w_tv, p->w_edge, s2));
    ++nl;
  }
  TM(NST_K_PIXEL, -1);
  if (use_vgg) {
    // ---- main: VGG forward (conv1_1 reads the fp32 planar image of each member)
    TB(NST_K_CONV1_FWD);
    for (int m = 0; m < n_img; ++m) CK(conv1_forward(img[m], img_x(m), s));
    nl += n_img - 1;
    TM(NST_K_CONV1_FWD, 0);
    if (max_content == 0) CK(edge(EV_CONTENT_IN, s, s2));
    for (int i = 1; i < p->n_layers; ++i) {
      TB(NST_K_CONV_FWD);
      CK(launch_conv_tc(p->fwd[i], CONV_FWD, g_conv_sms, s));
      TM(NST_K_CONV_FWD, i);
      if (i == at_shallow) {
        // ---- side: Gram, style MSE and backward operand of the shallower style layers
        TB(NST_K_GRAM);
        CK(gram_shallow_launch());
        nl += gram_launches(gram_shallow);
        TM(NST_K_GRAM, 0);
      }
      if (i == at_content) {
        CK(edge(EV_CONTENT_IN, s, s2));
        for (int l = 0; l < p->n_content; ++l) {
          if (p->content_conv[l] == last) continue;
          TB(NST_K_CONTENT);
          CK(content_launch(l, 0, s2));
          ++nl;
          TM(NST_K_CONTENT, p->content_conv[l]);
        }
        if (grad != nullptr && conc) CK(cudaEventRecord(p->ev[EV_CONTENT], s2));
      }
    }
    nl += p->n_layers;
    if (max_shallow == 0) {
      TB(NST_K_GRAM);
      CK(gram_shallow_launch());
      nl += gram_launches(gram_shallow);
      TM(NST_K_GRAM, 0);
    }
    if (max_content == 0) {
      for (int l = 0; l < p->n_content; ++l) {
        if (p->content_conv[l] == last) continue;
        CK(content_launch(l, 0, s2));
        ++nl;
      }
      if (grad != nullptr && conc) CK(cudaEventRecord(p->ev[EV_CONTENT], s2));
    }
    // ---- main: the deepest layer's targets
    if (deep_style) {
      TB(NST_K_GRAM);
      CK(launch_gram(p->gram_deep, s));
      nl += gram_launches(p->gram_deep);
      TM(NST_K_GRAM, last);
    }
    if (shallow_after_deep) {
      TB(NST_K_GRAM);
      CK(gram_shallow_launch());
      nl += gram_launches(gram_shallow);
      TM(NST_K_GRAM, 0);
    }
    for (int l = 0; l < p->n_content; ++l) {
      if (p->content_conv[l] != last) continue;
      TB(NST_K_CONTENT);
      CK(content_launch(l, 0, s));
      ++nl;
      TM(NST_K_CONTENT, last);
    }
  }
  // ---- side: Gram-backward seeds of the shallower style layers
  if (grad != nullptr && use_vgg) {
    if (!shallow_joined && n_shallow > 0) {
      CK(cudaStreamWaitEvent(s2, p->ev[EV_GRAM_SHALLOW], 0));
      shallow_joined = true;
    }
    // shallowest (largest, bandwidth-bound) first: it then overlaps the latency-bound Gram chain of the deepest layer on
    // the main stream instead of the first data gradients; all seeds are awaited together before conv4_2's data gradient
    for (int l = 0; l < n_shallow; ++l) {
      const int i = p->style_conv[l];
      if (p->seed_folded[i]) continue;  // computed inside the data gradient of conv i+1
      TB(NST_K_GRAM_BWD);
      CK(launch_conv_tc(p->scale[i], CONV_SCALE, g_conv_sms, s2));
      ++nl;
      TM(NST_K_GRAM_BWD, i);
      const int cl = content_index(p, i);
      if (cl >= 0) {
        TB(NST_K_CONTENT);
        CK(content_launch(cl, 1, s2));
        ++nl;
        TM(NST_K_CONTENT, i);
      }
    }
    if (conc) CK(cudaEventRecord(p->ev[EV_SEEDS], s2));
  }
  // ---- side: loss assembly (needs the deepest layer's Gram MSE / content partials from main)
  if (!shallow_joined && n_shallow > 0) {
    CK(cudaStreamWaitEvent(s2, p->ev[EV_GRAM_SHALLOW], 0));
    shallow_joined = true;
  }
  CK(edge(EV_GRAM, s, s2));
  TB(NST_K_ASSEMBLE);
  for (int m = 0; m < n_img; ++m) {
  nst_plan* q = img[m];
  LossAssembleArgs a;
  memset(&a, 0, sizeof(a));
  a.tv_part = q->tv_part;
  a.n_tv = pixel_blocks(p->H, p->W);
  a.edge_part = q->edge_part;
  a.n_edge = pixel_blocks(p->H, p->W);
  a.content_part = q->content_part;
  a.n_content = use_vgg ? p->content_part_off[p->n_content] : 0;
  a.num_style = use_vgg ? p->n_style : 0;
  for (int l = 0; l < p->n_style; ++l) {
    const double c = kCout[p->style_conv[l]];
    // the Gram launches are the head's: image m's per-block sums follow image m - 1's inside each set
    a.style_fin[l] = narrow ? p->style_fin_narrow[l] + static_cast<size_t>(m) * p->style_fin_stride_narrow[l]
                            : p->style_fin[l] + static_cast<size_t>(m) * p->style_fin_stride[l];
    a.style_fin_n[l] = narrow ? p->style_fin_n_narrow[l] : p->style_fin_n[l];
    a.style_inv_cc[l] = static_cast<float>(1.0 / (c * c));
  }
  a.w_style = p->w_style;
  a.w_content = p->w_content;
  a.w_tv = p->w_tv;
  a.w_edge = p->w_edge;
  a.tv_norm = 1.0 / (3.0 * p->H * p->W);
  a.edge_norm = (p->H > 2 && p->W > 2) ? 1.0 / (static_cast<double>(p->H - 2) * (p->W - 2)) : 0.0;
  // mean over the content layers of the per-layer MSEs; each layer's partial sums are already divided by its own numel
  if (p->n_content > 0) a.content_norm = 1.0 / p->n_content;
  a.out = q->losses;
  a.counter = is_head ? &q->lb.ctl->closure_calls : counter;
  a.stop_flag = is_head ? &q->lb.ctl->stop : stop_flag;
  a.trace = q->trace;
  a.trace_cap = q->trace_cap;
  CK(launch_loss_assemble(a, s2));
  ++nl;
  }
  TM(NST_K_ASSEMBLE, -1);
  if (conc) CK(cudaEventRecord(p->ev[EV_JOIN], s2));
  // ---- main: backward chain
  if (grad != nullptr) {
    if (use_vgg) {
      if (deep_style) {
        TB(NST_K_GRAM_BWD);
        CK(launch_conv_tc(p->scale[last], CONV_SCALE, g_conv_sms, s));
        ++nl;
        TM(NST_K_GRAM_BWD, last);
        const int cl = content_index(p, last);
        if (cl >= 0) {
          TB(NST_K_CONTENT);
          CK(content_launch(cl, 1, s));
          ++nl;
          TM(NST_K_CONTENT, last);
        }
      }
      for (int i = last; i >= 1; --i) {
        // conv i's data gradient adds the seed of conv i-1
        if (conc && i - 1 == max_content) CK(cudaStreamWaitEvent(s, p->ev[EV_CONTENT], 0));
        if (conc && i - 1 == max_shallow) CK(cudaStreamWaitEvent(s, p->ev[EV_SEEDS], 0));
        TB(NST_K_CONV_DGRAD);
        CK(launch_conv_tc(p->dgrad[i], CONV_DGRAD, g_conv_sms, s));
        ++nl;
        TM(NST_K_CONV_DGRAD, i);
      }
      if (conc) CK(cudaStreamWaitEvent(s, p->ev[EV_JOIN], 0));
      {
        static const bool cuda_core_conv1 = getenv("NST_CONV1_CUDA_CORES") != nullptr;
        if (cuda_core_conv1) {
          if (is_head) return fail(NST_ERR_UNSUPPORTED, "NST_CONV1_CUDA_CORES with a batch head");
          TB(NST_K_CONV1_DGRAD);
          CK(launch_conv1_dgrad(p->gpre[0], p->net->w32[0], p->grad_pix, grad, p->H, p->W, p->pc, s));
        } else {
          TB(NST_K_CONV1_DGRAD);
          for (int m = 0; m < n_img; ++m) {
            ConvParams d = img[m]->dgrad[0];
            d.out_pix = img_grad(m);
            for (int c = 0; c < 3; ++c) d.inv_std[c] = 1.f / p->pc.stdv[c];
            CK(launch_conv_tc(d, CONV_DGRAD_PIX, g_conv_sms, s));
          }
          nl += n_img - 1;
        }
      }
      ++nl;
      TM(NST_K_CONV1_DGRAD, 0);
    } else {
      for (int m = 0; m < n_img; ++m) {
        CK(cudaMemcpyAsync(img_grad(m), img[m]->grad_pix, static_cast<size_t>(3) * p->H * p->W * sizeof(float),
                           cudaMemcpyDeviceToDevice, s));
        ++nl;
      }
    }
  } else if (conc) {
    CK(cudaStreamWaitEvent(s, p->ev[EV_JOIN], 0));
  }
  if (launches) *launches += nl;
  return NST_OK;
}

extern "C" int nst_plan_eval(nst_plan* p, const float* x, float* losses, float* grad, void* stream) {
  if (!p || !x) return fail(NST_ERR_ARG, "nst_plan_eval: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CKI(eval_enqueue(p, x, grad, nullptr, nullptr, nullptr, s));
  if (losses) CK(cudaMemcpyAsync(losses, p->losses, NST_LOSS_COUNT * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return NST_OK;
}

__global__ void spin_kernel(long long cycles) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {
  }
}

static int timer_collect(LaunchTimer& tm, nst_launch_time* out, int max_out, cudaStream_t s) {
  if (tm.grouped && tm.begin(NST_K_START) != 0) return fail(NST_ERR_CUDA, "event record failed");  // closes the last group
  CK(cudaStreamSynchronize(s));
  int n = 0;
  if (tm.grouped) {
    // row = one run of same-kind launches: kind, number of launches, duration of the whole run
    for (size_t i = 0; i + 1 < tm.ev.size() && n < max_out; ++i) {
      if (tm.kind[i] == NST_K_START) continue;
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, tm.ev[i], tm.ev[i + 1]));
      out[n].kind = tm.kind[i];
      out[n].layer = tm.layer[i];
      out[n].ms = ms;
      ++n;
    }
    return n;
  }
  for (size_t i = 1; i < tm.ev.size() && n < max_out; ++i) {
    if (tm.kind[i] == NST_K_START) continue;
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, tm.ev[i - 1], tm.ev[i]));
    out[n].kind = tm.kind[i];
    out[n].layer = tm.layer[i];
    out[n].ms = ms;
    ++n;
  }
  return n;
}

extern "C" int nst_plan_eval_timed(nst_plan* p, const float* x, float* grad, nst_launch_time* out, int max_out,
                                   void* stream) {
  if (!p || !x || !out || max_out < 1) return fail(NST_ERR_ARG, "nst_plan_eval_timed: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  LaunchTimer tm;
  tm.s = s;
  // let the host run ahead of the device so that event-to-event times contain no launch gaps
  spin_kernel<<<1, 1, 0, s>>>(600000);
  CK(cudaGetLastError());
  CKI(eval_enqueue(p, x, grad, nullptr, nullptr, nullptr, s, &tm));
  return timer_collect(tm, out, max_out, s);
}

// forward declaration (defined with the L-BFGS loop below)
static int step_enqueue(nst_plan* p, int* launches, cudaStream_t s, int max_evals, LaunchTimer* tm);
#ifdef NST_INSTRUMENT
static void timeline_arm(nst_plan* p, bool on);
#endif

static int step_timed(nst_plan* p, nst_launch_time* out, int max_out, void* stream, bool grouped);
extern "C" int nst_lbfgs_step_timed(nst_plan* p, nst_launch_time* out, int max_out, void* stream) {
  return step_timed(p, out, max_out, stream, false);
}
extern "C" int nst_lbfgs_step_timed_grouped(nst_plan* p, nst_launch_time* out, int max_out, void* stream) {
  return step_timed(p, out, max_out, stream, true);
}
static int step_timed(nst_plan* p, nst_launch_time* out, int max_out, void* stream, bool grouped) {
  if (!p || !out || max_out < 64) return fail(NST_ERR_ARG, "nst_lbfgs_step_timed: bad arguments");
  if (!p->with_grad) return fail(NST_ERR_STATE, "plan was created without optimizer state");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  LaunchTimer tm;
  tm.s = s;
  tm.grouped = grouped;
  // let the host run ahead of the device so that event-to-event times contain no launch gaps
  spin_kernel<<<1, 1, 0, s>>>(3000000);
  CK(cudaGetLastError());
  int nl = 0;
  CKI(step_enqueue(p, &nl, s, 20, &tm));
  return timer_collect(tm, out, max_out, s);
}

#ifdef NST_INSTRUMENT
// Debug / tuning aid: runs one convolution launch of the plan (mode 0 forward, 1 data gradient) with phase
// timestamps of CTA 0 -> out[7] (SM clock): 0 start, 1 setup done, 2 first operands landed, 3 last MMA issued,
// 4 accumulator complete, 5 epilogue done, 6 exit.
extern "C" int nst_plan_conv_phases(nst_plan* p, int conv, int mode, long long* out7, void* stream) {
  if (!p || conv < 0 || conv >= p->n_layers || !out7) return fail(NST_ERR_ARG, "nst_plan_conv_phases: bad arguments");
  if (mode == 1 && !p->with_grad) return fail(NST_ERR_STATE, "plan has no gradient buffers");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  long long* d = nullptr;
  CK(cudaMalloc(&d, (16 + 160) * sizeof(long long)));
  cudaError_t e = cudaMemsetAsync(d, 0, (16 + 160) * sizeof(long long), s);
  ConvParams c = mode == 0 ? p->fwd[conv] : p->dgrad[conv];
  c.dbg = d;
  c.dbg_flags = getenv("NST_DBG_FLAGS") ? atoi(getenv("NST_DBG_FLAGS")) : 0;
  c.tl = reinterpret_cast<unsigned long long*>(d + 14);  // out[14], out[15]: launch span in %globaltimer ns
  const unsigned long long tl_init[2] = {~0ull, 0ull};
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + 14, tl_init, sizeof(tl_init), cudaMemcpyHostToDevice, s);
  if (mode == 1 && conv == 0) {
    c.out_pix = p->lb.g != nullptr ? p->lb.g : p->grad_pix;
    for (int k = 0; k < 3; ++k) c.inv_std[k] = 1.f / p->pc.stdv[k];
  }
  if (mode == 0 && conv == 0) {
    c.tl = nullptr;
    if (e == cudaSuccess) e = c.tma_out ? launch_conv1_tc(c, p->lb.x != nullptr ? p->lb.x : p->grad_pix, p->net->w32[0], p->pc, g_num_sms, s) : cudaErrorNotSupported;
  } else if (e == cudaSuccess) {
    e = launch_conv_tc(c, mode == 0 ? CONV_FWD : (conv == 0 ? CONV_DGRAD_PIX : CONV_DGRAD), g_num_sms, s);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out7, d, (16 + 160) * sizeof(long long), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(d);
  if (e != cudaSuccess) return fail(NST_ERR_CUDA, "nst_plan_conv_phases: %s", cudaGetErrorString(e));
  return NST_OK;
}

// Debug / tuning aid: SM clock at the phase boundaries of the last L-BFGS controller launch (slots: lbfgs_ctl.h).
extern "C" int nst_lbfgs_ctl_clocks(nst_plan* p, long long* out8, void* stream) {
  if (!p || !out8 || !p->with_grad) return fail(NST_ERR_ARG, "nst_lbfgs_ctl_clocks: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaStreamSynchronize(s));
  CK(cudaMemcpy(out8, reinterpret_cast<const char*>(p->lb.ctl) + offsetof(NstLbfgsCtl, clk), 8 * sizeof(long long), cudaMemcpyDeviceToHost));
  return NST_OK;
}

// Debug / tuning aid: where do the convolution launches sit inside the captured step?  enable = 1 gives every conv launch
// of the plan a slot {earliest CTA start, latest CTA end} in %globaltimer ns (slot = conv for forward, 16 + conv for data
// gradients, 32 + conv for Gram backward) and drops the captured graph so that the next step re-captures with the slots;
// out96 != nullptr reads the slots back (8 x 48 values: start, end, earliest return from
// griddepcontrol.wait, latest completion of a CTA's last accumulator, earliest / latest first MMA, latest last MMA issue, spare) and re-arms them.
extern "C" int nst_plan_timeline(nst_plan* p, int enable, unsigned long long* out96, void* stream) {
  if (!p) return fail(NST_ERR_ARG, "nst_plan_timeline: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (enable && !p->timeline) CKI(plan_alloc_t(p, &p->timeline, 384));
  unsigned long long init[384];
  for (int i = 0; i < 48; ++i) {
    init[8 * i] = ~0ull;      // earliest CTA start
    init[8 * i + 1] = 0ull;   // latest CTA end
    init[8 * i + 2] = ~0ull;  // earliest return from griddepcontrol.wait
    init[8 * i + 3] = 0ull;   // latest "accumulator of the CTA's last tile complete"
    init[8 * i + 4] = ~0ull;  // earliest first MMA (first patch and first weight stage have arrived)
    init[8 * i + 5] = 0ull;   // latest first MMA
    init[8 * i + 6] = 0ull;   // latest "last MMA of the CTA's last tile issued"
    init[8 * i + 7] = 0ull;
  }
  if (out96 && p->timeline) {
    CK(cudaStreamSynchronize(s));
    CK(cudaMemcpy(out96, p->timeline, sizeof(init), cudaMemcpyDeviceToHost));
  }
  if (p->timeline) CK(cudaMemcpy(p->timeline, init, sizeof(init), cudaMemcpyHostToDevice));
  p->timeline_on = enable != 0;
  if (!p->timeline_on) timeline_arm(p, false);
  drop_graph(p);
  return NST_OK;
}

#endif  // NST_INSTRUMENT

// ------------------------------------------------------------------------------------------------
// mask compositing (the step after the loop in six of the reference's call sites)
// ------------------------------------------------------------------------------------------------
extern "C" int nst_mask_composite(const uint8_t* content, const uint8_t* style, const uint8_t* mask, int H, int W, int C,
                                  int edge_smoothing, uint8_t* out, void* stream) {
  if (!content || !style || !mask || !out || H < 1 || W < 1 || C < 1 || C > 4 || edge_smoothing < 0)
    return fail(NST_ERR_ARG, "nst_mask_composite: bad arguments");
  CKI(nst_device_check());
  int k = edge_smoothing;
  if (k != 0 && (k & 1) == 0) k += 1;  // segmentation_style_transfer.py:76-77
  if (k > MASK_MAX_K) return fail(NST_ERR_UNSUPPORTED, "nst_mask_composite: edge_smoothing %d > %d", edge_smoothing, MASK_MAX_K);
  CK(launch_mask_composite(content, style, mask, out, H, W, C, k, static_cast<cudaStream_t>(stream)));
  return NST_OK;
}

extern "C" int nst_mask_gaussian_weights(int k, int* w) {
  if (!w || mask_gaussian_weights(k, w) != 0) return fail(NST_ERR_ARG, "nst_mask_gaussian_weights: k must be odd, 1..%d", MASK_MAX_K);
  return NST_OK;
}

// ------------------------------------------------------------------------------------------------
// video back end: RGB -> BGR + cross-dissolved in-between frames (app.py:800-806, 820-840)
// ------------------------------------------------------------------------------------------------
extern "C" int nst_video_assemble(const uint8_t* frames_rgb, int F, int H, int W, int n_interp, uint8_t* out_bgr, void* stream) {
  if (!frames_rgb || !out_bgr || F < 1 || H < 1 || W < 1 || n_interp < 0) return fail(NST_ERR_ARG, "nst_video_assemble: bad arguments");
  if (n_interp > VIDEO_MAX_INTERP)
    return fail(NST_ERR_UNSUPPORTED, "nst_video_assemble: %d interpolation frames > %d", n_interp, VIDEO_MAX_INTERP);
  CKI(nst_device_check());
  CK(launch_video_assemble(frames_rgb, F, static_cast<size_t>(H) * W, n_interp, out_bgr, static_cast<cudaStream_t>(stream)));
  return NST_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-plane depth style transfer: split into depth bins / merge the stylised planes (components/style_transfer_depth/util.py)
// ------------------------------------------------------------------------------------------------
static int mip_bins(MipBins* b, int n, const double* lo, const double* hi, int depth_f64, int dmin, int dmax, const char* who) {
  if (n < 1 || !lo || !hi) return fail(NST_ERR_ARG, "%s: bad bins", who);
  if (n > MIP_MAX_PLANES) return fail(NST_ERR_UNSUPPORTED, "%s: %d planes > %d", who, n, MIP_MAX_PLANES);
  if (!depth_f64 && (dmin < 0 || dmax > 255 || dmax < dmin)) return fail(NST_ERR_ARG, "%s: bad depth range", who);
  b->n = n;
  b->dmin = dmin;
  b->range = dmax - dmin;
  for (int i = 0; i < n; ++i) b->lo[i] = lo[i], b->hi[i] = hi[i];
  return NST_OK;
}

extern "C" int nst_mip_split(const uint8_t* image, const void* depth, int depth_f64, int dmin, int dmax, int H, int W, int C, int n,
                             const double* lo, const double* hi, uint8_t* out, void* stream) {
  if (!image || !depth || !out || H < 1 || W < 1 || C < 1 || C > 4) return fail(NST_ERR_ARG, "nst_mip_split: bad arguments");
  MipBins b;
  CKI(mip_bins(&b, n, lo, hi, depth_f64, dmin, dmax, "nst_mip_split"));
  CKI(nst_device_check());
  CK(launch_mip_split(image, depth, depth_f64, static_cast<size_t>(H) * W, C, b, out, static_cast<cudaStream_t>(stream)));
  return NST_OK;
}

extern "C" int nst_mip_merge(const uint8_t* planes, const void* depth, int depth_f64, int dmin, int dmax, int H, int W, int n,
                             const double* lo, const double* hi, uint8_t* out, void* stream) {
  if (!planes || !depth || !out || H < 1 || W < 1) return fail(NST_ERR_ARG, "nst_mip_merge: bad arguments");
  MipBins b;
  CKI(mip_bins(&b, n, lo, hi, depth_f64, dmin, dmax, "nst_mip_merge"));
  CKI(nst_device_check());
  CK(launch_mip_merge(planes, depth, depth_f64, static_cast<size_t>(H) * W, b, out, static_cast<cudaStream_t>(stream)));
  return NST_OK;
}

// ------------------------------------------------------------------------------------------------
// function-level mirrors
// ------------------------------------------------------------------------------------------------
__global__ void tv_edge_finish_kernel(const double* tv_part, const double* edge_part, int n, double tv_norm,
                                      double edge_norm, float* out2) {
  // single block of 32 threads: fixed-order sums
  double tv = 0.0, ed = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) {
    tv += tv_part[i];
    ed += edge_part[i];
  }
  for (int o = 16; o > 0; o >>= 1) {
    tv += __shfl_xor_sync(0xffffffffu, tv, o);
    ed += __shfl_xor_sync(0xffffffffu, ed, o);
  }
  if (threadIdx.x == 0) {
    out2[0] = static_cast<float>(tv * tv_norm);
    out2[1] = static_cast<float>(0.5 * ed * edge_norm);
  }
}

extern "C" int nst_tv_edge(const float* x, const float* edge_target, int H, int W, const float mean[3],
                           const float std[3], float* out2, void* stream) {
  if (!x || !out2 || H < 1 || W < 1 || !mean || !std) return fail(NST_ERR_ARG, "nst_tv_edge: bad arguments");
  CKI(nst_device_check());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PixelConsts pc;
  for (int c = 0; c < 3; ++c) {
    pc.mean[c] = mean[c];
    pc.stdv[c] = std[c];
  }
  const int nb = pixel_blocks(H, W);
  const size_t n = static_cast<size_t>(3) * H * W;
  float* scratch = nullptr;
  double* parts = nullptr;
  float* tz = nullptr;
  CK(cudaMalloc(&scratch, n * sizeof(float)));
  cudaError_t e = cudaMalloc(&parts, static_cast<size_t>(2) * nb * sizeof(double));
  if (e == cudaSuccess && !edge_target) {
    e = cudaMalloc(&tz, static_cast<size_t>(2) * H * W * sizeof(float));
    if (e == cudaSuccess) e = cudaMemsetAsync(tz, 0, static_cast<size_t>(2) * H * W * sizeof(float), s);
  }
  if (e == cudaSuccess)
    e = launch_pixel_losses(x, edge_target ? edge_target : tz, scratch, parts, parts + nb, H, W, pc, 1.f, 1.f, s);
  if (e == cudaSuccess) {
    const double en = (H > 2 && W > 2) ? 1.0 / (static_cast<double>(H - 2) * (W - 2)) : 0.0;
    tv_edge_finish_kernel<<<1, 32, 0, s>>>(parts, parts + nb, nb, 1.0 / (3.0 * H * W), en, out2);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(scratch);
  cudaFree(parts);
  cudaFree(tz);
  if (e != cudaSuccess) return fail(NST_ERR_CUDA, "nst_tv_edge: %s", cudaGetErrorString(e));
  return NST_OK;
}

extern "C" int nst_edge_images(const float* img, int channels, int H, int W, float* out, void* stream) {
  if (!img || !out || (channels != 1 && channels != 3)) return fail(NST_ERR_ARG, "nst_edge_images: channels must be 1 or 3");
  CKI(nst_device_check());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PixelConsts pc = {{0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}};
  if (channels == 3) {
    CK(launch_edge_target(img, out, H, W, pc, s));
    return NST_OK;
  }
  // grayscale input: replicate the plane so that the channel mean returns it
  float* tmp = nullptr;
  const size_t plane = static_cast<size_t>(H) * W;
  CK(cudaMalloc(&tmp, 3 * plane * sizeof(float)));
  cudaError_t e = cudaSuccess;
  for (int c = 0; c < 3 && e == cudaSuccess; ++c)
    e = cudaMemcpyAsync(tmp + c * plane, img, plane * sizeof(float), cudaMemcpyDeviceToDevice, s);
  if (e == cudaSuccess) e = launch_edge_target(tmp, out, H, W, pc, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(tmp);
  if (e != cudaSuccess) return fail(NST_ERR_CUDA, "nst_edge_images: %s", cudaGetErrorString(e));
  return NST_OK;
}

// ------------------------------------------------------------------------------------------------
// L-BFGS loop
// ------------------------------------------------------------------------------------------------
__global__ void clamp_copy_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = fminf(fmaxf(in[i], 0.f), 1.f);
}

__global__ void lbfgs_ctl_reset_kernel(NstLbfgsCtl* c, int trace_cap) {
  c->lr = 1.0;            // torch.optim.LBFGS defaults (lbfgs.py:246-271), run_style_transfer.py:90
  c->tol_grad = 1e-7;
  c->tol_change = 1e-9;
  c->history_size = NST_LBFGS_HISTORY;
  c->H_diag = 1.0;
  c->trace_cap = trace_cap;
}

#define NOT_A_HEAD(p, who)                                                                                                   \
  do {                                                                                                                       \
    if ((p)->batch > 1) return fail(NST_ERR_STATE, who ": a batch head has no optimizer state of its own - use its members"); \
  } while (0)

extern "C" int nst_lbfgs_init(nst_plan* p, const float* x0, int trace_capacity, void* stream) {
  if (!p || !x0) return fail(NST_ERR_ARG, "nst_lbfgs_init: bad arguments");
  if (!p->with_grad) return fail(NST_ERR_STATE, "plan was created without optimizer state");
  NOT_A_HEAD(p, "nst_lbfgs_init");
  if (p->head != nullptr && trace_capacity > p->trace_cap) drop_graph(p->head);  // the trace pointer is baked into the head's graph too
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  LbfgsBuffers& b = p->lb;
  const size_t n = static_cast<size_t>(3) * p->H * p->W;
  if (trace_capacity > p->trace_cap) {
    drop_graph(p);  // the trace pointer is baked into the captured evaluation
    CKI(plan_alloc_t(p, &p->trace, static_cast<size_t>(trace_capacity) * 5, true));
    p->trace_cap = trace_capacity;
  }
  // the control block is reset on the device (torch defaults: lr 1, tolerance_grad 1e-7, tolerance_change 1e-9, history 100):
  // no host staging, hence no host synchronisation per frame
  CK(cudaMemsetAsync(b.ctl, 0, sizeof(NstLbfgsCtl), s));
  lbfgs_ctl_reset_kernel<<<1, 1, 0, s>>>(b.ctl, p->trace_cap);
  CK(cudaGetLastError());
  CK(cudaMemsetAsync(b.R, 0, NST_CTL_MAT_DOUBLES * sizeof(double), s));
  CK(cudaMemsetAsync(b.YY, 0, NST_CTL_MAT_DOUBLES * sizeof(double), s));
  CK(cudaMemsetAsync(b.x, 0, b.n_pad * sizeof(float), s));
  CK(cudaMemsetAsync(b.g, 0, b.n_pad * sizeof(float), s));
  CK(cudaMemsetAsync(b.g_prev, 0, b.n_pad * sizeof(float), s));
  CK(cudaMemsetAsync(b.d, 0, b.n_pad * sizeof(float), s));
  // the closure clamps before its first forward (run_style_transfer.py:108-109)
  clamp_copy_kernel<<<592, 256, 0, s>>>(x0, b.x, n);
  CK(cudaGetLastError());
  return NST_OK;
}

// One optimizer.step(closure).  max_evals < 20 truncates the step after that many evaluations (used to time an
// exact number of evaluations; the optimizer state stays valid - it looks like a step that ended early).
// arms (or disarms) the launch-span slots of every convolution launch of the plan; see nst_plan_timeline
#ifdef NST_INSTRUMENT
static void timeline_arm(nst_plan* p, bool on) {
  for (int i = 0; i < p->n_layers; ++i) {
    // NST_TL_FLAGS: the wrong-result timing experiments of ConvParams::dbg_flags for the launches of the traced evaluation
    static const int tl_flags = getenv("NST_TL_FLAGS") ? atoi(getenv("NST_TL_FLAGS")) : 0;
    p->fwd[i].dbg_flags = p->dgrad[i].dbg_flags = on ? tl_flags : 0;
    p->fwd[i].tl = on ? p->timeline + 8 * i : nullptr;
    p->dgrad[i].tl = on ? p->timeline + 8 * (16 + i) : nullptr;
    p->scale[i].tl = on ? p->timeline + 8 * (32 + i) : nullptr;
  }
}
#endif

// One optimizer.step() of every member of a head: the evaluations are shared launches (eval_enqueue), each member's
// optimizer iteration (four launches, ~25 us of them serial latency) runs on the member's own side stream beside the others'.
static int step_enqueue_head(nst_plan* p, int* launches, cudaStream_t s, int max_evals) {
  int nl = 0, evals = 0;
  const int B = p->batch;
  if (static_cast<int>(p->members.size()) != B) return fail(NST_ERR_STATE, "batch head without members");
  nst_plan* m0 = p->members[0];
  for (int m = 0; m < B; ++m) {
    CK(launch_lbfgs_step_begin(p->members[m]->lb, s));
    ++nl;
  }
  CKI(eval_enqueue(p, m0->lb.x, m0->lb.g, nullptr, nullptr, &nl, s, nullptr));
  ++evals;
  const int max_iter = 20;
  for (int k = 1; k <= max_iter && evals < max_evals + (max_evals >= 20 ? 1 : 0); ++k) {
    const int mode = k == 1 ? NST_CTL_BEGIN : NST_CTL_MID;
    // fork: ev[0] is free between evaluations (EV_FORK is recorded and consumed at the start of each)
    CK(cudaEventRecord(p->ev[0], s));
    for (int m = 0; m < B; ++m) {
      nst_plan* q = p->members[m];
      cudaStream_t sm = q->side != nullptr ? q->side : s;
      if (sm != s) CK(cudaStreamWaitEvent(sm, p->ev[0], 0));
      CK(launch_lbfgs_iteration(q->lb, mode, sm));
      nl += 4;
      if (sm != s) {
        CK(cudaEventRecord(q->ev[0], sm));
        CK(cudaStreamWaitEvent(s, q->ev[0], 0));
      }
    }
    if (k != max_iter) {
      CKI(eval_enqueue(p, m0->lb.x, m0->lb.g, nullptr, nullptr, &nl, s, nullptr));
      ++evals;
    }
  }
  if (launches) *launches = nl;
  return NST_OK;
}

static int step_enqueue(nst_plan* p, int* launches, cudaStream_t s, int max_evals, LaunchTimer* tm) {
  if (p->batch > 1) {
    if (tm != nullptr) return fail(NST_ERR_UNSUPPORTED, "per-launch timing of a batch head");
    return step_enqueue_head(p, launches, s, max_evals);
  }
  LbfgsBuffers& b = p->lb;
  int nl = 0, evals = 0;
  CK(launch_lbfgs_step_begin(b, s));
  ++nl;
  CKI(eval_enqueue(p, b.x, b.g, &b.ctl->closure_calls, &b.ctl->stop, &nl, s, tm));
  ++evals;
  const int max_iter = 20;  // torch.optim.LBFGS default, run_style_transfer.py:90
  for (int k = 1; k <= max_iter && evals < max_evals + (max_evals >= 20 ? 1 : 0); ++k) {
    const int mode = k == 1 ? NST_CTL_BEGIN : NST_CTL_MID;
    if (tm == nullptr) {
      CK(launch_lbfgs_iteration(b, mode, s));
    } else {
      TB(NST_K_START);
      TM(NST_K_START, -1);
      TB(NST_K_LBFGS_PASS1);
      CK(launch_lbfgs_pass1(b, s));
      TM(NST_K_LBFGS_PASS1, -1);
      TB(NST_K_LBFGS_REDUCE);
      CK(launch_lbfgs_reduce(b, s));
      TM(NST_K_LBFGS_REDUCE, -1);
      TB(NST_K_LBFGS_CONTROL);
      CK(launch_lbfgs_control(b, mode, s));
      TM(NST_K_LBFGS_CONTROL, -1);
      TB(NST_K_LBFGS_PASS2);
      CK(launch_lbfgs_pass2(b, s));
      TM(NST_K_LBFGS_PASS2, -1);
    }
    nl += 4;
    if (k != max_iter) {
#ifdef NST_INSTRUMENT
      if (p->timeline_on) timeline_arm(p, evals == 10);  // the spans of ONE evaluation in the middle of the step
#endif
      CKI(eval_enqueue(p, b.x, b.g, &b.ctl->closure_calls, &b.ctl->stop, &nl, s, tm));  // lbfgs.py:493-502
      ++evals;
    }
  }
  if (launches) *launches = nl;
  return NST_OK;
}

extern "C" int nst_lbfgs_partial_step(nst_plan* p, int n_evals, void* stream) {
  if (!p || n_evals < 1 || n_evals > 20) return fail(NST_ERR_ARG, "nst_lbfgs_partial_step: n_evals must be 1..20");
  if (!p->with_grad) return fail(NST_ERR_STATE, "plan was created without optimizer state");
  int nl = 0;
  CKI(step_enqueue(p, &nl, static_cast<cudaStream_t>(stream), n_evals, nullptr));
  return nl;
}

// Captures and instantiates the CUDA graph of one optimizer.step() (about 800 kernel nodes) if the plan has none.
static int ensure_step_graph(nst_plan* p, cudaStream_t s) {
  if (p->step_graph) return NST_OK;
  cudaGraph_t graph = nullptr;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  const int rc = step_enqueue(p, &p->launches_per_step, s, 20, nullptr);
  cudaError_t e = cudaStreamEndCapture(s, &graph);
  if (rc != NST_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return fail(NST_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(&p->step_graph, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail(NST_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  e = cudaGraphUpload(p->step_graph, s);  // the first launch then costs what every later launch costs
  if (e != cudaSuccess) return fail(NST_ERR_CUDA, "cudaGraphUpload: %s", cudaGetErrorString(e));
  return NST_OK;
}

extern "C" int nst_lbfgs_prepare_graph(nst_plan* p, void* stream) {
  if (!p) return fail(NST_ERR_ARG, "nst_lbfgs_prepare_graph: null plan");
  if (!p->with_grad) return fail(NST_ERR_STATE, "plan was created without optimizer state");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static const bool no_graph = getenv("NST_NO_GRAPH") != nullptr;
  if (no_graph || s == nullptr) return NST_OK;  // steps are enqueued directly: nothing to prepare
  return ensure_step_graph(p, s);
}

extern "C" int nst_lbfgs_step(nst_plan* p, void* stream) {
  if (!p) return fail(NST_ERR_ARG, "nst_lbfgs_step: null plan");
  if (!p->with_grad) return fail(NST_ERR_STATE, "plan was created without optimizer state");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static const bool no_graph = getenv("NST_NO_GRAPH") != nullptr;
  if (no_graph || s == nullptr) {
    // legacy default stream cannot be captured: enqueue directly
    return step_enqueue(p, &p->launches_per_step, s, 20, nullptr);
  }
  CKI(ensure_step_graph(p, s));
  CK(cudaGraphLaunch(p->step_graph, s));
  return NST_OK;
}

extern "C" int nst_lbfgs_launches_per_step(const nst_plan* p) { return p ? p->launches_per_step : 0; }

extern "C" int nst_lbfgs_status(nst_plan* p, nst_status* out, void* stream) {
  if (!p || !out) return fail(NST_ERR_ARG, "nst_lbfgs_status: bad arguments");
  if (!p->with_grad) return fail(NST_ERR_STATE, "plan was created without optimizer state");
  NOT_A_HEAD(p, "nst_lbfgs_status");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static thread_local NstLbfgsCtl h;
  CK(cudaMemcpyAsync(&h, p->lb.ctl, sizeof(h), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(out->losses, p->losses, NST_LOSS_COUNT * sizeof(float), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  out->n_iter = h.n_iter;
  out->func_evals = h.func_evals;
  out->closure_calls = h.closure_calls;
  out->stop = h.stop;
  out->hist_len = h.hist_len;
  out->reserved = 0;
  out->loss = h.loss;
  out->prev_loss = h.prev_loss;
  out->t = h.t;
  out->H_diag = h.H_diag;
  out->gtd = h.gtd;
  out->gmax = h.gmax;
  out->max_td = h.max_td;
  return NST_OK;
}

extern "C" int nst_lbfgs_get_x(nst_plan* p, float* out, void* stream) {
  if (!p || !out) return fail(NST_ERR_ARG, "nst_lbfgs_get_x: bad arguments");
  if (!p->with_grad) return fail(NST_ERR_STATE, "plan was created without optimizer state");
  NOT_A_HEAD(p, "nst_lbfgs_get_x");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  clamp_copy_kernel<<<592, 256, 0, s>>>(p->lb.x, out, static_cast<size_t>(3) * p->H * p->W);  // run_style_transfer.py:153-155
  CK(cudaGetLastError());
  return NST_OK;
}

extern "C" int nst_lbfgs_trace(nst_plan* p, float* host_out, int max_rows, void* stream) {
  if (!p || !host_out || max_rows < 0) return fail(NST_ERR_ARG, "nst_lbfgs_trace: bad arguments");
  if (!p->with_grad) return fail(NST_ERR_STATE, "plan was created without optimizer state");
  NOT_A_HEAD(p, "nst_lbfgs_trace");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int calls = 0;
  CK(cudaMemcpyAsync(&calls, &p->lb.ctl->closure_calls, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  int rows = calls < p->trace_cap ? calls : p->trace_cap;
  if (rows > max_rows) rows = max_rows;
  if (rows > 0) {
    CK(cudaMemcpyAsync(host_out, p->trace, static_cast<size_t>(rows) * 5 * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  return rows;
}

// ------------------------------------------------------------------------------------------------
// host-buffer frame path
// ------------------------------------------------------------------------------------------------
__global__ void u8_hwc_to_f32_chw_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int H, int W) {
  const size_t HW = static_cast<size_t>(H) * W;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < HW;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // transforms.ToTensor(): uint8 -> float32 / 255   (run_style_transfer.py:5-11)
    out[i] = static_cast<float>(in[3 * i]) / 255.f;
    out[HW + i] = static_cast<float>(in[3 * i + 1]) / 255.f;
    out[2 * HW + i] = static_cast<float>(in[3 * i + 2]) / 255.f;
  }
}
__global__ void f32_chw_to_u8_hwc_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, int H, int W) {
  const size_t HW = static_cast<size_t>(H) * W;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < HW;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // transforms.ToPILImage(): clamp (already in [0,1]) . mul(255) . byte() -> truncation   (run_style_transfer.py:157)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = fminf(fmaxf(in[c * HW + i], 0.f), 1.f) * 255.f;
      out[3 * i + c] = static_cast<uint8_t>(v);
    }
  }
}

// one frame = upload + targets + optimizer reset (enqueue only) ... the step loop ... download (enqueue only)
static int frame_begin(nst_plan* p, const uint8_t* content_u8, int num_steps, int channel_attention, const float* ca_w1,
                       const float* ca_w2, cudaStream_t s) {
  if (!p || !content_u8 || num_steps < 0) return fail(NST_ERR_ARG, "nst_run_frame(s)_host: bad arguments");
  if (!p->with_grad) return fail(NST_ERR_STATE, "plan was created without optimizer state");
  if (channel_attention && (!ca_w1 || !ca_w2)) return fail(NST_ERR_ARG, "channel attention needs its two weight matrices");
  const size_t HW = static_cast<size_t>(p->H) * p->W;
  if (!p->u8_dev) CKI(plan_alloc_t(p, &p->u8_dev, 3 * HW));
  if (!p->img_dev) CKI(plan_alloc_t(p, &p->img_dev, 3 * HW));
  if (!p->gate_dev) CKI(plan_alloc_t(p, &p->gate_dev, 512, true));
  CK(cudaMemcpyAsync(p->u8_dev, content_u8, 3 * HW, cudaMemcpyHostToDevice, s));
  u8_hwc_to_f32_chw_kernel<<<592, 256, 0, s>>>(p->u8_dev, p->img_dev, p->H, p->W);
  CK(cudaGetLastError());
  // targets from the content image (run_style_transfer.py:71-80, 94-96)
  CKI(nst_plan_set_edge_target(p, p->img_dev, s));
  CKI(forward_enqueue(p, p->img_dev, s));
  for (int l = 0; l < p->n_content; ++l) {
    const int conv = p->content_conv[l];
    const float* gate = nullptr;
    if (channel_attention) {
      CKI(nst_plan_channel_gate(p, conv, ca_w1, ca_w2, 2, p->gate_dev, s));
      gate = p->gate_dev;
    }
    CKI(nst_plan_set_content_target(p, conv, p, gate, s));
  }
  const int evals = 20 * (num_steps / 20 + 1);
  CKI(nst_lbfgs_init(p, p->img_dev, evals + 32, s));
  return NST_OK;
}

static int frame_end(nst_plan* p, uint8_t* out_u8, cudaStream_t s) {
  const size_t HW = static_cast<size_t>(p->H) * p->W;
  CKI(nst_lbfgs_get_x(p, p->img_dev, s));
  f32_chw_to_u8_hwc_kernel<<<592, 256, 0, s>>>(p->img_dev, p->u8_dev, p->H, p->W);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out_u8, p->u8_dev, 3 * HW, cudaMemcpyDeviceToHost, s));
  return NST_OK;
}

extern "C" int nst_run_frame_host(nst_plan* p, const uint8_t* content_u8, uint8_t* out_u8, int num_steps,
                                  int channel_attention, const float* ca_w1, const float* ca_w2, void* stream) {
  if (!out_u8) return fail(NST_ERR_ARG, "nst_run_frame_host: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CKI(frame_begin(p, content_u8, num_steps, channel_attention, ca_w1, ca_w2, s));
  nst_status st;
  memset(&st, 0, sizeof(st));
  int guard = 0;
  while (st.closure_calls <= num_steps) {  // run_style_transfer.py:100
    CKI(nst_lbfgs_step(p, s));
    CKI(nst_lbfgs_status(p, &st, s));
    if (st.stop == NST_STOP_NONFINITE) return fail(NST_ERR_STATE, "non-finite loss at evaluation %d", st.closure_calls);
    if (++guard > num_steps + 2) break;  // every step() performs at least one evaluation
  }
  CKI(frame_end(p, out_u8, s));
  CK(cudaStreamSynchronize(s));
  return st.closure_calls;
}

// ---- several images per launch ------------------------------------------------------------------------------
extern "C" int nst_batch_create(nst_plan** out, const nst_net* net, int H, int W, uint32_t tap_mask, uint32_t style_mask,
                                uint32_t content_mask, int batch) {
  if (!out) return fail(NST_ERR_ARG, "nst_batch_create: bad arguments");
  if (batch < 2 || batch > 8) return fail(NST_ERR_ARG, "nst_batch_create: batch must be 2..8 (got %d)", batch);
  // tiles (16 x 8 pixels) and 2 x 2 pooling windows must not straddle two images at any of the five resolutions
  if (H % 16 != 0 || W % 16 != 0) return fail(NST_ERR_UNSUPPORTED, "nst_batch_create: H and W must be multiples of 16 (got %dx%d)", H, W);
  nst_plan* head = nullptr;
  CKI(plan_create_impl(&head, net, H, W, tap_mask, style_mask, content_mask, 1, batch, nullptr, 0));
  for (int m = 0; m < batch; ++m) {
    nst_plan* q = nullptr;
    const int rc = plan_create_impl(&q, net, H, W, tap_mask, style_mask, content_mask, 1, 1, head, m);
    if (rc != NST_OK) {
      nst_plan_destroy(head);
      return rc;
    }
    head->members.push_back(q);
    head->bytes += q->bytes;
  }
  *out = head;
  return NST_OK;
}

extern "C" int nst_batch_size(const nst_plan* p) { return p ? p->batch : 0; }

extern "C" nst_plan* nst_batch_member(nst_plan* head, int index) {
  if (!head || index < 0 || index >= static_cast<int>(head->members.size())) {
    fail(NST_ERR_ARG, "nst_batch_member: no member %d", index);
    return nullptr;
  }
  return head->members[index];
}

extern "C" int nst_lbfgs_freeze(nst_plan* p, int frozen, void* stream) {
  if (!p || !p->with_grad) return fail(NST_ERR_ARG, "nst_lbfgs_freeze: bad arguments");
  NOT_A_HEAD(p, "nst_lbfgs_freeze");
  CK(cudaMemsetAsync(&p->lb.ctl->frozen, frozen ? 1 : 0, sizeof(int), static_cast<cudaStream_t>(stream)));
  return NST_OK;
}

// `count` (<= batch) frames through one head: frame k on member k, the remaining members frozen.  Every optimizer.step()
// of all frames is ONE graph launch; a frame whose own loop has ended (run_style_transfer.py:100) is frozen while the others
// finish.  Each frame's trajectory is the one nst_run_frame_host gives on a single-image plan up to the order of the
// floating-point sums inside the shared launches (the tile -> CTA assignment differs).
extern "C" int nst_run_batch_host(nst_plan* head, int count, const uint8_t* const* content_u8, uint8_t* const* out_u8, int num_steps,
                                  int channel_attention, const float* ca_w1, const float* ca_w2, void* stream, int* closure_calls) {
  if (!head || head->batch < 2 || !content_u8 || !out_u8 || count < 1 || count > head->batch)
    return fail(NST_ERR_ARG, "nst_run_batch_host: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int B = head->batch;
  nst_status st[8];
  int guard[8];
  memset(st, 0, sizeof(st));
  memset(guard, 0, sizeof(guard));
  for (int k = 0; k < B; ++k) {
    nst_plan* q = head->members[k];
    if (k < count) {
      if (!content_u8[k] || !out_u8[k]) return fail(NST_ERR_ARG, "nst_run_batch_host: bad arguments");
      CKI(frame_begin(q, content_u8[k], num_steps, channel_attention, ca_w1, ca_w2, s));
    } else {
      if (!q->img_dev) CKI(plan_alloc_t(q, &q->img_dev, static_cast<size_t>(3) * q->H * q->W, true));
      CKI(nst_lbfgs_init(q, q->img_dev, 0, s));
      CKI(nst_lbfgs_freeze(q, 1, s));
    }
  }
  for (;;) {
    bool active[8], any = false;
    for (int k = 0; k < count; ++k) {
      active[k] = st[k].closure_calls <= num_steps && guard[k] <= num_steps + 2;
      any |= active[k];
    }
    if (!any) break;
    for (int k = 0; k < count; ++k)
      if (!active[k]) CKI(nst_lbfgs_freeze(head->members[k], 1, s));
    CKI(nst_lbfgs_step(head, s));
    for (int k = 0; k < count; ++k) {
      if (!active[k]) continue;
      CKI(nst_lbfgs_status(head->members[k], &st[k], s));
      if (st[k].stop == NST_STOP_NONFINITE) return fail(NST_ERR_STATE, "frame %d: non-finite loss at evaluation %d", k, st[k].closure_calls);
      ++guard[k];
    }
  }
  for (int k = 0; k < count; ++k) CKI(frame_end(head->members[k], out_u8[k], s));
  CK(cudaStreamSynchronize(s));
  for (int k = 0; k < count; ++k)
    if (closure_calls) closure_calls[k] = st[k].closure_calls;
  return NST_OK;
}

// K independent frames on K plans / streams, stepped in lock step: every optimizer.step() of every frame is enqueued
// before the host waits for the first, so the launches of one frame fill the ramps and tails of the others' (small
// frames are bound by per-launch latency: +46 % at 256 x 256 with two frames, +79 % with four; +1 % at 720p).
// Each frame's result is bit-identical to nst_run_frame_host on the same plan.
extern "C" int nst_run_frames_host(nst_plan* const* plans, int count, const uint8_t* const* content_u8, uint8_t* const* out_u8,
                                   int num_steps, int channel_attention, const float* ca_w1, const float* ca_w2, void* const* streams,
                                   int* closure_calls) {
  constexpr int MAX_FRAMES = 16;
  if (!plans || !content_u8 || !out_u8 || !streams || count < 1) return fail(NST_ERR_ARG, "nst_run_frames_host: bad arguments");
  if (count > MAX_FRAMES) return fail(NST_ERR_UNSUPPORTED, "nst_run_frames_host: %d frames > %d", count, MAX_FRAMES);
  for (int k = 0; k < count; ++k) {
    if (!plans[k] || !out_u8[k]) return fail(NST_ERR_ARG, "nst_run_frames_host: bad arguments");
    for (int j = 0; j < k; ++j)
      if (plans[j] == plans[k] || streams[j] == streams[k]) return fail(NST_ERR_ARG, "nst_run_frames_host: plans and streams must be distinct");
  }
  nst_status st[MAX_FRAMES];
  int guard[MAX_FRAMES];
  memset(st, 0, sizeof(st));
  memset(guard, 0, sizeof(guard));
#ifndef NST_EXP_NO_SHARED   // experiment builds only
  if (count > 1)
    for (int k = 0; k < count; ++k) CKI(nst_plan_set_shared_gpu(plans[k], 1));   // several chains of pair launches side by side
#endif
  for (int k = 0; k < count; ++k)
    CKI(frame_begin(plans[k], content_u8[k], num_steps, channel_attention, ca_w1, ca_w2, static_cast<cudaStream_t>(streams[k])));
  for (;;) {
    bool active[MAX_FRAMES], any = false;
    for (int k = 0; k < count; ++k) {
      active[k] = st[k].closure_calls <= num_steps && guard[k] <= num_steps + 2;  // run_style_transfer.py:100
      any |= active[k];
    }
    if (!any) break;
    for (int k = 0; k < count; ++k)
      if (active[k]) CKI(nst_lbfgs_step(plans[k], streams[k]));
    for (int k = 0; k < count; ++k) {
      if (!active[k]) continue;
      CKI(nst_lbfgs_status(plans[k], &st[k], streams[k]));
      if (st[k].stop == NST_STOP_NONFINITE) return fail(NST_ERR_STATE, "frame %d: non-finite loss at evaluation %d", k, st[k].closure_calls);
      ++guard[k];
    }
  }
  for (int k = 0; k < count; ++k) CKI(frame_end(plans[k], out_u8[k], static_cast<cudaStream_t>(streams[k])));
  for (int k = 0; k < count; ++k) {
    CK(cudaStreamSynchronize(static_cast<cudaStream_t>(streams[k])));
    if (closure_calls) closure_calls[k] = st[k].closure_calls;
  }
  return NST_OK;
}
