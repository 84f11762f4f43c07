// Bandwidth-bound pixel-space kernels: TV + edge losses and their gradients, conv1_1 forward and
// data-gradient on CUDA cores (3 channels: tensor-core hostile), content loss, loss assembly, and
// the step-0 helpers (layout conversion, weight packing, channel-attention gate, StyleMixer).
//
// Reference behaviour reproduced (all under /root/reference/multi_style_transfer/):
//   normalize                style_transfer_losses.py:9-28
//   total_variation_loss     style_transfer_losses.py:149-174   (on the NORMALISED image, run_style_transfer.py:129)
//   to_grayscale             helper_functions.py:104-113
//   get_gradient_imgs        style_transfer_losses.py:177-204
//   edge_loss                style_transfer_losses.py:207-225   (on the RAW image, run_style_transfer.py:134-136)
//   content_loss             style_transfer_losses.py:31-67
//   ChannelAttention.forward ChannelAttention.py:23-40
//   StyleMixer.mix           StyleMixer.py:25-38
#include "pixel.cuh"
#include "common.cuh"

namespace nst {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sgnf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

static constexpr int PT_W = 32, PT_H = 8;  // pixel-loss tile
int pixel_blocks(int H, int W) { return ((W + PT_W - 1) / PT_W) * ((H + PT_H - 1) / PT_H); }
static constexpr int CL_THREADS = 256;
static constexpr int CL_PER_THREAD = 8;
int content_blocks(size_t numel) {
  const size_t per = static_cast<size_t>(CL_THREADS) * CL_PER_THREAD;
  size_t b = (numel + per - 1) / per;
  if (b > 1184) b = 1184;  // 8 blocks per SM; grid-stride beyond
  return static_cast<int>(b < 1 ? 1 : b);
}

// ------------------------------------------------------------------------------------------------
// edge target
// ------------------------------------------------------------------------------------------------
__global__ void edge_target_kernel(const float* __restrict__ c, float* __restrict__ tedge, int H, int W,
                                   PixelConsts pc) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = blockIdx.y;
  if (w >= W) return;
  const size_t HW = static_cast<size_t>(H) * W;
  auto gray = [&](int hh, int ww) {
    const size_t o = static_cast<size_t>(hh) * W + ww;
    const float a = (c[o] - pc.mean[0]) / pc.stdv[0];
    const float b = (c[HW + o] - pc.mean[1]) / pc.stdv[1];
    const float d = (c[2 * HW + o] - pc.mean[2]) / pc.stdv[2];
    return ((a + b) + d) / 3.f;
  };
  float dx = 0.f, dy = 0.f;
  if (h >= 1 && h <= H - 2 && w >= 1 && w <= W - 2) {
    dx = gray(h, w + 1) - gray(h, w - 1);
    dy = gray(h + 1, w) - gray(h - 1, w);
  }
  tedge[static_cast<size_t>(h) * W + w] = dx;
  tedge[HW + static_cast<size_t>(h) * W + w] = dy;
}

cudaError_t launch_edge_target(const float* content, float* tedge, int H, int W, PixelConsts pc, cudaStream_t s) {
  dim3 grid((W + 127) / 128, H);
  edge_target_kernel<<<grid, 128, 0, s>>>(content, tedge, H, W, pc);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// TV + edge losses and gradient w.r.t. the raw image
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT_W* PT_H) pixel_losses_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ tedge,
                                                                  float* __restrict__ grad_pix,
                                                                  double* __restrict__ tv_part,
                                                                  double* __restrict__ edge_part, int H, int W,
                                                                  PixelConsts pc, float w_tv, float w_edge) {
  constexpr int SW = PT_W + 4, SH = PT_H + 4;
  __shared__ float xs[3][SH][SW + 1];
  __shared__ float gs[SH][SW + 1];
  __shared__ double red[2][PT_W * PT_H / 32];
  const int tx = threadIdx.x % PT_W, ty = threadIdx.x / PT_W;
  const int tiles_w = (W + PT_W - 1) / PT_W;
  const int w0 = (blockIdx.x % tiles_w) * PT_W, h0 = (blockIdx.x / tiles_w) * PT_H;
  const size_t HW = static_cast<size_t>(H) * W;
  for (int i = threadIdx.x; i < SH * SW; i += PT_W * PT_H) {
    const int sy = i / SW, sx = i - sy * SW;
    const int hh = h0 + sy - 2, ww = w0 + sx - 2;
    float a = 0.f, b = 0.f, c = 0.f;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
      const size_t o = static_cast<size_t>(hh) * W + ww;
      a = x[o];
      b = x[HW + o];
      c = x[2 * HW + o];
    }
    xs[0][sy][sx] = a;
    xs[1][sy][sx] = b;
    xs[2][sy][sx] = c;
    gs[sy][sx] = ((a + b) + c) / 3.f;  // helper_functions.py:113 (mean over the channel dim)
  }
  __syncthreads();
  const int h = h0 + ty, w = w0 + tx;
  const int sy = ty + 2, sx = tx + 2;
  float tv = 0.f, ed = 0.f;
  if (h < H && w < W) {
    const float ctv = w_tv / (3.f * static_cast<float>(H) * static_cast<float>(W));
    float gr[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float m = pc.mean[c], sd = pc.stdv[c];
      const float y = (xs[c][sy][sx] - m) / sd;
      float gsum = 0.f;
      if (h > 0) gsum += sgnf(y - (xs[c][sy - 1][sx] - m) / sd);
      if (w > 0) gsum += sgnf(y - (xs[c][sy][sx - 1] - m) / sd);
      if (h < H - 1) {
        const float dlt = (xs[c][sy + 1][sx] - m) / sd - y;
        tv += fabsf(dlt);
        gsum -= sgnf(dlt);
      }
      if (w < W - 1) {
        const float dlt = (xs[c][sy][sx + 1] - m) / sd - y;
        tv += fabsf(dlt);
        gsum -= sgnf(dlt);
      }
      gr[c] = ctv * gsum / sd;
    }
    if (H > 2 && W > 2) {
      const float cedge = w_edge / (static_cast<float>(H - 2) * static_cast<float>(W - 2));  // 0.5 * 2 (d/dd of (d-t)^2)
      const float* tdx = tedge;
      const float* tdy = tedge + HW;
      auto interior = [&](int hh, int ww) { return hh >= 1 && hh <= H - 2 && ww >= 1 && ww <= W - 2; };
      // residuals of the four difference pixels that contain g(h, w)
      float gg = 0.f;
      if (interior(h, w - 1)) gg += (gs[sy][sx] - gs[sy][sx - 2]) - tdx[static_cast<size_t>(h) * W + (w - 1)];
      if (interior(h, w + 1)) gg -= (gs[sy][sx + 2] - gs[sy][sx]) - tdx[static_cast<size_t>(h) * W + (w + 1)];
      if (interior(h - 1, w)) gg += (gs[sy][sx] - gs[sy - 2][sx]) - tdy[static_cast<size_t>(h - 1) * W + w];
      if (interior(h + 1, w)) gg -= (gs[sy + 2][sx] - gs[sy][sx]) - tdy[static_cast<size_t>(h + 1) * W + w];
      const float ge = cedge * gg / 3.f;
      gr[0] += ge;
      gr[1] += ge;
      gr[2] += ge;
      if (interior(h, w)) {
        const float ex = (gs[sy][sx + 1] - gs[sy][sx - 1]) - tdx[static_cast<size_t>(h) * W + w];
        const float ey = (gs[sy + 1][sx] - gs[sy - 1][sx]) - tdy[static_cast<size_t>(h) * W + w];
        ed = ex * ex + ey * ey;
      }
    }
    const size_t o = static_cast<size_t>(h) * W + w;
    grad_pix[o] = gr[0];
    grad_pix[HW + o] = gr[1];
    grad_pix[2 * HW + o] = gr[2];
  }
  double tvd = warp_sum(static_cast<double>(tv));
  double edd = warp_sum(static_cast<double>(ed));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    red[0][warp] = tvd;
    red[1][warp] = edd;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int i = 0; i < PT_W * PT_H / 32; ++i) {
      a += red[0][i];
      b += red[1][i];
    }
    tv_part[blockIdx.x] = a;
    edge_part[blockIdx.x] = b;
  }
}

cudaError_t launch_pixel_losses(const float* x, const float* tedge, float* grad_pix, double* tv_part,
                                double* edge_part, int H, int W, PixelConsts pc, float w_tv, float w_edge,
                                cudaStream_t s) {
  pixel_losses_kernel<<<pixel_blocks(H, W), PT_W * PT_H, 0, s>>>(x, tedge, grad_pix, tv_part, edge_part, H, W, pc,
                                                                w_tv, w_edge);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// conv1_1 forward: 3 -> 64 channels, fp32 on CUDA cores, fused with normalize()
//   thread = (pixel column, 8-channel group); every thread walks the 8 rows of a 32 x 8 tile
// ------------------------------------------------------------------------------------------------
static constexpr int C1_TW = 32, C1_TH = 8;

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Persistent: the grid is two blocks per SM and every block walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The
// weights are staged once per block, and the next tile's 3 x 10 x 34 input patch is fetched into registers before the
// current tile is computed - the one-launch-per-tile version spent half of every block waiting for its own prologue
// (ncu: fma pipe 39 % active, profiles/r01b_ncu_full_misc_summary.csv).
static constexpr int C1_IN = 3 * (C1_TH + 2) * (C1_TW + 2);   // 1020 patch values
static constexpr int C1_IN_PER_THREAD = (C1_IN + 255) / 256;  // 4

__global__ void __launch_bounds__(256) conv1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wgt,
                                                        const float* __restrict__ bias, __half* __restrict__ out_tap,
                                                        __half* __restrict__ out_act, int H, int W, PixelConsts pc) {
  __shared__ float in[2][3][C1_TH + 2][C1_TW + 2];
  // [c*9 + r*3 + s][half][channel group j][4]: the eight channel groups of a quarter-warp read 128 contiguous bytes
  __shared__ __align__(16) float ws[27][2][8][4];
  const int tiles_w = (W + C1_TW - 1) / C1_TW;
  const int tiles = tiles_w * ((H + C1_TH - 1) / C1_TH);
  const size_t HW = static_cast<size_t>(H) * W;
  for (int i = threadIdx.x; i < 27 * 64; i += 256) {
    const int k = i / 64, n = i - k * 64;  // torch layout [n][c][r][s] -> k = c*9 + r*3 + s
    ws[k][(n >> 2) & 1][n >> 3][n & 3] = __ldg(wgt + n * 27 + k);
  }
  const float inv_std[3] = {1.f / pc.stdv[0], 1.f / pc.stdv[1], 1.f / pc.stdv[2]};
  // patch element e of a tile -> normalised value (zero padding lives in the normalised domain)
  auto fetch = [&](int tile, float (&v)[C1_IN_PER_THREAD]) {
    const int w0 = (tile % tiles_w) * C1_TW, h0 = (tile / tiles_w) * C1_TH;
#pragma unroll
    for (int q = 0; q < C1_IN_PER_THREAD; ++q) {
      const int i = threadIdx.x + q * 256;
      v[q] = 0.f;
      if (i < C1_IN) {
        const int c = i / ((C1_TH + 2) * (C1_TW + 2));
        const int rem = i - c * ((C1_TH + 2) * (C1_TW + 2));
        const int sy = rem / (C1_TW + 2), sx = rem - sy * (C1_TW + 2);
        const int hh = h0 + sy - 1, ww = w0 + sx - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) v[q] = (__ldg(x + c * HW + static_cast<size_t>(hh) * W + ww) - pc.mean[c]) / pc.stdv[c];
      }
    }
  };
  auto stash = [&](int buf, const float (&v)[C1_IN_PER_THREAD]) {
#pragma unroll
    for (int q = 0; q < C1_IN_PER_THREAD; ++q) {
      const int i = threadIdx.x + q * 256;
      if (i < C1_IN) (&in[buf][0][0][0])[i] = v[q];
    }
  };
  (void)inv_std;
  const int j = threadIdx.x & 7;    // channel group: channels 8j .. 8j+7
  const int px = threadIdx.x >> 3;  // pixel column inside the tile
  const float4 b0 = *reinterpret_cast<const float4*>(bias + 8 * j);
  const float4 b1 = *reinterpret_cast<const float4*>(bias + 8 * j + 4);
  float nxt[C1_IN_PER_THREAD];
  int buf = 0;
  if (static_cast<int>(blockIdx.x) < tiles) {
    fetch(blockIdx.x, nxt);
    stash(0, nxt);
  }
  __syncthreads();
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, buf ^= 1) {
    const bool more = tile + static_cast<int>(gridDim.x) < tiles;
    if (more) fetch(tile + gridDim.x, nxt);  // in flight while this tile is computed
    const int w0 = (tile % tiles_w) * C1_TW, h0 = (tile / tiles_w) * C1_TH;
    // packed fp32 (fma.rn.f32x2, sm_100): two channels per instruction, bit-identical to two fmaf
    float2 acc[C1_TH][4];
#pragma unroll
    for (int i = 0; i < C1_TH; ++i) {
      acc[i][0] = make_float2(b0.x, b0.y);
      acc[i][1] = make_float2(b0.z, b0.w);
      acc[i][2] = make_float2(b1.x, b1.y);
      acc[i][3] = make_float2(b1.z, b1.w);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        float col[C1_TH + 2];
#pragma unroll
        for (int r = 0; r < C1_TH + 2; ++r) col[r] = in[buf][c][r][px + s];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float4 wa = *reinterpret_cast<const float4*>(ws[c * 9 + r * 3 + s][0][j]);
          const float4 wb = *reinterpret_cast<const float4*>(ws[c * 9 + r * 3 + s][1][j]);
          const float2 q0 = make_float2(wa.x, wa.y), q1 = make_float2(wa.z, wa.w);
          const float2 q2 = make_float2(wb.x, wb.y), q3 = make_float2(wb.z, wb.w);
#pragma unroll
          for (int i = 0; i < C1_TH; ++i) {
            const float2 v = make_float2(col[i + r], col[i + r]);
            acc[i][0] = __ffma2_rn(v, q0, acc[i][0]);
            acc[i][1] = __ffma2_rn(v, q1, acc[i][1]);
            acc[i][2] = __ffma2_rn(v, q2, acc[i][2]);
            acc[i][3] = __ffma2_rn(v, q3, acc[i][3]);
          }
        }
      }
    }
    const int w = w0 + px;
    if (w < W) {
#pragma unroll
      for (int i = 0; i < C1_TH; ++i) {
        const int h = h0 + i;
        if (h >= H) break;
        const size_t o = (static_cast<size_t>(h) * W + w) * 64 + 8 * j;
        uint4 u;
        if (out_tap != nullptr) {
          u.x = pack_half2(acc[i][0].x, acc[i][0].y);
          u.y = pack_half2(acc[i][1].x, acc[i][1].y);
          u.z = pack_half2(acc[i][2].x, acc[i][2].y);
          u.w = pack_half2(acc[i][3].x, acc[i][3].y);
          *reinterpret_cast<uint4*>(out_tap + o) = u;
        }
        if (out_act != nullptr) {
          u.x = pack_half2(fmaxf(acc[i][0].x, 0.f), fmaxf(acc[i][0].y, 0.f));
          u.y = pack_half2(fmaxf(acc[i][1].x, 0.f), fmaxf(acc[i][1].y, 0.f));
          u.z = pack_half2(fmaxf(acc[i][2].x, 0.f), fmaxf(acc[i][2].y, 0.f));
          u.w = pack_half2(fmaxf(acc[i][3].x, 0.f), fmaxf(acc[i][3].y, 0.f));
          *reinterpret_cast<uint4*>(out_act + o) = u;
        }
      }
    }
    if (more) stash(buf ^ 1, nxt);  // the other buffer: nobody reads it during this iteration
    __syncthreads();
  }
}

cudaError_t launch_conv1_fwd(const float* x, const float* w, const float* b, __half* out_tap, __half* out_act, int H,
                             int W, PixelConsts pc, cudaStream_t s) {
  const int blocks = ((W + C1_TW - 1) / C1_TW) * ((H + C1_TH - 1) / C1_TH);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int grid = blocks < 2 * sms ? blocks : 2 * sms;
  conv1_fwd_kernel<<<grid, 256, 0, s>>>(x, w, b, out_tap, out_act, H, W, pc);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// conv1_1 data-gradient: 64 -> 3 channels, + d normalize + TV/edge gradient -> flat fp32 gradient
//   gx[c,h,w] = sum_{n,r,s} gy[h+1-r, w+1-s, n] * W[n,c,r,s]
// ------------------------------------------------------------------------------------------------
static constexpr int D1_TW = 32, D1_TH = 4;
static constexpr int D1_PIX_STRIDE = 72;  // bf16 elements per staged pixel: 64 + 8 pad (144 B) -> conflict-free 16 B reads

__global__ void __launch_bounds__(D1_TW* D1_TH) conv1_dgrad_kernel(const __nv_bfloat16* __restrict__ gy,
                                                                   const float* __restrict__ wgt,
                                                                   const float* __restrict__ grad_pix,
                                                                   float* __restrict__ grad, int H, int W,
                                                                   PixelConsts pc) {
  __shared__ __align__(16) __nv_bfloat16 tile[(D1_TH + 2) * (D1_TW + 2) * D1_PIX_STRIDE];
  __shared__ __align__(16) float ws[9][64][4];  // [r*3+s][n][c (3 used)]
  const int tiles_w = (W + D1_TW - 1) / D1_TW;
  const int w0 = (blockIdx.x % tiles_w) * D1_TW, h0 = (blockIdx.x / tiles_w) * D1_TH;
  const size_t HW = static_cast<size_t>(H) * W;
  for (int i = threadIdx.x; i < 9 * 64; i += D1_TW * D1_TH) {
    const int tap = i / 64, n = i - tap * 64;
    const int r = tap / 3, s = tap - 3 * r;
    ws[tap][n][0] = wgt[((n * 3 + 0) * 3 + r) * 3 + s];
    ws[tap][n][1] = wgt[((n * 3 + 1) * 3 + r) * 3 + s];
    ws[tap][n][2] = wgt[((n * 3 + 2) * 3 + r) * 3 + s];
    ws[tap][n][3] = 0.f;
  }
  // stage the (TH+2) x (TW+2) x 64 halo patch, 16 B per thread per step
  for (int i = threadIdx.x; i < (D1_TH + 2) * (D1_TW + 2) * 8; i += D1_TW * D1_TH) {
    const int pix = i >> 3, q = i & 7;
    const int sy = pix / (D1_TW + 2), sx = pix - sy * (D1_TW + 2);
    const int hh = h0 + sy - 1, ww = w0 + sx - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (hh >= 0 && hh < H && ww >= 0 && ww < W)
      v = __ldg(reinterpret_cast<const uint4*>(gy + (static_cast<size_t>(hh) * W + ww) * 64) + q);
    *reinterpret_cast<uint4*>(tile + pix * D1_PIX_STRIDE + q * 8) = v;
  }
  __syncthreads();
  const int tx = threadIdx.x % D1_TW, ty = threadIdx.x / D1_TW;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      // output pixel (ty, tx) reads gy at tile position (ty + 1 - r + 1, tx + 1 - s + 1)
      const __nv_bfloat16* src = tile + ((ty + 2 - r) * (D1_TW + 2) + (tx + 2 - s)) * D1_PIX_STRIDE;
      const float(*wt)[4] = ws[r * 3 + s];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + q * 8);
        const __nv_bfloat162* bb = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(bb[e]);
          const float4 wa = *reinterpret_cast<const float4*>(wt[q * 8 + 2 * e]);
          const float4 wb = *reinterpret_cast<const float4*>(wt[q * 8 + 2 * e + 1]);
          a0 = fmaf(f.x, wa.x, a0);
          a1 = fmaf(f.x, wa.y, a1);
          a2 = fmaf(f.x, wa.z, a2);
          a0 = fmaf(f.y, wb.x, a0);
          a1 = fmaf(f.y, wb.y, a1);
          a2 = fmaf(f.y, wb.z, a2);
        }
      }
    }
  }
  const int h = h0 + ty, w = w0 + tx;
  if (h >= H || w >= W) return;
  const size_t o = static_cast<size_t>(h) * W + w;
  float g0 = a0 / pc.stdv[0], g1 = a1 / pc.stdv[1], g2 = a2 / pc.stdv[2];
  if (grad_pix != nullptr) {
    g0 += grad_pix[o];
    g1 += grad_pix[HW + o];
    g2 += grad_pix[2 * HW + o];
  }
  grad[o] = g0;
  grad[HW + o] = g1;
  grad[2 * HW + o] = g2;
}

cudaError_t launch_conv1_dgrad(const __nv_bfloat16* gy, const float* w, const float* grad_pix, float* grad, int H,
                               int W, PixelConsts pc, cudaStream_t s) {
  const int blocks = ((W + D1_TW - 1) / D1_TW) * ((H + D1_TH - 1) / D1_TH);
  conv1_dgrad_kernel<<<blocks, D1_TW * D1_TH, 0, s>>>(gy, w, grad_pix, grad, H, W, pc);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// content loss + tap gradient
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CL_THREADS) content_loss_kernel(const __half* __restrict__ y,
                                                                  const float* __restrict__ yc,
                                                                  __nv_bfloat16* __restrict__ addend,
                                                                  double* __restrict__ part, size_t numel,
                                                                  float gcoef, int accumulate, double part_scale) {
  __shared__ double scratch[CL_THREADS / 32];
  float acc = 0.f;
  const size_t nvec = numel / 8;
  for (size_t v = static_cast<size_t>(blockIdx.x) * CL_THREADS + threadIdx.x; v < nvec;
       v += static_cast<size_t>(gridDim.x) * CL_THREADS) {
    const uint4 yy = __ldg(reinterpret_cast<const uint4*>(y) + v);
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(yc) + 2 * v);
    const float4 c1 = __ldg(reinterpret_cast<const float4*>(yc) + 2 * v + 1);
    const __half2* hh = reinterpret_cast<const __half2*>(&yy);
    const float2 f0 = __half22float2(hh[0]), f1 = __half22float2(hh[1]), f2 = __half22float2(hh[2]),
                 f3 = __half22float2(hh[3]);
    const float d[8] = {f0.x - c0.x, f0.y - c0.y, f1.x - c0.z, f1.y - c0.w,
                        f2.x - c1.x, f2.y - c1.y, f3.x - c1.z, f3.y - c1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) acc = fmaf(d[e], d[e], acc);
    if (addend != nullptr) {
      uint4 o;
      float e[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) e[q] = gcoef * d[q];
      if (accumulate) {
        const uint4 old = *(reinterpret_cast<const uint4*>(addend) + v);
        const __nv_bfloat162* ob = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __bfloat1622float2(ob[q]);
          e[2 * q] += f.x;
          e[2 * q + 1] += f.y;
        }
      }
      __nv_bfloat162 b0 = __floats2bfloat162_rn(e[0], e[1]);
      __nv_bfloat162 b1 = __floats2bfloat162_rn(e[2], e[3]);
      __nv_bfloat162 b2 = __floats2bfloat162_rn(e[4], e[5]);
      __nv_bfloat162 b3 = __floats2bfloat162_rn(e[6], e[7]);
      o.x = *reinterpret_cast<uint32_t*>(&b0);
      o.y = *reinterpret_cast<uint32_t*>(&b1);
      o.z = *reinterpret_cast<uint32_t*>(&b2);
      o.w = *reinterpret_cast<uint32_t*>(&b3);
      *(reinterpret_cast<uint4*>(addend) + v) = o;
    }
  }
  const double tot = block_sum(static_cast<double>(acc), scratch);
  // every layer's partial sums carry that layer's own 1 / numel: the assembled value is the mean over the content layers
  // of the per-layer MSEs (style_transfer_losses.py:53-65) also when the layers differ in size
  if (threadIdx.x == 0) part[blockIdx.x] = tot * part_scale;
}

cudaError_t launch_content_loss(const __half* y, const float* yc, __nv_bfloat16* addend, double* part, size_t numel,
                                float gcoef, int accumulate, cudaStream_t s) {
  if (numel % 8 != 0) return cudaErrorInvalidValue;  // C is a multiple of 64
  content_loss_kernel<<<content_blocks(numel), CL_THREADS, 0, s>>>(y, yc, addend, part, numel, gcoef, accumulate,
                                                                   1.0 / static_cast<double>(numel));
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// loss assembly: one block, fixed summation order
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_assemble_kernel(LossAssembleArgs a) {
  __shared__ double scratch[8];
  double tv = 0.0, ed = 0.0, ct = 0.0;
  for (int i = threadIdx.x; i < a.n_tv; i += 256) tv += a.tv_part[i];
  for (int i = threadIdx.x; i < a.n_edge; i += 256) ed += a.edge_part[i];
  for (int i = threadIdx.x; i < a.n_content; i += 256) ct += a.content_part[i];
  tv = block_sum(tv, scratch);
  ed = block_sum(ed, scratch);
  ct = block_sum(ct, scratch);
  // per-layer Gram MSE (style_transfer_losses.py:138-144): the Gram kernels leave per-block sums of (G - T)^2
  __shared__ float s_style[5];
  for (int l = 0; l < a.num_style; ++l) {
    double sq = 0.0;
    for (int i = threadIdx.x; i < a.style_fin_n[l]; i += 256) sq += static_cast<double>(a.style_fin[l][i]);
    sq = block_sum(sq, scratch);
    if (threadIdx.x == 0) s_style[l] = static_cast<float>(sq * static_cast<double>(a.style_inv_cc[l]));
  }
  if (threadIdx.x != 0) return;
  // run_style_transfer.py:115-139: each term is formed in fp32, weighted, then summed s + c + tv + e
  float style = 0.f;
  for (int l = 0; l < a.num_style; ++l) {
    style += s_style[l];
    a.out[5 + l] = s_style[l];
  }
  if (a.num_style > 0) style /= static_cast<float>(a.num_style);
  const float content = static_cast<float>(ct * a.content_norm);
  const float tvl = static_cast<float>(tv * a.tv_norm);
  const float edl = static_cast<float>(0.5 * ed * a.edge_norm);
  const float s_loss = a.w_style > 0.f ? a.w_style * style : 0.f;
  const float c_loss = a.w_content > 0.f ? a.w_content * content : 0.f;
  const float t_loss = a.w_tv > 0.f ? a.w_tv * tvl : 0.f;
  const float e_loss = a.w_edge > 0.f ? a.w_edge * edl : 0.f;
  float total = 0.f;
  total = total + s_loss;
  total = total + c_loss;
  total = total + t_loss;
  total = total + e_loss;
  a.out[0] = total;
  a.out[1] = c_loss;
  a.out[2] = s_loss;
  a.out[3] = t_loss;
  a.out[4] = e_loss;
  a.out[10] = content;
  a.out[11] = style;
  a.out[12] = tvl;
  a.out[13] = edl;
  if (a.counter != nullptr && (a.stop_flag == nullptr || *a.stop_flag == 0)) {
    const int k = *a.counter;
    if (a.trace != nullptr && k < a.trace_cap) {
      float* t = a.trace + static_cast<size_t>(k) * 5;
      t[0] = total;
      t[1] = c_loss;
      t[2] = s_loss;
      t[3] = t_loss;
      t[4] = e_loss;
    }
    *a.counter = k + 1;
  }
}

cudaError_t launch_loss_assemble(const LossAssembleArgs& a, cudaStream_t s) {
  loss_assemble_kernel<<<1, 256, 0, s>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// layout / precision conversion (setup and tests; not on the per-step path)
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 32 x 32 (pixel x channel) transposes through shared memory so both sides stay coalesced
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, size_t P, int C) {
  __shared__ float t[32][33];
  const size_t p0 = static_cast<size_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const size_t p = p0 + i;
    const int c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (p < P && c < C) ? to_f<T>(in[p * C + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const size_t p = p0 + threadIdx.x;
    if (p < P && c < C) out[static_cast<size_t>(c) * P + p] = t[threadIdx.x][i];
  }
}
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, size_t P, int C, int Cp) {
  __shared__ float t[32][33];
  const size_t p0 = static_cast<size_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const size_t p = p0 + threadIdx.x;
    t[i][threadIdx.x] = (p < P && c < C) ? in[static_cast<size_t>(c) * P + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const size_t p = p0 + i;
    const int c = c0 + threadIdx.x;
    if (p < P && c < C) out[p * Cp + c] = from_f<T>(t[threadIdx.x][i]);
  }
}

template <typename T>
static cudaError_t launch_to_nchw(const T* in, float* out, int H, int W, int C, cudaStream_t s) {
  const size_t P = static_cast<size_t>(H) * W;
  dim3 grid(static_cast<unsigned>((P + 31) / 32), (C + 31) / 32), block(32, 8);
  nhwc_to_nchw_kernel<T><<<grid, block, 0, s>>>(in, out, P, C);
  return cudaGetLastError();
}
template <typename T>
static cudaError_t launch_to_nhwc(const float* in, T* out, int H, int W, int C, cudaStream_t s, int Cp = 0) {
  const size_t P = static_cast<size_t>(H) * W;
  dim3 grid(static_cast<unsigned>((P + 31) / 32), (C + 31) / 32), block(32, 8);
  nchw_to_nhwc_kernel<T><<<grid, block, 0, s>>>(in, out, P, C, Cp > 0 ? Cp : C);
  return cudaGetLastError();
}
cudaError_t launch_nhwc_half_to_nchw_float(const __half* in, float* out, int H, int W, int C, cudaStream_t s) {
  return launch_to_nchw<__half>(in, out, H, W, C, s);
}
cudaError_t launch_nchw_float_to_nhwc_half(const float* in, __half* out, int H, int W, int C, cudaStream_t s) {
  return launch_to_nhwc<__half>(in, out, H, W, C, s);
}
cudaError_t launch_nchw_float_to_nhwc_half_padded(const float* in, __half* out, int H, int W, int C, int Cp,
                                                  cudaStream_t s) {
  return launch_to_nhwc<__half>(in, out, H, W, C, s, Cp);
}
cudaError_t launch_nhwc_bf16_to_nchw_float(const __nv_bfloat16* in, float* out, int H, int W, int C, cudaStream_t s) {
  return launch_to_nchw<__nv_bfloat16>(in, out, H, W, C, s);
}
cudaError_t launch_nchw_float_to_nhwc_bf16(const float* in, __nv_bfloat16* out, int H, int W, int C, cudaStream_t s) {
  return launch_to_nhwc<__nv_bfloat16>(in, out, H, W, C, s);
}

// torch [Cout,Cin,3,3] -> forward operand [9][Cout][Cin] fp16, data-gradient operand [9][Cin][Cout] bf16 (taps flipped)
__global__ void pack_weights_kernel(const float* __restrict__ w, __half* __restrict__ wf, __nv_bfloat16* __restrict__ wb,
                                    int Cout, int Cin) {
  const size_t total = static_cast<size_t>(Cout) * Cin * 9;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int tap = static_cast<int>(i % 9);
    const size_t nk = i / 9;
    const int k = static_cast<int>(nk % Cin);
    const int n = static_cast<int>(nk / Cin);
    const float v = w[i];
    if (wf != nullptr) wf[(static_cast<size_t>(tap) * Cout + n) * Cin + k] = __float2half_rn(v);
    if (wb != nullptr) wb[(static_cast<size_t>(8 - tap) * Cin + k) * Cout + n] = __float2bfloat16_rn(v);
  }
}
// Forward weights: fp16 with ERROR-FEEDBACK rounding along the reduction.  Round-to-nearest makes the rounding error of every
// weight independent, and the Gram error of every layer was dominated by exactly that (measured by rounding source on the
// CPU: weights 3.2e-4 .. 6.0e-4 relative Frobenius error at conv2_1 .. conv5_1, activations 0.2e-4 .. 1.5e-4): the inputs
// of a convolution are post-ReLU activations - positive mean, smooth across the nine taps of an input channel - so the
// output error sum_k dw_k a_k is mostly (local mean of a) x (sum of dw over neighbouring k).  Carrying each weight's rounding
// error into the next weight of the same filter (torch order: input channel major, then the 3 x 3 taps) keeps every such
// partial sum of dw within half an ulp: 1.1e-4 .. 2.6e-4 at the same layers.  Each stored weight is still one of the two fp16
// neighbours of the fp32 weight; nothing changes at run time.  One thread per output channel, once per network.
__global__ void pack_weights_fwd_feedback_kernel(const float* __restrict__ w, __half* __restrict__ wf, int Cout, int Cin) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Cout) return;
  const float* src = w + static_cast<size_t>(n) * Cin * 9;
  float carry = 0.f;
  for (int k = 0; k < Cin; ++k) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float v = src[k * 9 + tap] + carry;
      const __half h = __float2half_rn(v);
      carry = v - __half2float(h);
      wf[(static_cast<size_t>(tap) * Cout + n) * Cin + k] = h;
    }
  }
}

cudaError_t launch_pack_weights(const float* w, __half* w_fwd, __nv_bfloat16* w_bwd, int Cout, int Cin, cudaStream_t s) {
  pack_weights_kernel<<<592, 256, 0, s>>>(w, nullptr, w_bwd, Cout, Cin);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || w_fwd == nullptr) return e;
  pack_weights_fwd_feedback_kernel<<<(Cout + 63) / 64, 64, 0, s>>>(w, w_fwd, Cout, Cin);
  return cudaGetLastError();
}

__global__ void pack_weights_conv1_bwd_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wb) {
  // wb[tap'][c][n] = w[n][c][2-r'][2-s'] for c < 3, zero for the 13 padding channels
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 9 * 16 * 64; i += gridDim.x * blockDim.x) {
    const int n = i % 64, c = (i / 64) % 16, tap = i / (64 * 16);
    float v = 0.f;
    if (c < 3) v = w[(n * 3 + c) * 9 + (8 - tap)];
    wb[i] = __float2bfloat16_rn(v);
  }
}
cudaError_t launch_pack_weights_conv1_bwd(const float* w, __nv_bfloat16* w_bwd, cudaStream_t s) {
  pack_weights_conv1_bwd_kernel<<<36, 256, 0, s>>>(w, w_bwd);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// content target and channel attention
// ------------------------------------------------------------------------------------------------
__global__ void make_content_target_kernel(const __half* __restrict__ y, const float* __restrict__ gate,
                                           float* __restrict__ yc, size_t numel, int C) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float g = gate != nullptr ? gate[i % C] : 1.f;
    yc[i] = __half2float(y[i]) * g;
  }
}
cudaError_t launch_make_content_target(const __half* y, const float* gate, float* yc, size_t pixels, int C,
                                       cudaStream_t s) {
  make_content_target_kernel<<<592, 256, 0, s>>>(y, gate, yc, pixels * C, C);
  return cudaGetLastError();
}

// column means of a [pixels, C] fp16 matrix: one block per 32 channels, fixed reduction order
__global__ void __launch_bounds__(256) channel_mean_kernel(const __half* __restrict__ y, float* __restrict__ pooled,
                                                           size_t pixels, int C) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < C)
    for (size_t p = threadIdx.y; p < pixels; p += 8) acc += __half2float(y[p * C + c]);
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    pooled[c] = t / static_cast<float>(pixels);
  }
}
// gate = sigmoid(relu(W2 relu(W1 pooled)))      ChannelAttention.py:30-37
__global__ void __launch_bounds__(512) channel_gate_kernel(const float* __restrict__ pooled,
                                                           const float* __restrict__ w1, const float* __restrict__ w2,
                                                           float* __restrict__ gate, int C, int Cr) {
  extern __shared__ float sm[];  // pooled[C] | hidden[Cr]
  float* sp = sm;
  float* sh = sm + C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) sp[i] = pooled[i];
  __syncthreads();
  for (int o = threadIdx.x; o < Cr; o += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < C; ++k) acc = fmaf(w1[static_cast<size_t>(o) * C + k], sp[k], acc);
    sh[o] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < C; o += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < Cr; ++k) acc = fmaf(w2[static_cast<size_t>(o) * Cr + k], sh[k], acc);
    acc = fmaxf(acc, 0.f);
    gate[o] = 1.f / (1.f + expf(-acc));
  }
}
cudaError_t launch_channel_gate(const __half* y, const float* w1, const float* w2, float* pooled, float* gate,
                                size_t pixels, int C, int Cr, cudaStream_t s) {
  channel_mean_kernel<<<(C + 31) / 32, dim3(32, 8), 0, s>>>(y, pooled, pixels, C);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  channel_gate_kernel<<<1, 512, (C + Cr) * sizeof(float), s>>>(pooled, w1, w2, gate, C, Cr);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// StyleMixer: out = (1 - wgt) * resize(a) + wgt * resize(b), bilinear, align_corners=True
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src(int o, int in, int out, int& i0, int& i1, float& l1) {
  // at::native area_pixel_compute_scale(align_corners=true) = (in - 1) / (out - 1), source = scale * o
  const float scale = out > 1 ? static_cast<float>(in - 1) / static_cast<float>(out - 1) : 0.f;
  const float src = scale * static_cast<float>(o);
  i0 = static_cast<int>(src);
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = src - static_cast<float>(i0);
}
__device__ __forceinline__ float bilerp_at(const __half* f, int Wf, int C, int c, int h0, int h1, float lh, int w0,
                                           int w1, float lw) {
  const float v00 = __half2float(f[(static_cast<size_t>(h0) * Wf + w0) * C + c]);
  const float v01 = __half2float(f[(static_cast<size_t>(h0) * Wf + w1) * C + c]);
  const float v10 = __half2float(f[(static_cast<size_t>(h1) * Wf + w0) * C + c]);
  const float v11 = __half2float(f[(static_cast<size_t>(h1) * Wf + w1) * C + c]);
  const float l0h = 1.f - lh, l0w = 1.f - lw;
  return l0h * (l0w * v00 + lw * v01) + lh * (l0w * v10 + lw * v11);
}
__global__ void style_mix_kernel(const __half* __restrict__ a, int Ha, int Wa, const __half* __restrict__ b, int Hb,
                                 int Wb, __half* __restrict__ out, int Ho, int Wo, int C, float wgt) {
  const size_t total = static_cast<size_t>(Ho) * Wo * C;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t pix = i / C;
    const int wo = static_cast<int>(pix % Wo), ho = static_cast<int>(pix / Wo);
    int h0, h1, w0, w1;
    float lh, lw;
    bilinear_src(ho, Ha, Ho, h0, h1, lh);
    bilinear_src(wo, Wa, Wo, w0, w1, lw);
    const float va = bilerp_at(a, Wa, C, c, h0, h1, lh, w0, w1, lw);
    bilinear_src(ho, Hb, Ho, h0, h1, lh);
    bilinear_src(wo, Wb, Wo, w0, w1, lw);
    const float vb = bilerp_at(b, Wb, C, c, h0, h1, lh, w0, w1, lw);
    out[i] = __float2half_rn((1.f - wgt) * va + wgt * vb);  // StyleMixer.py:23,37
  }
}
cudaError_t launch_style_mix(const __half* a, int Ha, int Wa, const __half* b, int Hb, int Wb, __half* out, int Ho,
                             int Wo, int C, float wgt, cudaStream_t s) {
  style_mix_kernel<<<1184, 256, 0, s>>>(a, Ha, Wa, b, Hb, Wb, out, Ho, Wo, C, wgt);
  return cudaGetLastError();
}

}  // namespace nst
