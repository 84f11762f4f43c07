// Function-level mirrors of multi_style_transfer/style_transfer_losses.py on plain [b,c,h,w] fp32 device
// tensors (torch layout).  These back the Python functions of the same names; the optimisation loop itself
// uses the fused kernels in pixel.cu / gram.cu instead.  All reductions are two-stage with a fixed order.
#include "../../include/nst_b200.h"

#include <cuda_runtime.h>
#include <stdio.h>

#include "common.cuh"

using namespace nst;

namespace {

constexpr int RT = 256;
constexpr int RB = 592;  // 4 blocks per SM

__global__ void normalize_kernel(const float* __restrict__ x, float* __restrict__ y, size_t plane, int C, size_t total,
                                 float m0, float m1, float m2, float s0, float s1, float s2) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>((i / plane) % C);
    const float m = c == 0 ? m0 : (c == 1 ? m1 : m2);
    const float s = c == 0 ? s0 : (c == 1 ? s1 : s2);
    y[i] = (x[i] - m) / s;  // style_transfer_losses.py:26
  }
}

__global__ void gray_kernel(const float* __restrict__ x, float* __restrict__ y, size_t plane, int C) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < plane;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc += x[c * plane + i];
    y[i] = acc / static_cast<float>(C);  // helper_functions.py:113
  }
}

// stage 1 of sum((a-b)^2) or, with b == nullptr and tv != 0, of the total variation of [planes,H,W]
__global__ void __launch_bounds__(RT) sqdiff_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            size_t n, double* __restrict__ part) {
  __shared__ double scratch[RT / 32];
  float acc = 0.f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * RT + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * RT) {
    const float d = a[i] - b[i];
    acc = fmaf(d, d, acc);
  }
  const double t = block_sum(static_cast<double>(acc), scratch);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}
__global__ void __launch_bounds__(RT) tv_partial_kernel(const float* __restrict__ y, int planes, int H, int W,
                                                        double* __restrict__ part) {
  __shared__ double scratch[RT / 32];
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t n = plane * planes;
  float acc = 0.f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * RT + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * RT) {
    const size_t r = i % plane;
    const int h = static_cast<int>(r / W), w = static_cast<int>(r % W);
    const float v = y[i];
    if (h + 1 < H) acc += fabsf(y[i + W] - v);  // style_transfer_losses.py:164
    if (w + 1 < W) acc += fabsf(y[i + 1] - v);  // :167
  }
  const double t = block_sum(static_cast<double>(acc), scratch);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}
__global__ void finish_kernel(const double* __restrict__ part, int n, double scale, float* __restrict__ out) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) acc += part[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) *out = static_cast<float>(acc * scale);
}

// x * sigmoid(relu(W2 relu(W1 mean_hw(x))))  on [C,H,W] fp32      ChannelAttention.py:23-40
__global__ void __launch_bounds__(RT) plane_mean_kernel(const float* __restrict__ x, float* __restrict__ pooled,
                                                        size_t plane) {
  __shared__ double scratch[RT / 32];
  const float* p = x + static_cast<size_t>(blockIdx.x) * plane;
  float acc = 0.f;
  for (size_t i = threadIdx.x; i < plane; i += RT) acc += p[i];
  const double t = block_sum(static_cast<double>(acc), scratch);
  if (threadIdx.x == 0) pooled[blockIdx.x] = static_cast<float>(t / static_cast<double>(plane));
}
__global__ void __launch_bounds__(512) gate_kernel(const float* __restrict__ pooled, const float* __restrict__ w1,
                                                   const float* __restrict__ w2, float* __restrict__ gate, int C,
                                                   int Cr) {
  extern __shared__ float sm[];
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
    gate[o] = 1.f / (1.f + expf(-fmaxf(acc, 0.f)));
  }
}
__global__ void scale_planes_kernel(const float* __restrict__ x, const float* __restrict__ gate, float* __restrict__ y,
                                    size_t plane, size_t total) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    y[i] = x[i] * gate[i / plane];
}

__device__ __forceinline__ void src_index(int o, int in, int out, int& i0, int& i1, float& l1) {
  const float scale = out > 1 ? static_cast<float>(in - 1) / static_cast<float>(out - 1) : 0.f;
  const float src = scale * static_cast<float>(o);
  i0 = static_cast<int>(src);
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = src - static_cast<float>(i0);
}
__device__ __forceinline__ float bilerp(const float* p, int W, int h0, int h1, float lh, int w0, int w1, float lw) {
  const float v00 = p[static_cast<size_t>(h0) * W + w0], v01 = p[static_cast<size_t>(h0) * W + w1];
  const float v10 = p[static_cast<size_t>(h1) * W + w0], v11 = p[static_cast<size_t>(h1) * W + w1];
  return (1.f - lh) * ((1.f - lw) * v00 + lw * v01) + lh * ((1.f - lw) * v10 + lw * v11);
}
// StyleMixer.mix on [C,H,W] fp32 tensors      StyleMixer.py:25-38
__global__ void mix_chw_kernel(const float* __restrict__ a, int Ha, int Wa, const float* __restrict__ b, int Hb, int Wb,
                               float* __restrict__ out, int Ho, int Wo, int C, float wb) {
  const size_t total = static_cast<size_t>(C) * Ho * Wo;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int wo = static_cast<int>(i % Wo);
    const int ho = static_cast<int>((i / Wo) % Ho);
    const int c = static_cast<int>(i / (static_cast<size_t>(Wo) * Ho));
    int h0, h1, w0, w1;
    float lh, lw;
    src_index(ho, Ha, Ho, h0, h1, lh);
    src_index(wo, Wa, Wo, w0, w1, lw);
    const float va = bilerp(a + static_cast<size_t>(c) * Ha * Wa, Wa, h0, h1, lh, w0, w1, lw);
    src_index(ho, Hb, Ho, h0, h1, lh);
    src_index(wo, Wb, Wo, w0, w1, lw);
    const float vb = bilerp(b + static_cast<size_t>(c) * Hb * Wb, Wb, h0, h1, lh, w0, w1, lw);
    out[i] = (1.f - wb) * va + wb * vb;
  }
}

int reduce_finish(double* part, int nblk, double scale, float* out, cudaStream_t s) {
  finish_kernel<<<1, 32, 0, s>>>(part, nblk, scale, out);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(part);
  return e == cudaSuccess ? NST_OK : NST_ERR_CUDA;
}

}  // namespace

extern "C" int nst_normalize(const float* x, float* y, int planes, int C, int H, int W, const float mean[3],
                             const float std[3], void* stream) {
  if (!x || !y || C != 3 || planes % C != 0) return NST_ERR_ARG;
  const size_t plane = static_cast<size_t>(H) * W;
  normalize_kernel<<<RB, RT, 0, static_cast<cudaStream_t>(stream)>>>(x, y, plane, C, plane * planes, mean[0], mean[1],
                                                                   mean[2], std[0], std[1], std[2]);
  return cudaGetLastError() == cudaSuccess ? NST_OK : NST_ERR_CUDA;
}

extern "C" int nst_grayscale(const float* x, float* y, int C, int H, int W, void* stream) {
  if (!x || !y || C < 1) return NST_ERR_ARG;
  gray_kernel<<<RB, RT, 0, static_cast<cudaStream_t>(stream)>>>(x, y, static_cast<size_t>(H) * W, C);
  return cudaGetLastError() == cudaSuccess ? NST_OK : NST_ERR_CUDA;
}

extern "C" int nst_mse(const float* a, const float* b, size_t n, float* out, void* stream) {
  if (!a || !b || !out || n == 0) return NST_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* part = nullptr;
  if (cudaMalloc(&part, RB * sizeof(double)) != cudaSuccess) return NST_ERR_CUDA;
  sqdiff_partial_kernel<<<RB, RT, 0, s>>>(a, b, n, part);
  return reduce_finish(part, RB, 1.0 / static_cast<double>(n), out, s);
}

extern "C" int nst_total_variation(const float* y, int planes, int H, int W, float* out, void* stream) {
  if (!y || !out || planes < 1 || H < 1 || W < 1) return NST_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* part = nullptr;
  if (cudaMalloc(&part, RB * sizeof(double)) != cudaSuccess) return NST_ERR_CUDA;
  tv_partial_kernel<<<RB, RT, 0, s>>>(y, planes, H, W, part);
  // style_transfer_losses.py:160: normalised by c * h * w (not by the batch size)
  return reduce_finish(part, RB, 1.0 / (static_cast<double>(planes) * H * W), out, s);
}

extern "C" int nst_channel_attention_chw(const float* x, int C, int H, int W, const float* w1, const float* w2,
                                         int reduction, float* y, void* stream) {
  if (!x || !y || !w1 || !w2 || reduction < 1 || C % reduction != 0) return NST_ERR_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* tmp = nullptr;
  if (cudaMalloc(&tmp, static_cast<size_t>(2) * C * sizeof(float)) != cudaSuccess) return NST_ERR_CUDA;
  const size_t plane = static_cast<size_t>(H) * W;
  plane_mean_kernel<<<C, RT, 0, s>>>(x, tmp, plane);
  gate_kernel<<<1, 512, (C + C / reduction) * sizeof(float), s>>>(tmp, w1, w2, tmp + C, C, C / reduction);
  scale_planes_kernel<<<RB, RT, 0, s>>>(x, tmp + C, y, plane, plane * C);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(tmp);
  return e == cudaSuccess ? NST_OK : NST_ERR_CUDA;
}

extern "C" int nst_style_mix_tensors(const float* a, int Ha, int Wa, const float* b, int Hb, int Wb, int C,
                                     float weight_b, float* out, void* stream) {
  if (!a || !b || !out || C < 1) return NST_ERR_ARG;
  const int Ho = Ha + Hb / 2, Wo = Wa + Wb / 2;  // StyleMixer.py:31-32
  mix_chw_kernel<<<RB, RT, 0, static_cast<cudaStream_t>(stream)>>>(a, Ha, Wa, b, Hb, Wb, out, Ho, Wo, C, weight_b);
  return cudaGetLastError() == cudaSuccess ? NST_OK : NST_ERR_CUDA;
}
