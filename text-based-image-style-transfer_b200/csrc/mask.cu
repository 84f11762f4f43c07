// Mask compositing: blends the stylised image into the original through a segmentation mask with Gaussian-softened edges.
//
// Replaces  text/segmentation_style_transfer.py:5-94  (segmentation_style_transfer, _edge_smoothing) - the step right after
// the style-transfer loop in six of the reference's eleven call sites (app.py:203,318,407,512,...) - i.e.
//   cv2.GaussianBlur(mask * 255, (k, k), 0)  ->  / 255.0  ->  content * (1 - m) + style * m  ->  astype(uint8)
// in ONE kernel, bit for bit: the blur in OpenCV's 8-bit fixed-point arithmetic (kernel with error diffusion, reflect-101
// border, 8.8 horizontal and 16.16 vertical pass, round half up), the blend in fp64 without fused multiply-adds (numpy does
// a double multiply, a double multiply and a double add), truncation to uint8.  Bandwidth bound: 7 bytes read and 3 written
// per pixel; the mask tile with its halo is staged in shared memory once.
#include "mask.cuh"

#include <math.h>
#include <stdint.h>

namespace nst {

static constexpr int MK_T = 32;  // output tile edge

struct MaskWeights {
  int k;
  int w[MASK_MAX_K];
};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * n - 2;
  i %= period;
  if (i < 0) i += period;
  return i >= n ? period - i : i;
}

__global__ void __launch_bounds__(256) mask_composite_kernel(const uint8_t* __restrict__ content, const uint8_t* __restrict__ style,
                                                             const uint8_t* __restrict__ mask, uint8_t* __restrict__ out, int H, int W,
                                                             int C, const __grid_constant__ MaskWeights mw) {
  extern __shared__ __align__(16) uint8_t mk_smem[];
  const int k = mw.k, r = k >> 1;
  const int E = MK_T + 2 * r;                       // staged tile edge
  uint8_t* sm = mk_smem;                            // [E][E] mask values 0 / 255
  int* sh = reinterpret_cast<int*>(mk_smem + ((E * E + 15) & ~15));  // [E][MK_T] horizontal pass, 8.8 fixed point
  // per-block tables: byte -> double, and the two blend factors of every possible blurred mask value, computed with exactly
  // the reference's operations (blurred / 255.0, 1 - that): the per-pixel work is then two multiplies and an add in fp64
  __shared__ double s_nb[256], s_na[256];
  {
    const double nb = __ddiv_rn(static_cast<double>(threadIdx.x), 255.0);          // :86
    s_nb[threadIdx.x] = nb;
    s_na[threadIdx.x] = __dsub_rn(1.0, nb);
  }
  const int x0 = blockIdx.x * MK_T, y0 = blockIdx.y * MK_T;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads: no integer division in the loops below
  for (int yy = ty; yy < E; yy += 8) {
    int gy = y0 + yy - r;
    if (static_cast<unsigned>(gy) >= static_cast<unsigned>(H)) gy = reflect101(gy, H);
    const uint8_t* mrow = mask + static_cast<size_t>(gy) * W;
    for (int xx = tx; xx < E; xx += 32) {
      int gx = x0 + xx - r;
      if (static_cast<unsigned>(gx) >= static_cast<unsigned>(W)) gx = reflect101(gx, W);
      sm[yy * E + xx] = mrow[gx] ? 255 : 0;                                        // :80
    }
  }
  __syncthreads();
  for (int yy = ty; yy < E; yy += 8) {
    int acc = 0;
    for (int j = 0; j < k; ++j) acc += mw.w[j] * sm[yy * E + tx + j];
    sh[yy * MK_T + tx] = acc;
  }
  __syncthreads();
  const int gx = x0 + tx;
  for (int yy = ty; yy < MK_T; yy += 8) {
    const int gy = y0 + yy;
    if (gy >= H || gx >= W) continue;
    int acc = 0;
    for (int j = 0; j < k; ++j) acc += mw.w[j] * sh[(yy + j) * MK_T + tx];
    const int blurred = (acc + 32768) >> 16;                                       // :83
    const double nb = s_nb[blurred], na = s_na[blurred];
    const size_t o = (static_cast<size_t>(gy) * W + gx) * C;
    for (int c = 0; c < C; ++c) {
      const double v = __dadd_rn(__dmul_rn(static_cast<double>(content[o + c]), na), __dmul_rn(static_cast<double>(style[o + c]), nb));
      out[o + c] = static_cast<uint8_t>(static_cast<int>(v));                     // :91 astype(np.uint8) truncates
    }
  }
}

__global__ void mask_select_kernel(const uint8_t* __restrict__ content, const uint8_t* __restrict__ style,
                                   const uint8_t* __restrict__ mask, uint8_t* __restrict__ out, size_t npix, int C) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const uint8_t* src = mask[i] ? style : content;                                  // :52 np.where(mask > 0, style, content)
  for (int c = 0; c < C; ++c) out[i * C + c] = src[i * C + c];
}

// the 8-bit fixed-point kernel of cv2.GaussianBlur(uint8, (k, k), 0): OpenCV smooth.dispatch.cpp getGaussianKernelBitExact +
// getGaussianKernelFixedPoint_ED (error diffusion on the first half, centre = 256 - 2 * sum)
int mask_gaussian_weights(int k, int* w) {
  if (k < 1 || k > MASK_MAX_K || (k & 1) == 0) return -1;
  static const double small_tab[4][7] = {{1.0}, {0.25, 0.5, 0.25}, {0.0625, 0.25, 0.375, 0.25, 0.0625},
                                         {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125}};
  double g[MASK_MAX_K];
  if (k <= 7) {
    for (int i = 0; i < k; ++i) g[i] = small_tab[k >> 1][i];
  } else {
    const double sigma = 0.3 * ((k - 1) * 0.5 - 1) + 0.8;
    const double scale2x = -0.5 / (sigma * sigma);
    double s = 0.0;
    for (int i = 0; i < k; ++i) {
      const double x = i - (k - 1) * 0.5;
      g[i] = exp(scale2x * x * x);
      s += g[i];
    }
    for (int i = 0; i < k; ++i) g[i] = g[i] / s;
  }
  const int n2 = k / 2;
  double err = 0.0;
  long long tot = 0;
  for (int i = 0; i < n2; ++i) {
    const double adj = g[i] * 256.0 + err;
    const long long v0 = static_cast<long long>(nearbyint(adj));
    err = adj - static_cast<double>(v0);
    w[i] = w[k - 1 - i] = static_cast<int>(v0);
    tot += v0;
  }
  w[n2] = static_cast<int>(256 - 2 * tot);
  return 0;
}

cudaError_t launch_mask_composite(const uint8_t* content, const uint8_t* style, const uint8_t* mask, uint8_t* out, int H, int W, int C,
                                  int k, cudaStream_t s) {
  if (k == 0) {
    const size_t npix = static_cast<size_t>(H) * W;
    mask_select_kernel<<<static_cast<unsigned>((npix + 255) / 256), 256, 0, s>>>(content, style, mask, out, npix, C);
    return cudaGetLastError();
  }
  MaskWeights mw;
  mw.k = k;
  if (mask_gaussian_weights(k, mw.w) != 0) return cudaErrorInvalidValue;
  const int E = MK_T + 2 * (k >> 1);
  const size_t smem = ((static_cast<size_t>(E) * E + 15) & ~static_cast<size_t>(15)) + static_cast<size_t>(E) * MK_T * sizeof(int);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mask_composite_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  dim3 grid((W + MK_T - 1) / MK_T, (H + MK_T - 1) / MK_T);
  mask_composite_kernel<<<grid, 256, smem, s>>>(content, style, mask, out, H, W, C, mw);
  return cudaGetLastError();
}

}  // namespace nst
