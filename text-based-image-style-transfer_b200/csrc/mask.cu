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
  int seen_on = 0, seen_off = 0;
  for (int yy = ty; yy < E; yy += 8) {
    int gy = y0 + yy - r;
    if (static_cast<unsigned>(gy) >= static_cast<unsigned>(H)) gy = reflect101(gy, H);
    const uint8_t* mrow = mask + static_cast<size_t>(gy) * W;
    for (int xx = tx; xx < E; xx += 32) {
      int gx = x0 + xx - r;
      if (static_cast<unsigned>(gx) >= static_cast<unsigned>(W)) gx = reflect101(gx, W);
      const bool on = mrow[gx] != 0;
      seen_on |= on;
      seen_off |= !on;
      sm[yy * E + xx] = on ? 255 : 0;                                              // :80
    }
  }
  // A tile whose whole neighbourhood is inside (or outside) the mask blurs to 255 (0) everywhere: the blend returns the
  // style (content) pixel exactly, so the tile is a plain copy.  Most tiles of a segmentation mask are of this kind.
  const int any_on = __syncthreads_or(seen_on), any_off = __syncthreads_or(seen_off);
  if (!(any_on && any_off)) {
    const uint8_t* src = any_on ? style : content;
    const bool words_ = ((static_cast<size_t>(W) * C) & 3) == 0 && (x0 + MK_T <= W) &&
                        ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out)) & 3) == 0;
    if (words_) {
      uint32_t wv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int gy = y0 + ty + 8 * a;
        wv[a] = (gy < H && tx < MK_T * C / 4) ? __ldg(reinterpret_cast<const uint32_t*>(src + (static_cast<size_t>(gy) * W + x0) * C) + tx) : 0u;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int gy = y0 + ty + 8 * a;
        if (gy < H && tx < MK_T * C / 4) reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(gy) * W + x0) * C)[tx] = wv[a];
      }
      return;
    }
    for (int yy = ty; yy < MK_T; yy += 8) {
      const int gy = y0 + yy;
      if (gy >= H) break;
      const size_t rowo = (static_cast<size_t>(gy) * W + x0) * C;
      if (x0 + tx < W)
        for (int c = 0; c < C; ++c) out[rowo + static_cast<size_t>(tx) * C + c] = src[rowo + static_cast<size_t>(tx) * C + c];
    }
    return;
  }
  for (int yy = ty; yy < E; yy += 8) {
    int acc = 0;
    for (int j = 0; j < k; ++j) acc += mw.w[j] * sm[yy * E + tx + j];
    sh[yy * MK_T + tx] = acc;
  }
  __syncthreads();
  // ---- blend.  Per channel the reference computes trunc(fl(fl(c * na) + fl(s * nb))) in fp64 with nb = fl(m / 255), na = fl(1 - nb).
  // The exact value is (255 c + m (s - c)) / 255; the fp64 result is within 1e-13 of it, so its floor is the integer quotient
  // unless the division is exact (remainder 0) - only then can the rounding of the fp64 sequence decide between q and q - 1,
  // and only then is that sequence executed (m = 0 and m = 255 give c and s exactly).  Rows whose byte length is a multiple
  // of four go through shared memory as 32-bit words (a warp's 32 pixels x 3 channels = 24 words), others byte by byte.
  __shared__ uint32_t s_io[8][3][MK_T];  // per warp: content, style / result words (up to 4 channels x 32 pixels = 32 words)
  const int gx = x0 + tx;
  const bool words = ((static_cast<size_t>(W) * C) & 3) == 0 && (x0 + MK_T <= W) &&
                     ((reinterpret_cast<uintptr_t>(content) | reinterpret_cast<uintptr_t>(style) | reinterpret_cast<uintptr_t>(out)) & 3) == 0;
  const int nwords = MK_T * C / 4;  // 24 for RGB
  for (int yy = ty; yy < MK_T; yy += 8) {
    const int gy = y0 + yy;
    if (gy >= H) break;  // warp-uniform
    const size_t rowo = (static_cast<size_t>(gy) * W + x0) * C;
    const uint8_t* cb;
    const uint8_t* sb;
    uint8_t* ob;
    if (words) {
      if (tx < nwords) {
        s_io[ty][0][tx] = __ldg(reinterpret_cast<const uint32_t*>(content + rowo) + tx);
        s_io[ty][1][tx] = __ldg(reinterpret_cast<const uint32_t*>(style + rowo) + tx);
      }
      __syncwarp();
      cb = reinterpret_cast<const uint8_t*>(s_io[ty][0]) + tx * C;
      sb = reinterpret_cast<const uint8_t*>(s_io[ty][1]) + tx * C;
      ob = reinterpret_cast<uint8_t*>(s_io[ty][2]) + tx * C;
    } else {
      cb = content + rowo + static_cast<size_t>(tx) * C;
      sb = style + rowo + static_cast<size_t>(tx) * C;
      ob = out + rowo + static_cast<size_t>(tx) * C;
    }
    if (gx < W) {
      int acc = 0;
      for (int j = 0; j < k; ++j) acc += mw.w[j] * sh[(yy + j) * MK_T + tx];
      const int m = (acc + 32768) >> 16;                                           // :83 blurred mask value
      for (int c = 0; c < C; ++c) {
        const int cv = cb[c], sv = sb[c];
        int res;
        if (m == 0) {
          res = cv;
        } else if (m == 255) {
          res = sv;
        } else {
          const int num = 255 * cv + m * (sv - cv);  // 0 .. 65025
          const int q = num / 255;
          if (num - 255 * q != 0) {
            res = q;
          } else {
            const double v = __dadd_rn(__dmul_rn(static_cast<double>(cv), s_na[m]), __dmul_rn(static_cast<double>(sv), s_nb[m]));
            res = static_cast<int>(v);                                             // :91 astype(np.uint8) truncates
          }
        }
        ob[c] = static_cast<uint8_t>(res);
      }
    }
    if (words) {
      __syncwarp();
      if (tx < nwords) reinterpret_cast<uint32_t*>(out + rowo)[tx] = s_io[ty][2][tx];
      __syncwarp();
    }
  }
}

// ---- the common case as its own kernel: RGB, k <= 9, row length a multiple of 64 pixels.
// 64 x 64 tiles; every global access is a 16-byte (image) or 4-byte (mask) word and all of a phase's loads are issued before
// the first is used - a block keeps ~12 KB in flight, which is what it takes to approach the HBM rate with byte images (the
// general kernel above moves one 4-byte word per thread per step and is latency bound at ~1.3 TB/s).
static constexpr int MF_T = 64;              // tile edge
static constexpr int MF_R = 4;               // largest radius (k <= 9)
static constexpr int MF_EW = MF_T + 2 * MF_R;  // staged mask row: 72 bytes = 18 aligned words starting at x0 - 4
static constexpr int MF_ROWB = MF_T * 3;     // 192 bytes of RGB per tile row = 12 uint4
static constexpr int MF_SMEM = MF_EW * MF_EW + MF_EW * MF_T * 2 + 2 * MF_T * MF_ROWB;  // mask + horizontal pass + two image tiles

__global__ void __launch_bounds__(256) mask_composite_fast_kernel(const uint8_t* __restrict__ content, const uint8_t* __restrict__ style,
                                                                  const uint8_t* __restrict__ mask, uint8_t* __restrict__ out, int H, int W,
                                                                  const __grid_constant__ MaskWeights mw) {
  extern __shared__ __align__(16) uint8_t mk_smem[];
  const int k = mw.k, r = k >> 1;
  const int E = MF_T + 2 * r;  // staged rows
  uint8_t* sm = mk_smem;                                                        // [E][72] mask 0 / 255
  uint16_t* sh = reinterpret_cast<uint16_t*>(mk_smem + MF_EW * MF_EW);          // [E][64] horizontal pass, 8.8 fixed point (<= 65280)
  uint8_t* sc = mk_smem + MF_EW * MF_EW + MF_EW * MF_T * 2;                     // [64][192] content tile, then the result
  uint8_t* ss = sc + MF_T * MF_ROWB;                                            // [64][192] style tile
  const int x0 = blockIdx.x * MF_T, y0 = blockIdx.y * MF_T;
  const int tid = threadIdx.x;
  const size_t rowb = static_cast<size_t>(W) * 3;

  // ---- mask rows y0 - r .. y0 + 63 + r, columns x0 - 4 .. x0 + 67, as 18 words per row
  constexpr int MW_PER_ROW = MF_EW / 4, M_STEPS = (MF_EW * MW_PER_ROW + 255) / 256;  // 18, 6
  uint32_t mv[M_STEPS];
#pragma unroll
  for (int i = 0; i < M_STEPS; ++i) {
    const int idx = tid + 256 * i;
    const int yy = idx / MW_PER_ROW, wq = idx - yy * MW_PER_ROW;
    mv[i] = 0;
    if (yy < E) {
      int gy = y0 + yy - r;
      if (static_cast<unsigned>(gy) >= static_cast<unsigned>(H)) gy = reflect101(gy, H);
      const uint8_t* mrow = mask + static_cast<size_t>(gy) * W;
      const int gx = x0 - MF_R + 4 * wq;
      if (gx >= 0 && gx + 4 <= W) {
        mv[i] = __ldg(reinterpret_cast<const uint32_t*>(mrow + gx));
      } else {  // the word left of the first or right of the last column: mirrored bytes
        uint32_t w = 0;
        for (int b = 0; b < 4; ++b) w |= static_cast<uint32_t>(mrow[reflect101(gx + b, W)]) << (8 * b);
        mv[i] = w;
      }
    }
  }
  uint32_t all_and = 0xffffffffu, all_or = 0u;
#pragma unroll
  for (int i = 0; i < M_STEPS; ++i) {
    const int idx = tid + 256 * i;
    const int yy = idx / MW_PER_ROW, wq = idx - yy * MW_PER_ROW;
    if (yy < E) {
      const uint32_t on = __vcmpne4(mv[i], 0u);                                   // :80 mask * 255, per byte
      all_and &= on;
      all_or |= on;
      reinterpret_cast<uint32_t*>(sm + yy * MF_EW)[wq] = on;
    }
  }
  // A tile whose whole neighbourhood is inside (or outside) the mask blurs to 255 (0) everywhere: the blend returns the
  // style (content) pixel exactly, so the tile is a plain copy.  Most tiles of a segmentation mask are of this kind.
  const int any_off = __syncthreads_or(all_and != 0xffffffffu), any_on = __syncthreads_or(all_or != 0u);
  constexpr int Q_PER_ROW = MF_ROWB / 16, Q_STEPS = MF_T * Q_PER_ROW / 256;  // 12 uint4 per row, 3 per thread
  const size_t tile_off = static_cast<size_t>(y0) * rowb + static_cast<size_t>(x0) * 3;
  if (!(any_on && any_off)) {
    const uint8_t* src = (any_on ? style : content) + tile_off;
    uint4 v[Q_STEPS];
#pragma unroll
    for (int i = 0; i < Q_STEPS; ++i) {
      const int idx = tid + 256 * i;
      const int yy = idx / Q_PER_ROW, q = idx - yy * Q_PER_ROW;
      if (y0 + yy < H) v[i] = __ldg(reinterpret_cast<const uint4*>(src + yy * rowb) + q);
    }
#pragma unroll
    for (int i = 0; i < Q_STEPS; ++i) {
      const int idx = tid + 256 * i;
      const int yy = idx / Q_PER_ROW, q = idx - yy * Q_PER_ROW;
      if (y0 + yy < H) reinterpret_cast<uint4*>(out + tile_off + yy * rowb)[q] = v[i];
    }
    return;
  }
  // ---- mixed tile: both images into shared memory (loads in flight during the blur), blur, blend in place, store
  {
    uint4 vc[Q_STEPS], vs[Q_STEPS];
#pragma unroll
    for (int i = 0; i < Q_STEPS; ++i) {
      const int idx = tid + 256 * i;
      const int yy = idx / Q_PER_ROW, q = idx - yy * Q_PER_ROW;
      vc[i] = vs[i] = make_uint4(0, 0, 0, 0);
      if (y0 + yy < H) {
        vc[i] = __ldg(reinterpret_cast<const uint4*>(content + tile_off + yy * rowb) + q);
        vs[i] = __ldg(reinterpret_cast<const uint4*>(style + tile_off + yy * rowb) + q);
      }
    }
#pragma unroll
    for (int i = 0; i < Q_STEPS; ++i) {
      const int idx = tid + 256 * i;
      reinterpret_cast<uint4*>(sc)[idx] = vc[i];
      reinterpret_cast<uint4*>(ss)[idx] = vs[i];
    }
  }
  __shared__ double s_nb[256], s_na[256];
  {
    const double nb = __ddiv_rn(static_cast<double>(tid), 255.0);                  // :86
    s_nb[tid] = nb;
    s_na[tid] = __dsub_rn(1.0, nb);
  }
  const int tx = tid & 63, tg = tid >> 6;  // column, row group
  for (int yy = tg; yy < E; yy += 4) {
    int acc = 0;
    for (int j = 0; j < k; ++j) acc += mw.w[j] * sm[yy * MF_EW + tx + (MF_R - r) + j];
    sh[yy * MF_T + tx] = static_cast<uint16_t>(acc);
  }
  __syncthreads();
  // blend: see the general kernel for why the integer quotient is the reference's fp64 result unless the division is exact
  for (int yy = tg; yy < MF_T; yy += 4) {
    int acc = 0;
    for (int j = 0; j < k; ++j) acc += mw.w[j] * sh[(yy + j) * MF_T + tx];
    const int m = (acc + 32768) >> 16;                                             // :83 blurred mask value
    uint8_t* cb = sc + yy * MF_ROWB + tx * 3;
    const uint8_t* sb = ss + yy * MF_ROWB + tx * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int cv = cb[c], sv = sb[c];
      int res;
      if (m == 0) {
        res = cv;
      } else if (m == 255) {
        res = sv;
      } else {
        const int num = 255 * cv + m * (sv - cv);
        const int q = num / 255;
        if (num - 255 * q != 0) {
          res = q;
        } else {
          const double v = __dadd_rn(__dmul_rn(static_cast<double>(cv), s_na[m]), __dmul_rn(static_cast<double>(sv), s_nb[m]));
          res = static_cast<int>(v);                                               // :91 astype(np.uint8) truncates
        }
      }
      cb[c] = static_cast<uint8_t>(res);
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < Q_STEPS; ++i) {
    const int idx = tid + 256 * i;
    const int yy = idx / Q_PER_ROW, q = idx - yy * Q_PER_ROW;
    if (y0 + yy < H) reinterpret_cast<uint4*>(out + tile_off + yy * rowb)[q] = reinterpret_cast<const uint4*>(sc)[idx];
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
  const uintptr_t al = reinterpret_cast<uintptr_t>(content) | reinterpret_cast<uintptr_t>(style) | reinterpret_cast<uintptr_t>(out);
  if (C == 3 && k <= 2 * MF_R + 1 && (W % MF_T) == 0 && (al & 15) == 0 && (reinterpret_cast<uintptr_t>(mask) & 3) == 0) {
    dim3 grid(W / MF_T, (H + MF_T - 1) / MF_T);
    mask_composite_fast_kernel<<<grid, 256, MF_SMEM, s>>>(content, style, mask, out, H, W, mw);
    return cudaGetLastError();
  }
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
