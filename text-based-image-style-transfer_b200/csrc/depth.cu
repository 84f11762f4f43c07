// Multi-plane ("MIP") depth style transfer: the byte work on either side of the n style-transfer loops.
//
// Replaces  components/style_transfer_depth/util.py:9-35  (mask_image_depth),  :53-66  (generate_mip_layers) and  :69-88
// (reconstruct_mip_image).  The reference normalises the depth map (depth - min) / (max - min) in fp64, keeps the pixels with
// lo <= depth <= hi (both inclusive, so a pixel exactly on an inner bin edge belongs to two planes) and zeroes the others;
// the reconstruction masks every stylised plane with its bin again and adds the planes up in uint8, wrapping around
// where two planes overlap (mip += img on a uint8 array, util.py:85-87).  Both are kept bit for bit.
// Bandwidth bound: split reads (C + 1) bytes and writes n * C per pixel; merge reads 1 + ~3 (only the planes a pixel group
// belongs to are loaded) and writes 3.
#include "depth.cuh"

namespace nst {

template <bool F64>
__device__ __forceinline__ double norm_depth(const void* depth, size_t i, const MipBins& b) {
  if (F64) return static_cast<const double*>(depth)[i];
  const int d = static_cast<const uint8_t*>(depth)[i];
  return __ddiv_rn(static_cast<double>(d - b.dmin), static_cast<double>(b.range));            // util.py:27 (0 / 0 = nan: no plane)
}

// bit i of the result: pixel belongs to plane i
__device__ __forceinline__ uint32_t plane_bits(double d, const MipBins& b) {
  uint32_t m = 0;
  for (int i = 0; i < b.n; ++i) m |= (d >= b.lo[i] && d <= b.hi[i]) ? (1u << i) : 0u;           // util.py:30
  return m;
}

// the three 32-bit words of four RGB pixels: byte masks from four per-pixel flags
__device__ __forceinline__ void rgb4_masks(bool p0, bool p1, bool p2, bool p3, uint32_t m[3]) {
  m[0] = (p0 ? 0x00FFFFFFu : 0u) | (p1 ? 0xFF000000u : 0u);
  m[1] = (p1 ? 0x0000FFFFu : 0u) | (p2 ? 0xFFFF0000u : 0u);
  m[2] = (p2 ? 0x000000FFu : 0u) | (p3 ? 0xFFFFFF00u : 0u);
}

// ---- four RGB pixels per thread, 32-bit accesses
template <bool F64>
__global__ void __launch_bounds__(256) mip_split_rgb4_kernel(const uint8_t* __restrict__ image, const void* __restrict__ depth, size_t groups,
                                                             size_t pixels, const __grid_constant__ MipBins b, uint8_t* __restrict__ out) {
  const size_t g = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (g >= groups) return;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(image) + 3 * g;
  const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
  uint32_t bits[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) bits[p] = plane_bits(norm_depth<F64>(depth, 4 * g + p, b), b);
  for (int i = 0; i < b.n; ++i) {
    uint32_t m[3];
    rgb4_masks((bits[0] >> i) & 1, (bits[1] >> i) & 1, (bits[2] >> i) & 1, (bits[3] >> i) & 1, m);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + static_cast<size_t>(i) * pixels * 3) + 3 * g;
    dst[0] = w0 & m[0];                                                                          // util.py:32-33
    dst[1] = w1 & m[1];
    dst[2] = w2 & m[2];
  }
}

template <bool F64>
__global__ void __launch_bounds__(256) mip_merge_rgb4_kernel(const uint8_t* __restrict__ planes, const void* __restrict__ depth, size_t groups,
                                                             size_t pixels, const __grid_constant__ MipBins b, uint8_t* __restrict__ out) {
  const size_t g = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (g >= groups) return;
  uint32_t bits[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) bits[p] = plane_bits(norm_depth<F64>(depth, 4 * g + p, b), b);
  const uint32_t any = bits[0] | bits[1] | bits[2] | bits[3];
  uint32_t acc0 = 0, acc1 = 0, acc2 = 0;
  for (int i = 0; i < b.n; ++i) {
    if (!((any >> i) & 1)) continue;  // none of the four pixels lies in plane i: its bytes are not read
    uint32_t m[3];
    rgb4_masks((bits[0] >> i) & 1, (bits[1] >> i) & 1, (bits[2] >> i) & 1, (bits[3] >> i) & 1, m);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(planes + static_cast<size_t>(i) * pixels * 3) + 3 * g;
    acc0 = __vadd4(acc0, __ldg(src) & m[0]);                                                     // util.py:86-87: uint8 +=, wraps
    acc1 = __vadd4(acc1, __ldg(src + 1) & m[1]);
    acc2 = __vadd4(acc2, __ldg(src + 2) & m[2]);
  }
  uint32_t* dst = reinterpret_cast<uint32_t*>(out) + 3 * g;
  dst[0] = acc0, dst[1] = acc1, dst[2] = acc2;
}

// ---- uint8 depth, sixteen RGB pixels per thread, 16-byte accesses.  A uint8 map has 256 possible depths: their plane bits
// are tabulated once on the host (256 fp64 divisions instead of one per pixel) and copied to shared memory by every block;
// a thread looks its sixteen depth bytes up and has all its loads in flight before the first use.
struct MipTable {
  uint32_t bits[256];  // plane bits of every uint8 depth, built on the host with the same fp64 division and comparisons
};
__device__ __forceinline__ void mip_table(uint32_t* tab, const MipTable& t) {
  tab[threadIdx.x] = t.bits[threadIdx.x];
  __syncthreads();
}

__global__ void __launch_bounds__(256) mip_split_rgb16_kernel(const uint8_t* __restrict__ image, const uint8_t* __restrict__ depth,
                                                              size_t chunks, size_t pixels, int n, const __grid_constant__ MipTable b,
                                                              uint8_t* __restrict__ out) {
  __shared__ uint32_t tab[256];
  __shared__ uint4 stage[8][96];  // per warp: its 32 x 48 output bytes, re-read so that a store instruction writes 512 contiguous bytes
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t c = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  const bool live = c < chunks;
  uint4 dv = make_uint4(0, 0, 0, 0), v0 = dv, v1 = dv, v2 = dv;
  if (live) {  // in flight while the table is copied
    const uint4* src = reinterpret_cast<const uint4*>(image) + 3 * c;
    dv = __ldg(reinterpret_cast<const uint4*>(depth) + c);
    v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2);
  }
  mip_table(tab, b);
  const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
  const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
  uint32_t bits[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) bits[p] = tab[(dw[p >> 2] >> (8 * (p & 3))) & 0xFFu];
  const size_t q0 = 3 * (c - lane), q_end = 3 * chunks;  // the warp's first 16-byte word, the end of the plane
  for (int i = 0; i < n; ++i) {
    uint32_t o[12];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t m[3];
      rgb4_masks((bits[4 * q] >> i) & 1, (bits[4 * q + 1] >> i) & 1, (bits[4 * q + 2] >> i) & 1, (bits[4 * q + 3] >> i) & 1, m);
      o[3 * q] = w[3 * q] & m[0], o[3 * q + 1] = w[3 * q + 1] & m[1], o[3 * q + 2] = w[3 * q + 2] & m[2];          // util.py:32-33
    }
    stage[warp][3 * lane] = make_uint4(o[0], o[1], o[2], o[3]);
    stage[warp][3 * lane + 1] = make_uint4(o[4], o[5], o[6], o[7]);
    stage[warp][3 * lane + 2] = make_uint4(o[8], o[9], o[10], o[11]);
    __syncwarp();
    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(i) * pixels * 3) + q0;
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (q0 + 32 * j + lane < q_end) dst[32 * j + lane] = stage[warp][32 * j + lane];
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) mip_merge_rgb16_kernel(const uint8_t* __restrict__ planes, const uint8_t* __restrict__ depth,
                                                              size_t chunks, size_t pixels, const __grid_constant__ MipTable b,
                                                              uint8_t* __restrict__ out) {
  __shared__ uint32_t tab[256];
  __shared__ uint4 stage[8][96];  // see mip_split_rgb16_kernel
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t c = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  const bool live = c < chunks;
  uint4 dv = make_uint4(0, 0, 0, 0);
  if (live) dv = __ldg(reinterpret_cast<const uint4*>(depth) + c);  // in flight while the table is built
  mip_table(tab, b);
  const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
  uint32_t bits[16], any = 0;
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    bits[p] = tab[(dw[p >> 2] >> (8 * (p & 3))) & 0xFFu];
    any |= bits[p];
  }
  if (!live) any = 0;  // lanes past the end read nothing and store nothing, but take part in the warp's staged store
  uint32_t acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0;
  // planes that hold none of the sixteen pixels are not read; neighbouring pixels nearly always share one plane
  while (any) {
    const int i = __ffs(any) - 1;
    any &= any - 1;
    const uint4* src = reinterpret_cast<const uint4*>(planes + static_cast<size_t>(i) * pixels * 3) + 3 * c;
    const uint4 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2);
    const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t m[3];
      rgb4_masks((bits[4 * q] >> i) & 1, (bits[4 * q + 1] >> i) & 1, (bits[4 * q + 2] >> i) & 1, (bits[4 * q + 3] >> i) & 1, m);
      acc[3 * q] = __vadd4(acc[3 * q], w[3 * q] & m[0]);                                        // util.py:86-87: uint8 +=, wraps
      acc[3 * q + 1] = __vadd4(acc[3 * q + 1], w[3 * q + 1] & m[1]);
      acc[3 * q + 2] = __vadd4(acc[3 * q + 2], w[3 * q + 2] & m[2]);
    }
  }
  stage[warp][3 * lane] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
  stage[warp][3 * lane + 1] = make_uint4(acc[4], acc[5], acc[6], acc[7]);
  stage[warp][3 * lane + 2] = make_uint4(acc[8], acc[9], acc[10], acc[11]);
  __syncwarp();
  const size_t q0 = 3 * (c - lane), q_end = 3 * chunks;
  uint4* dst = reinterpret_cast<uint4*>(out) + q0;
#pragma unroll
  for (int j = 0; j < 3; ++j)
    if (q0 + 32 * j + lane < q_end) dst[32 * j + lane] = stage[warp][32 * j + lane];
}

// ---- any channel count / size / alignment: one pixel per thread
template <bool F64>
__global__ void mip_split_px_kernel(const uint8_t* __restrict__ image, const void* __restrict__ depth, size_t first, size_t pixels, int C,
                                    const __grid_constant__ MipBins b, uint8_t* __restrict__ out) {
  const size_t px = first + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (px >= pixels) return;
  const uint32_t bits = plane_bits(norm_depth<F64>(depth, px, b), b);
  for (int i = 0; i < b.n; ++i)
    for (int c = 0; c < C; ++c)
      out[(static_cast<size_t>(i) * pixels + px) * C + c] = ((bits >> i) & 1) ? image[px * C + c] : 0;
}

template <bool F64>
__global__ void mip_merge_px_kernel(const uint8_t* __restrict__ planes, const void* __restrict__ depth, size_t first, size_t pixels,
                                    const __grid_constant__ MipBins b, uint8_t* __restrict__ out) {
  const size_t px = first + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (px >= pixels) return;
  const uint32_t bits = plane_bits(norm_depth<F64>(depth, px, b), b);
  for (int c = 0; c < 3; ++c) {
    uint32_t acc = 0;
    for (int i = 0; i < b.n; ++i)
      if ((bits >> i) & 1) acc += planes[(static_cast<size_t>(i) * pixels + px) * 3 + c];
    out[px * 3 + c] = static_cast<uint8_t>(acc);
  }
}

static MipTable host_table(const MipBins& b) {
  MipTable t;
  for (int v = 0; v < 256; ++v) {
    const volatile double num = static_cast<double>(v - b.dmin), den = static_cast<double>(b.range);
    const double d = num / den;                                                                  // util.py:27 (0 / 0 = nan: no plane)
    uint32_t m = 0;
    for (int i = 0; i < b.n; ++i) m |= (d >= b.lo[i] && d <= b.hi[i]) ? (1u << i) : 0u;          // util.py:30
    t.bits[v] = m;
  }
  return t;
}

static bool aligned16(const void* a, const void* b, const void* c) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}
static bool aligned4(const void* a, const void* b) { return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 3) == 0; }

cudaError_t launch_mip_split(const uint8_t* image, const void* depth, int depth_f64, size_t pixels, int C, const MipBins& bins,
                             uint8_t* out, cudaStream_t s) {
  if (bins.n < 1 || bins.n > MIP_MAX_PLANES || pixels < 1 || C < 1) return cudaErrorInvalidValue;
  size_t done = 0;
  if (C == 3 && !depth_f64 && aligned16(image, out, depth) && (pixels % 16) == 0) {
    const size_t chunks = pixels / 16;
    mip_split_rgb16_kernel<<<static_cast<unsigned>((chunks + 255) / 256), 256, 0, s>>>(image, static_cast<const uint8_t*>(depth), chunks,
                                                                                       pixels, bins.n, host_table(bins), out);
    return cudaGetLastError();
  }
  // every plane starts at a multiple of 4 bytes only if pixels * 3 is one: then groups of four pixels are whole words
  if (C == 3 && aligned4(image, out) && (pixels % 4) == 0) {
    const size_t groups = pixels / 4;
    const unsigned blocks = static_cast<unsigned>((groups + 255) / 256);
    if (depth_f64) mip_split_rgb4_kernel<true><<<blocks, 256, 0, s>>>(image, depth, groups, pixels, bins, out);
    else mip_split_rgb4_kernel<false><<<blocks, 256, 0, s>>>(image, depth, groups, pixels, bins, out);
    done = pixels;
  }
  if (done < pixels) {
    const unsigned blocks = static_cast<unsigned>((pixels - done + 255) / 256);
    if (depth_f64) mip_split_px_kernel<true><<<blocks, 256, 0, s>>>(image, depth, done, pixels, C, bins, out);
    else mip_split_px_kernel<false><<<blocks, 256, 0, s>>>(image, depth, done, pixels, C, bins, out);
  }
  return cudaGetLastError();
}

cudaError_t launch_mip_merge(const uint8_t* planes, const void* depth, int depth_f64, size_t pixels, const MipBins& bins, uint8_t* out,
                             cudaStream_t s) {
  if (bins.n < 1 || bins.n > MIP_MAX_PLANES || pixels < 1) return cudaErrorInvalidValue;
  if (!depth_f64 && aligned16(planes, out, depth) && (pixels % 16) == 0) {
    const size_t chunks = pixels / 16;
    mip_merge_rgb16_kernel<<<static_cast<unsigned>((chunks + 255) / 256), 256, 0, s>>>(planes, static_cast<const uint8_t*>(depth), chunks,
                                                                                       pixels, host_table(bins), out);
  } else if (aligned4(planes, out) && (pixels % 4) == 0) {
    const size_t groups = pixels / 4;
    const unsigned blocks = static_cast<unsigned>((groups + 255) / 256);
    if (depth_f64) mip_merge_rgb4_kernel<true><<<blocks, 256, 0, s>>>(planes, depth, groups, pixels, bins, out);
    else mip_merge_rgb4_kernel<false><<<blocks, 256, 0, s>>>(planes, depth, groups, pixels, bins, out);
  } else {
    const unsigned blocks = static_cast<unsigned>((pixels + 255) / 256);
    if (depth_f64) mip_merge_px_kernel<true><<<blocks, 256, 0, s>>>(planes, depth, 0, pixels, bins, out);
    else mip_merge_px_kernel<false><<<blocks, 256, 0, s>>>(planes, depth, 0, pixels, bins, out);
  }
  return cudaGetLastError();
}

}  // namespace nst
