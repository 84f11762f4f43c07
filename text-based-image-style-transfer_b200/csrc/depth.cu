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

static bool aligned4(const void* a, const void* b) { return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 3) == 0; }

cudaError_t launch_mip_split(const uint8_t* image, const void* depth, int depth_f64, size_t pixels, int C, const MipBins& bins,
                             uint8_t* out, cudaStream_t s) {
  if (bins.n < 1 || bins.n > MIP_MAX_PLANES || pixels < 1 || C < 1) return cudaErrorInvalidValue;
  size_t done = 0;
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
  if (aligned4(planes, out) && (pixels % 4) == 0) {
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
