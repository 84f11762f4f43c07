// Video back end: the frame list that apply_video_process hands to the encoder.
//
// Replaces  app.py:800-806  (cv2.cvtColor(frame, COLOR_RGB2BGR) of every stylised frame) and  app.py:820-840  (the
// cross-dissolve: between two consecutive frames, number_of_interpolations frames
// cv2.addWeighted(prev, 1 - alpha, frame, alpha, 0) with alpha = (i + 1) / (n + 1)) in ONE pass over the stylised frames
// while they are still on the device: each input frame is read twice (once as `prev`, once as `frame`), each output frame
// written once.  Bandwidth bound.
//
// Bit-exact with OpenCV 4.x on a CPU with FMA (every x86-64 since 2013; the dispatch this image's cv2 4.13 takes):
// addWeighted on 8-bit data converts alpha, beta to fp32 and computes  fma(a, alpha, b * beta)  in fp32, rounds to nearest
// even and saturates (pinned to cv2 on all 65 536 byte pairs for every slider value by the CPU tests).
#include "video.cuh"

namespace nst {

struct VideoWeights {
  int n;
  float a1[VIDEO_MAX_INTERP], a2[VIDEO_MAX_INTERP];  // (float)(1 - alpha_i), (float)alpha_i
};

// byte `pos` of w as a float, without the conversion unit: 2^23 + byte is a float whose bit pattern is 0x4B0000bb
__device__ __forceinline__ float byte_as_float(uint32_t w, int pos) {
  return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | static_cast<uint32_t>(pos))), 8388608.0f);
}
// round-to-nearest-even of 0 <= v < 2^22 in the low bits of v + 1.5 * 2^23
__device__ __forceinline__ uint32_t round_bits(float v) { return __float_as_uint(__fadd_rn(v, 12582912.0f)); }
__device__ __forceinline__ uint32_t pack4(float v0, float v1, float v2, float v3) {
  const uint32_t lo = __byte_perm(round_bits(v0), round_bits(v1), 0x0040u);
  const uint32_t hi = __byte_perm(round_bits(v2), round_bits(v3), 0x0040u);
  return __byte_perm(lo, hi, 0x5410u);
}

// One thread: 16 pixels = 48 bytes = three 16-byte words of `prev` and of `frame`, all six loads in flight at once.
__global__ void __launch_bounds__(256) video_assemble_kernel(const uint8_t* __restrict__ frames, uint8_t* __restrict__ out,
                                                             size_t frame_bytes, int F, const __grid_constant__ VideoWeights vw) {
  // a thread's 48 output bytes go through shared memory so that every store instruction of a warp writes 512 contiguous
  // bytes: 16-byte stores 48 bytes apart are half-sector writes and cost a third of the bandwidth (measured on the plane split)
  __shared__ uint4 stage[8][96];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int f = blockIdx.y;
  const size_t chunks = frame_bytes / 48;
  const size_t chunk = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  const bool live = chunk < chunks;
  if (chunk - lane >= chunks) return;  // whole warp past the end
  const int n = (f + 1 < F) ? vw.n : 0;  // the last frame has no successor
  const size_t ld = live ? chunk : chunks - 1;  // lanes past the end load the last chunk and store nothing
  const uint4* pa = reinterpret_cast<const uint4*>(frames + static_cast<size_t>(f) * frame_bytes) + 3 * ld;
  const uint4* pb = reinterpret_cast<const uint4*>(frames + static_cast<size_t>(f + (n ? 1 : 0)) * frame_bytes) + 3 * ld;
  uint32_t wa[12], wb[12];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const uint4 v = __ldg(pa + q);
    wa[4 * q] = v.x, wa[4 * q + 1] = v.y, wa[4 * q + 2] = v.z, wa[4 * q + 3] = v.w;
  }
  if (n) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const uint4 v = __ldg(pb + q);
      wb[4 * q] = v.x, wb[4 * q + 1] = v.y, wb[4 * q + 2] = v.z, wb[4 * q + 3] = v.w;
    }
  }
  uint8_t* po = out + static_cast<size_t>(f) * (vw.n + 1) * frame_bytes;
  const size_t q0 = 3 * (chunk - lane), q_end = 3 * chunks;  // the warp's first 16-byte word, the end of the frame
  auto store_frame = [&](uint8_t* frame_out, const uint32_t (&w)[12]) {
    stage[warp][3 * lane] = make_uint4(w[0], w[1], w[2], w[3]);
    stage[warp][3 * lane + 1] = make_uint4(w[4], w[5], w[6], w[7]);
    stage[warp][3 * lane + 2] = make_uint4(w[8], w[9], w[10], w[11]);
    __syncwarp();
    uint4* dst = reinterpret_cast<uint4*>(frame_out) + q0;
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (q0 + 32 * j + lane < q_end) dst[32 * j + lane] = stage[warp][32 * j + lane];
    __syncwarp();
  };
  // output byte j comes from input byte j - (j % 3) + 2 - (j % 3): RGB -> BGR          app.py:803
  float fa[48];
#pragma unroll
  for (int j = 0; j < 48; ++j) {
    const int sj = j - (j % 3) + 2 - (j % 3);
    fa[j] = byte_as_float(wa[sj >> 2], sj & 3);
  }
  {
    uint32_t w[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) w[e] = pack4(fa[4 * e], fa[4 * e + 1], fa[4 * e + 2], fa[4 * e + 3]);
    store_frame(po, w);
  }
  if (n == 0) return;
  float fb[48];
#pragma unroll
  for (int j = 0; j < 48; ++j) {
    const int sj = j - (j % 3) + 2 - (j % 3);
    fb[j] = byte_as_float(wb[sj >> 2], sj & 3);
  }
  for (int i = 0; i < n; ++i) {
    const float a1 = vw.a1[i], a2 = vw.a2[i];
    uint32_t w[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) {
      const int j = 4 * e;
      w[e] = pack4(__fmaf_rn(fa[j], a1, __fmul_rn(fb[j], a2)), __fmaf_rn(fa[j + 1], a1, __fmul_rn(fb[j + 1], a2)),
                   __fmaf_rn(fa[j + 2], a1, __fmul_rn(fb[j + 2], a2)), __fmaf_rn(fa[j + 3], a1, __fmul_rn(fb[j + 3], a2)));  // :833
    }
    store_frame(po + static_cast<size_t>(i + 1) * frame_bytes, w);
  }
}

// any size / alignment: one thread per pixel
__global__ void video_assemble_scalar_kernel(const uint8_t* __restrict__ frames, uint8_t* __restrict__ out, size_t pixels, int F,
                                             const __grid_constant__ VideoWeights vw) {
  const int f = blockIdx.y;
  const size_t px = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (px >= pixels) return;
  const size_t frame_bytes = pixels * 3;
  const int n = (f + 1 < F) ? vw.n : 0;
  const uint8_t* pa = frames + static_cast<size_t>(f) * frame_bytes + px * 3;
  const uint8_t* pb = pa + (n ? frame_bytes : 0);
  uint8_t* po = out + static_cast<size_t>(f) * (vw.n + 1) * frame_bytes + px * 3;
  float a[3], b[3];
  for (int c = 0; c < 3; ++c) {
    a[c] = static_cast<float>(pa[2 - c]);
    b[c] = static_cast<float>(pb[2 - c]);
    po[c] = pa[2 - c];
  }
  for (int i = 0; i < n; ++i) {
    po += frame_bytes;
    for (int c = 0; c < 3; ++c) {
      const int r = __float2int_rn(__fmaf_rn(a[c], vw.a1[i], __fmul_rn(b[c], vw.a2[i])));
      po[c] = static_cast<uint8_t>(min(max(r, 0), 255));
    }
  }
}

cudaError_t launch_video_assemble(const uint8_t* frames, int F, size_t pixels, int n_interp, uint8_t* out, cudaStream_t s) {
  if (F < 1 || pixels < 1 || n_interp < 0 || n_interp > VIDEO_MAX_INTERP) return cudaErrorInvalidValue;
  VideoWeights vw;
  vw.n = n_interp;
  for (int i = 0; i < n_interp; ++i) {
    const double alpha = static_cast<double>(i + 1) / static_cast<double>(n_interp + 1);  // app.py:829
    vw.a1[i] = static_cast<float>(1.0 - alpha);
    vw.a2[i] = static_cast<float>(alpha);
  }
  const size_t frame_bytes = pixels * 3;
  const uintptr_t al = reinterpret_cast<uintptr_t>(frames) | reinterpret_cast<uintptr_t>(out);
  if ((al & 15) == 0 && (frame_bytes % 48) == 0) {
    const size_t chunks = frame_bytes / 48;
    dim3 grid(static_cast<unsigned>((chunks + 255) / 256), F);
    video_assemble_kernel<<<grid, 256, 0, s>>>(frames, out, frame_bytes, F, vw);
  } else {
    dim3 grid(static_cast<unsigned>((pixels + 255) / 256), F);
    video_assemble_scalar_kernel<<<grid, 256, 0, s>>>(frames, out, pixels, F, vw);
  }
  return cudaGetLastError();
}

}  // namespace nst
