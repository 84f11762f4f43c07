// Epilogue building blocks of the tcgen05 convolution kernels (conv_tc.cu: one launch per layer; conv_chain.cu: a whole
// chain of layers in one persistent launch).  One thread = one pixel of the 16 x 8 tile.
#pragma once
#include "conv_tc.cuh"
#include "common.cuh"

namespace nst {

// ---------------------------------------------------------------------------------------------
// epilogue helpers (one thread = one pixel of the tile, 32 consecutive channels per call)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256).  One thread owns one pixel, so the lanes of a warp are a whole pixel
// row (>= 128 bytes) apart: with 16-byte accesses every instruction touches half of 32 different 32-byte sectors and a
// second instruction the other halves; with 32-byte accesses every lane reads / writes one whole sector.
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void ld_global_nc_256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void ld_global_cg_256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void store_h32(__half* dst, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; q += 2) {
    uint4 u[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      u[e].x = pack_h2(v[8 * (q + e) + 0], v[8 * (q + e) + 1]);
      u[e].y = pack_h2(v[8 * (q + e) + 2], v[8 * (q + e) + 3]);
      u[e].z = pack_h2(v[8 * (q + e) + 4], v[8 * (q + e) + 5]);
      u[e].w = pack_h2(v[8 * (q + e) + 6], v[8 * (q + e) + 7]);
    }
    st_global_256(reinterpret_cast<uint8_t*>(dst) + 16 * q, u[0], u[1]);
  }
}
template <int CH>
__device__ __forceinline__ void store_bf(__nv_bfloat16* dst, const float (&v)[CH]) {
  static_assert(CH % 16 == 0, "32-byte stores");
#pragma unroll
  for (int q = 0; q < CH / 8; q += 2) {
    uint4 u[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      u[e].x = pack_bf2(v[8 * (q + e) + 0], v[8 * (q + e) + 1]);
      u[e].y = pack_bf2(v[8 * (q + e) + 2], v[8 * (q + e) + 3]);
      u[e].z = pack_bf2(v[8 * (q + e) + 4], v[8 * (q + e) + 5]);
      u[e].w = pack_bf2(v[8 * (q + e) + 6], v[8 * (q + e) + 7]);
    }
    st_global_256(reinterpret_cast<uint8_t*>(dst) + 16 * q, u[0], u[1]);
  }
}

// ---------------------------------------------------------------------------------------------
// Output through shared memory + TMA store.  One thread owns one pixel, so a direct 16-byte store instruction of a warp
// touches 32 different 128-byte lines (one piece each): the epilogue of the shallow layers was bound by those
// transactions, not by bandwidth (profiles/r01_conv_phases_store_experiment.log).  Instead every lane writes its pixel's
// 64 (or 32) bytes into a small per-warp staging buffer in the layout of a TMA box [4 rows][8 pixels][channels]
// (64- or 32-byte swizzle, which also makes the 16-byte shared-memory writes conflict-free) and one lane issues a
// cp.async.bulk.tensor store: coalesced by the TMA engine, clipped at the image border, asynchronous.
// Two buffers per warp alternate; a buffer is rewritten only after the bulk store that read it has finished reading.
// ---------------------------------------------------------------------------------------------
static constexpr int EPI_STAGE_BYTES = 2048;  // 32 pixels x 64 B; two per warp (the pool-routing scatter uses both as one 4 KB box)
struct EpiStore {
  uint8_t* buf;   // 2 x EPI_STAGE_BYTES of this warp, 1024-byte aligned
  int next;       // buffer the next store uses
  uint8_t* cur;   // buffer being filled by a store that spans two calls (data gradient: two 16-channel chunks per row)
  int lane;
  int w0, h0;     // tile-row origin of this warp's 8 x 4 pixel block inside its image
  int b;          // image of a batched launch (the output maps are 4-D (C, W, H, image): boxes are clipped per image)
};
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// rows of 64 bytes, 64-byte swizzle: 16-byte piece j of row r lives at piece j ^ ((r >> 1) & 3)
__device__ __forceinline__ void stage_row64(uint8_t* buf, int row, const uint4 (&u)[4]) {
  const int x = (row >> 1) & 3;
  uint4* b = reinterpret_cast<uint4*>(buf + row * 64);
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j ^ x] = u[j];
}
// rows of 32 bytes, 32-byte swizzle: piece j of row r lives at piece j ^ ((r >> 2) & 1)
__device__ __forceinline__ void stage_row32(uint8_t* buf, int row, const uint4 (&u)[2]) {
  const int x = (row >> 2) & 1;
  uint4* b = reinterpret_cast<uint4*>(buf + row * 32);
  b[x] = u[0];
  b[x ^ 1] = u[1];
}
// the two halves of a warp-wide TMA store: acquire a buffer (its previous store must have finished reading it) ...
__device__ __forceinline__ uint8_t* epi_acquire(EpiStore& st, bool both) {
  if (st.lane == 0) {
    if (both) bulk_wait_read<0>();
    else bulk_wait_read<1>();
  }
  __syncwarp();
  uint8_t* b = both ? st.buf : st.buf + st.next * EPI_STAGE_BYTES;
  if (!both) st.next ^= 1;
  return b;
}
// ... and, once every lane has written its rows, hand it to the TMA engine
__device__ __forceinline__ void epi_submit(const EpiStore& st, const uint8_t* b, const CUtensorMap* tm, int c0, int w0, int h0,
                                           bool skip) {
  fence_async_smem();
  __syncwarp();
  if (st.lane == 0 && !skip) {
    tma_store_4d(tm, b, c0, w0, h0, st.b);
    bulk_commit();
  }
}
__device__ __forceinline__ void pack_h32(const float (&v)[32], uint4 (&u)[4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    u[q].x = pack_h2(v[8 * q + 0], v[8 * q + 1]);
    u[q].y = pack_h2(v[8 * q + 2], v[8 * q + 3]);
    u[q].z = pack_h2(v[8 * q + 4], v[8 * q + 5]);
    u[q].w = pack_h2(v[8 * q + 6], v[8 * q + 7]);
  }
}
template <int CH>
__device__ __forceinline__ void pack_bf(const float (&v)[CH], uint4 (&u)[CH / 8]) {
#pragma unroll
  for (int q = 0; q < CH / 8; ++q) {
    u[q].x = pack_bf2(v[8 * q + 0], v[8 * q + 1]);
    u[q].y = pack_bf2(v[8 * q + 2], v[8 * q + 3]);
    u[q].z = pack_bf2(v[8 * q + 4], v[8 * q + 5]);
    u[q].w = pack_bf2(v[8 * q + 6], v[8 * q + 7]);
  }
}

// forward epilogue for 32 channels of one pixel
__device__ __forceinline__ void epilogue_fwd(const ConvParams& p, float (&v)[32], int h, int w, int n, bool valid,
                                             int lane, const float* sbias /* 32 values of this chunk, shared memory */,
                                             EpiStore* st = nullptr /* non-null: outputs go through TMA stores */) {
  // bias (same address across the warp -> broadcast read); staged in shared memory while the main loop ran
  const float4* b4 = reinterpret_cast<const float4*>(sbias);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 b = b4[q];
    v[4 * q + 0] += b.x;
    v[4 * q + 1] += b.y;
    v[4 * q + 2] += b.z;
    v[4 * q + 3] += b.w;
  }
  const size_t pix = static_cast<size_t>(h) * p.W + w;
  if (st != nullptr) {
    if (p.out_tap != nullptr) {
      uint4 u[4];
      pack_h32(v, u);
      uint8_t* b = epi_acquire(*st, false);
      stage_row64(b, lane, u);
      epi_submit(*st, b, &p.tmO0, n, st->w0, st->h0, NST_DBG_FLAG(p, 1));
    }
  } else if (p.out_tap != nullptr && valid && !NST_DBG_FLAG(p, 1)) {
    store_h32(p.out_tap + pix * p.N + n, v);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.f;
  if (!p.pool) {
    if (st != nullptr) {
      if (p.out_act != nullptr) {
        uint4 u[4];
        pack_h32(v, u);
        uint8_t* b = epi_acquire(*st, false);
        stage_row64(b, lane, u);
        epi_submit(*st, b, &p.tmO1, n, st->w0, st->h0, NST_DBG_FLAG(p, 1));
      }
    } else if (p.out_act != nullptr && valid && !NST_DBG_FLAG(p, 1)) {
      store_h32(p.out_act + pix * p.N + n, v);
    }
    return;
  }
  // 2x2 max-pool across the four lanes {lane, lane^1, lane^8, lane^9}: tile rows are 8 pixels
  // wide and a warp owns four consecutive rows.  The window position is folded into the two low
  // mantissa bits so that one integer max gives both the value and PyTorch's first-max arg-max.
  const uint32_t pos = ((lane >> 3) & 1) * 2 + (lane & 1);
  float pooled[32];
  uint32_t win[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    uint32_t key = (__float_as_uint(v[j]) & ~3u) | (3u - pos);
    key = max(key, __shfl_xor_sync(0xffffffffu, key, 1));
    key = max(key, __shfl_xor_sync(0xffffffffu, key, 8));
    pooled[j] = __uint_as_float(key & ~3u);
    win[j] = pooled[j] > 0.f ? 3u - (key & 3u) : 4u;
  }
  const int Hp = p.H >> 1, Wp = p.W >> 1;
  const int hp = h >> 1, wp = w >> 1;
  // several images per launch: H and W are even (multiples of 16) and h counts rows of the stacked [batch * H] view
  if (p.batch > 1 ? valid : (hp < Hp && wp < Wp)) {
    const size_t ppix = static_cast<size_t>(hp) * Wp + wp;
    // each of the four lanes of a window stores 8 of the 32 channels
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      if (pos == static_cast<uint32_t>(jj)) {
        uint4 u;
        u.x = pack_h2(pooled[8 * jj + 0], pooled[8 * jj + 1]);
        u.y = pack_h2(pooled[8 * jj + 2], pooled[8 * jj + 3]);
        u.z = pack_h2(pooled[8 * jj + 4], pooled[8 * jj + 5]);
        u.w = pack_h2(pooled[8 * jj + 6], pooled[8 * jj + 7]);
        *reinterpret_cast<uint4*>(p.out_act + ppix * p.N + n + 8 * jj) = u;
        uint2 r;
        r.x = win[8 * jj + 0] | (win[8 * jj + 1] << 8) | (win[8 * jj + 2] << 16) | (win[8 * jj + 3] << 24);
        r.y = win[8 * jj + 4] | (win[8 * jj + 5] << 8) | (win[8 * jj + 6] << 16) | (win[8 * jj + 7] << 24);
        *reinterpret_cast<uint2*>(p.out_route + ppix * p.N + n + 8 * jj) = r;
      }
    }
  }
}

// Operands of the data-gradient epilogue for DG_CH channels of one pixel (mask + tap seed, or pool routing bytes).
// They do not depend on the accumulator, so they are fetched one chunk ahead - the first one before the
// accumulator is even complete - instead of serialising global-memory latencies on the critical path of a tile.
static constexpr int DG_CH = 16;
struct DgradAux {
  uint4 m[2];  // post-ReLU activation (fp16) whose sign masks the gradient   | m[0]: routing bytes
  uint4 a[2];  // tap seed (bf16) added to the gradient
};
__device__ __forceinline__ void dgrad_aux_load(const ConvParams& p, DgradAux& x, int h, int w, int n, bool valid) {
  if (!valid || NST_DBG_FLAG(p, 2)) return;
  const size_t pix = static_cast<size_t>(h) * p.W + w;
  if (p.route == nullptr) {
    ld_global_nc_256(p.mask_act + pix * p.N + n, x.m[0], x.m[1]);
    if (p.addend != nullptr) {
      // L2-coherent load: in the chained kernel the seed is written earlier in the same launch (by another SM)
      ld_global_cg_256(p.addend + pix * p.N + n, x.a[0], x.a[1]);
    }
  } else {
    x.m[0] = __ldg(reinterpret_cast<const uint4*>(p.route + pix * p.N + n));
  }
}

// data-gradient epilogue for DG_CH channels of one pixel
__device__ __forceinline__ void epilogue_dgrad(const ConvParams& p, float (&v)[DG_CH], int h, int w, int n, bool valid,
                                               const DgradAux& x, EpiStore* st = nullptr,
                                               const float* seed = nullptr /* DG_CH values: folded Gram backward, already scaled */) {
  if (!valid && st == nullptr) return;  // with TMA stores every lane takes part (the box is clipped at the border)
  const size_t pix = static_cast<size_t>(h) * p.W + w;
  if (p.route == nullptr) {
    // ReLU mask from the stored post-ReLU activation (PyTorch: grad * (result > 0))
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const __half2* hh = reinterpret_cast<const __half2*>(&x.m[q]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = __half22float2(hh[e]);
        if (!(f.x > 0.f)) v[8 * q + 2 * e] = 0.f;
        if (!(f.y > 0.f)) v[8 * q + 2 * e + 1] = 0.f;
      }
    }
    if (seed != nullptr) {
#pragma unroll
      for (int j = 0; j < DG_CH; ++j) v[j] += seed[j];
    }
    if (p.addend != nullptr) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const __nv_bfloat162* bb = reinterpret_cast<const __nv_bfloat162*>(&x.a[q]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = __bfloat1622float2(bb[e]);
          v[8 * q + 2 * e] += f.x;
          v[8 * q + 2 * e + 1] += f.y;
        }
      }
    }
    if (st != nullptr) {
      // two consecutive 16-channel chunks fill one 64-byte row; the store goes out with the second (a 1 KB store per
      // chunk was slower than the direct stores: twice the bulk operations, each waiting for a buffer)
      uint4 u[DG_CH / 8];
      pack_bf<DG_CH>(v, u);
      const int half = (n / DG_CH) & 1;
      if (half == 0) st->cur = epi_acquire(*st, false);
      const int x = (st->lane >> 1) & 3;
      uint4* row = reinterpret_cast<uint4*>(st->cur + st->lane * 64);
      row[(2 * half) ^ x] = u[0];
      row[(2 * half + 1) ^ x] = u[1];
      if (half == 1) epi_submit(*st, st->cur, &p.tmO0, n - DG_CH, st->w0, st->h0, NST_DBG_FLAG(p, 1));
    } else if (!NST_DBG_FLAG(p, 1)) {
      store_bf<DG_CH>(p.out_grad + pix * p.N + n, v);
    }
  } else {
    // max-pool routing: the gradient of pooled pixel (h, w) goes to the arg-max position of its
    // 2x2 window in the un-pooled map (and only if the pooled activation was > 0: ReLU mask).
    uint32_t r[4];
    r[0] = x.m[0].x; r[1] = x.m[0].y; r[2] = x.m[0].z; r[3] = x.m[0].w;
    // TMA path: the warp's 8 x 4 pooled pixels scatter into one 16 x 8 box of the un-pooled gradient (4 KB, both buffers)
    uint8_t* b = st != nullptr ? epi_acquire(*st, true) : nullptr;
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      float o[DG_CH];
#pragma unroll
      for (int j = 0; j < DG_CH; ++j) {
        const uint32_t rj = (r[j >> 2] >> (8 * (j & 3))) & 0xffu;
        o[j] = (valid && rj == static_cast<uint32_t>(pos)) ? v[j] : 0.f;
      }
      if (st != nullptr) {
        uint4 u[DG_CH / 8];
        pack_bf<DG_CH>(o, u);
        const int row = (2 * (st->lane >> 3) + (pos >> 1)) * 16 + 2 * (st->lane & 7) + (pos & 1);
        stage_row32(b, row, u);
      } else {
        const size_t upix = static_cast<size_t>(2 * h + (pos >> 1)) * p.Wup + (2 * w + (pos & 1));
        if (!NST_DBG_FLAG(p, 1)) store_bf<DG_CH>(p.out_grad + upix * p.N + n, o);
      }
    }
    if (st != nullptr) epi_submit(*st, b, &p.tmO0, n, 2 * st->w0, 2 * st->h0, NST_DBG_FLAG(p, 1));
  }
}

__device__ __forceinline__ void epilogue_scale(const ConvParams& p, float (&v)[32], int h, int w, int n, bool valid,
                                               float alpha, EpiStore* st = nullptr) {
  if (!valid && st == nullptr) return;
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] *= alpha;
  if (st != nullptr) {
    uint4 u[4];
    pack_bf<32>(v, u);
    uint8_t* b = epi_acquire(*st, false);
    stage_row64(b, st->lane, u);
    epi_submit(*st, b, &p.tmO0, n, st->w0, st->h0, NST_DBG_FLAG(p, 1));
    return;
  }
  const size_t pix = static_cast<size_t>(h) * p.W + w;
  store_bf<32>(p.out_grad + pix * p.N + n, v);
}

}  // namespace nst
