// Bandwidth-bound vector kernels of the pixel-space L-BFGS update.
//
// Replaces the ~4m+15 small ATen launches and 4-5 host syncs per iteration of
//   torch.optim.LBFGS.step   torch/optim/lbfgs.py:388-526   (called from run_style_transfer.py:151)
// by two streaming passes over the stored (s, y) pairs plus a one-warp controller (lbfgs_ctl.h):
//   pass 1  y = g - g_prev, s = t d (written to the spare history slot), and every dot product of
//           {s, y, g} with every stored s_i / y_i, with max|g| and sum|g|          reads (2m + 3) n floats
//   pass 2  d = sum_k coef_k basis_k, x <- clamp(x + t d, 0, 1), g_prev <- g, max|t d|   reads (2m + 2) n floats
// The [0,1] clamp of run_style_transfer.py:108-109 is folded into the update.
// No host synchronisation: loop bounds and the stop flag are read from the device control block.
#include "lbfgs.cuh"
#include <stdlib.h>
#include "common.cuh"

namespace nst {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// streaming 128-bit load (evict-first).  Deliberately NOT `asm volatile`: the compiler must be free to hoist the loads
// of the next stored pair above the shuffle reduction of the current one (the volatile version serialised
// load -> reduce -> load and ran pass 1 at a quarter of the HBM bandwidth).
__device__ __forceinline__ float4 ld4_stream(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  acc = fmaf(a.w, b.w, acc);
  return acc;
}

// Reduces 4 per-lane values across the warp with 6 shuffles; on return lane 8*k (k = 0..3) holds value k in a[0].
__device__ __forceinline__ void warp_reduce4(float (&a)[4], int lane) {
  const bool up16 = (lane & 16) != 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float send = up16 ? a[k] : a[k + 2];
    const float keep = up16 ? a[k + 2] : a[k];
    a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  const bool up8 = (lane & 8) != 0;
  {
    const float send = up8 ? a[0] : a[1];
    const float keep = up8 ? a[1] : a[0];
    a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  a[0] += __shfl_xor_sync(0xffffffffu, a[0], 4);
  a[0] += __shfl_xor_sync(0xffffffffu, a[0], 2);
  a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
}

// Reduces 8 per-lane values across the warp with 9 shuffles (recursive halving);
// on return lane 4*k (k = 0..7) holds the warp-wide sum of value k in a[0].
__device__ __forceinline__ void warp_reduce8(float (&a)[8], int lane) {
  const bool up16 = (lane & 16) != 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float send = up16 ? a[k] : a[k + 4];
    const float keep = up16 ? a[k + 4] : a[k];
    a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  const bool up8 = (lane & 8) != 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float send = up8 ? a[k] : a[k + 2];
    const float keep = up8 ? a[k + 2] : a[k];
    a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const bool up4 = (lane & 4) != 0;
  {
    const float send = up4 ? a[0] : a[1];
    const float keep = up4 ? a[1] : a[0];
    a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  a[0] += __shfl_xor_sync(0xffffffffu, a[0], 2);
  a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
}

__global__ void lbfgs_step_begin_kernel(NstLbfgsCtl* ctl) {
  ctl->stop = ctl->frozen != 0 ? NST_STOP_FROZEN : NST_RUN;
  ctl->run_pass2 = 0;
}

// ------------------------------------------------------------------------------------------------
// pass 1
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LB_THREADS) lbfgs_pass1_kernel(const LbfgsBuffers b) {
  __shared__ float wpart[LB_THREADS / 32][NST_LBFGS_SLOTS][NST_LBFGS_NDOT];
  __shared__ float wscal[LB_THREADS / 32][NST_LBFGS_NSCAL];
  // Programmatic dependent launch (all four optimizer kernels): scheduled while the previous kernel drains.  The control
  // block and the stored pairs were written by earlier optimizer kernels (complete: every kernel of the chain waits for
  // its predecessor), only the gradient comes from the kernel right before this one - so the ring of stored pairs is
  // primed BEFORE griddepcontrol.wait, while the last data-gradient kernel is still running.
  const NstLbfgsCtl* ctl = b.ctl;
  if (ctl->stop != NST_RUN) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    return;
  }
  const int len = ctl->hist_len, head = ctl->hist_head;
  const bool first = ctl->n_iter == 0;
  const float t = static_cast<float>(ctl->t);
  const int pn = (head + len) % NST_LBFGS_SLOTS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  const int nv = b.n_pad >> 2;
  const int v0 = blockIdx.x * b.vec_per_blk;
  const int v1 = min(nv, v0 + b.vec_per_blk);
  const size_t row = static_cast<size_t>(b.vec_per_blk) * 4;  // floats per (slot, s|y) row of this block's region
  float* hblk = b.hist + static_cast<size_t>(blockIdx.x) * (2 * NST_LBFGS_SLOTS) * row;

  // The stored pairs stream through a ring of LB_RING shared-memory stages filled by bulk copies (see pass 2)
  extern __shared__ __align__(128) uint8_t lb_smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(lb_smem);
  uint64_t* empty_bar = full_bar + LB_RING;
  uint8_t* ring = lb_smem + 128;
  const uint32_t pair_bytes = static_cast<uint32_t>(row) * 8u;  // S row + Y row
  const uint32_t stage_bytes = (pair_bytes + 127u) & ~127u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < LB_RING; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], LB_THREADS / 32);
    }
    mbar_fence_init();
  }
  auto issue = [&](int i) {  // thread 0: pair i (age order) -> stage i % LB_RING
    int p = head + i;
    if (p >= NST_LBFGS_SLOTS) p -= NST_LBFGS_SLOTS;
    const int st = i % LB_RING;
    mbar_arrive_expect_tx(&full_bar[st], pair_bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(ring + st * stage_bytes)),
                 "l"(hblk + static_cast<size_t>(2 * p) * row), "r"(pair_bytes), "r"(smem_u32(&full_bar[st]))
                 : "memory");
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < LB_RING && i < len; ++i) issue(i);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");


  float4 s4[LB_VEC_PER_THREAD], y4[LB_VEC_PER_THREAD], g4[LB_VEC_PER_THREAD];
  size_t off[LB_VEC_PER_THREAD];   // offset in the flat vectors (x, g, d, ...)
  int hoff[LB_VEC_PER_THREAD];     // offset inside a history row
  bool ok[LB_VEC_PER_THREAD];
  float sc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float gmax = 0.f;
#pragma unroll
  for (int i = 0; i < LB_VEC_PER_THREAD; ++i) {
    const int lv = i * LB_THREADS + threadIdx.x;
    const int vi = v0 + lv;
    ok[i] = vi < v1;
    off[i] = static_cast<size_t>(ok[i] ? vi : v0) * 4;
    hoff[i] = (ok[i] ? lv : 0) * 4;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    g4[i] = ok[i] ? ld4(b.g + off[i]) : z;
    s4[i] = z;
    y4[i] = z;
    if (!first && ok[i]) {
      const float4 gp = ld4(b.g_prev + off[i]);
      const float4 dd = ld4(b.d + off[i]);
      y4[i] = make_float4(g4[i].x - gp.x, g4[i].y - gp.y, g4[i].z - gp.z, g4[i].w - gp.w);  // lbfgs.py:404
      s4[i] = make_float4(dd.x * t, dd.y * t, dd.z * t, dd.w * t);                          // lbfgs.py:405
      st4(hblk + (2 * pn) * row + hoff[i], s4[i]);
      st4(hblk + (2 * pn + 1) * row + hoff[i], y4[i]);
    }
    sc[0] = dot4(s4[i], s4[i], sc[0]);
    sc[1] = dot4(s4[i], y4[i], sc[1]);
    sc[2] = dot4(y4[i], y4[i], sc[2]);
    sc[3] = dot4(s4[i], g4[i], sc[3]);
    sc[4] = dot4(y4[i], g4[i], sc[4]);
    sc[5] = dot4(g4[i], g4[i], sc[5]);
    sc[7] += fabsf(g4[i].x) + fabsf(g4[i].y) + fabsf(g4[i].z) + fabsf(g4[i].w);
    gmax = fmaxf(gmax, fmaxf(fmaxf(fabsf(g4[i].x), fabsf(g4[i].y)), fmaxf(fabsf(g4[i].z), fabsf(g4[i].w))));
  }
  gmax = warp_max(gmax);
  warp_reduce8(sc, lane);
  if ((lane & 3) == 0) wscal[warp][lane >> 2] = (lane >> 2) == 6 ? gmax : sc[0];

  __syncthreads();  // the ring's barriers (initialised by thread 0 at kernel entry) are visible to every thread
  for (int i = 0; i < len; ++i) {
    int p = head + i;
    if (p >= NST_LBFGS_SLOTS) p -= NST_LBFGS_SLOTS;
    const int st = i % LB_RING;
    const uint32_t ph = static_cast<uint32_t>(i / LB_RING) & 1u;
    if (threadIdx.x == 0 && i >= 1 && i - 1 + LB_RING < len) {
      mbar_wait(&empty_bar[(i - 1) % LB_RING], static_cast<uint32_t>((i - 1) / LB_RING) & 1u);
      issue(i - 1 + LB_RING);
    }
    mbar_wait(&full_bar[st], ph);
    const float* Sp = reinterpret_cast<const float*>(ring + st * stage_bytes);
    const float* Yp = Sp + row;
    // S_p.y  S_p.g  Y_p.y  Y_p.g : all the recursion needs (lbfgs_ctl.h)
    float r[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < LB_VEC_PER_THREAD; ++k) {
      if (!ok[k]) continue;
      const float4 a4 = ld4(Sp + hoff[k]);
      const float4 c4 = ld4(Yp + hoff[k]);
      r[0] = dot4(a4, y4[k], r[0]);
      r[1] = dot4(a4, g4[k], r[1]);
      r[2] = dot4(c4, y4[k], r[2]);
      r[3] = dot4(c4, g4[k], r[3]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[st]);
    warp_reduce4(r, lane);
    if ((lane & 7) == 0) wpart[warp][p][lane >> 3] = r[0];
  }
  __syncthreads();
  // cross-warp sums in a fixed order -> per-block partials
  // per-block partials, output-major: part[o][block] - the controller sums one output with coalesced loads
  float* out = b.part + blockIdx.x;
  const size_t ostride = static_cast<size_t>(b.nblocks);
  for (int idx = threadIdx.x; idx < len * NST_LBFGS_NDOT; idx += LB_THREADS) {
    const int i = idx / NST_LBFGS_NDOT, q = idx - NST_LBFGS_NDOT * i;
    int p = head + i;
    if (p >= NST_LBFGS_SLOTS) p -= NST_LBFGS_SLOTS;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < LB_THREADS / 32; ++w) acc += wpart[w][p][q];
    out[static_cast<size_t>(p * NST_LBFGS_NDOT + q) * ostride] = acc;
  }
  if (threadIdx.x < NST_LBFGS_NSCAL) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < LB_THREADS / 32; ++w)
      acc = threadIdx.x == 6 ? fmaxf(acc, wscal[w][threadIdx.x]) : acc + wscal[w][threadIdx.x];
    out[static_cast<size_t>(NST_LBFGS_SLOTS * NST_LBFGS_NDOT + threadIdx.x) * ostride] = acc;
  }
}

// One warp per output: sums the per-block partials of pass 1 in fp64 in a fixed order (lane-strided sums, then a shuffle tree).
// A kernel of its own on purpose: every load of a warp is independent, so the whole reduction is ONE round trip to L2 spread over
// ~50 SMs (6 us).  Folded into the one-block controller the same 0.5 MB take tens of microseconds: right after the streaming
// pass a load batch comes back after ~2 us, and one block needs many dependent batches (measured: 49 - 88 us for three
// variants of the folded form, profiles/r02_lbfgs_controller_experiments.md).
__global__ void __launch_bounds__(256) lbfgs_pass1_reduce_kernel(const LbfgsBuffers b) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const NstLbfgsCtl* ctl = b.ctl;
  if (ctl->stop != NST_RUN) return;
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= LB_PART_STRIDE) return;
  const bool is_scal = o >= NST_LBFGS_SLOTS * NST_LBFGS_NDOT;
  if (!is_scal) {
    // skip slots that hold no stored pair
    int rel = o / NST_LBFGS_NDOT - ctl->hist_head;
    if (rel < 0) rel += NST_LBFGS_SLOTS;
    if (rel >= ctl->hist_len) return;
  }
  const bool is_max = o == NST_LBFGS_SLOTS * NST_LBFGS_NDOT + 6;
  const float* src = b.part + static_cast<size_t>(o) * b.nblocks;   // output-major: the lanes read consecutive floats
  double acc = 0.0;
  for (int blk = lane; blk < b.nblocks; blk += 32) {
    const double val = static_cast<double>(src[blk]);
    acc = is_max ? fmax(acc, val) : acc + val;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, acc, s);
    acc = is_max ? fmax(acc, other) : acc + other;
  }
  if (lane == 0) {
    if (is_scal) b.scal[o - NST_LBFGS_SLOTS * NST_LBFGS_NDOT] = acc;
    else b.dots[o] = acc;
  }
}

__global__ void __launch_bounds__(LB_CTL_THREADS) lbfgs_control_kernel(const LbfgsBuffers b, int mode) {
  extern __shared__ double ctl_smem[];
  NstCtlWork w;
  w.R = ctl_smem;
  w.YY = w.R + NST_CTL_MAT_DOUBLES;
  w.Sg = w.YY + NST_CTL_MAT_DOUBLES;
  w.Yg = w.Sg + NST_LBFGS_SLOTS;
  w.al = w.Yg + NST_LBFGS_SLOTS;
  w.c = w.al + NST_LBFGS_SLOTS;
  w.yq = w.c + NST_LBFGS_SLOTS;
  w.ro = w.yq + NST_LBFGS_SLOTS;
  w.red = w.ro + NST_LBFGS_SLOTS;
  uint64_t* bar = reinterpret_cast<uint64_t*>(w.red + 4);
#ifdef NST_INSTRUMENT
  if (threadIdx.x == 0) b.ctl->clk[0] = clock64();
#endif
  const bool stage = b.ctl->stop == NST_RUN && b.ctl->n_iter > 0;
  if (stage && threadIdx.x == 0) {
    // stage the persistent dot-product matrices in shared memory with two bulk copies (2 x 80 KB, L2 resident):
    // the recurrences below touch every element once, on a dependent chain - it must not see global-memory latency
    constexpr uint32_t BYTES = NST_CTL_MAT_DOUBLES * 8;
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(bar, 2 * BYTES);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(w.R)),
                 "l"(b.R), "r"(BYTES), "r"(smem_u32(bar))
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(w.YY)),
                 "l"(b.YY), "r"(BYTES), "r"(smem_u32(bar))
                 : "memory");
  }
  // the matrices staged above are written by this kernel only (previous iteration); the dot products come from the
  // reduction kernel: wait for it here, after the 160 KB copy has been issued
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __syncthreads();
  if (stage) mbar_wait(bar, 0);
  __syncthreads();
  nst_lbfgs_control(b.ctl, w, b.R, b.YY, b.dots, b.scal, *b.eval_loss, b.td_part, b.nblocks, mode);
}

// ------------------------------------------------------------------------------------------------
// pass 2
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LB_THREADS) lbfgs_pass2_kernel(const LbfgsBuffers b) {
  __shared__ float coef[NST_LBFGS_NB + 3];
  __shared__ float wmax[LB_THREADS / 32];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const NstLbfgsCtl* ctl = b.ctl;
  if (ctl->run_pass2 == 0) return;
  const int len = ctl->hist_len, head = ctl->hist_head;
  const float t = ctl->t_apply;
  for (int k = threadIdx.x; k < NST_LBFGS_NB; k += LB_THREADS) coef[k] = ctl->coef[k];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  const int nv = b.n_pad >> 2;
  const int v0 = blockIdx.x * b.vec_per_blk;
  const int v1 = min(nv, v0 + b.vec_per_blk);
  const size_t row = static_cast<size_t>(b.vec_per_blk) * 4;
  const float* hblk = b.hist + static_cast<size_t>(blockIdx.x) * (2 * NST_LBFGS_SLOTS) * row;

  float4 acc[LB_VEC_PER_THREAD];
  size_t off[LB_VEC_PER_THREAD];
  int hoff[LB_VEC_PER_THREAD];
  bool ok[LB_VEC_PER_THREAD];
  const float cg = coef[NST_LBFGS_G];
#pragma unroll
  for (int i = 0; i < LB_VEC_PER_THREAD; ++i) {
    const int lv = i * LB_THREADS + threadIdx.x;
    const int vi = v0 + lv;
    ok[i] = vi < v1;
    off[i] = static_cast<size_t>(ok[i] ? vi : v0) * 4;
    hoff[i] = (ok[i] ? lv : 0) * 4;
    acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok[i]) {
      const float4 g = ld4(b.g + off[i]);
      st4(b.g_prev + off[i], g);  // lbfgs.py:444-447
      acc[i] = make_float4(cg * g.x, cg * g.y, cg * g.z, cg * g.w);
    }
  }
  // The stored pairs stream through a ring of LB_RING shared-memory stages filled by bulk copies (cp.async.bulk, one
  // 2 x row copy per pair: S_p and Y_p are adjacent in the chunk-major history): four pairs (~85 KB) per block are in
  // flight instead of the two a register pipeline can hold, which is what a pure read stream needs to saturate HBM.
  extern __shared__ __align__(128) uint8_t lb_smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(lb_smem);
  uint64_t* empty_bar = full_bar + LB_RING;
  uint8_t* ring = lb_smem + 128;
  const uint32_t pair_bytes = static_cast<uint32_t>(row) * 8u;  // S row + Y row
  const uint32_t stage_bytes = (pair_bytes + 127u) & ~127u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < LB_RING; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], LB_THREADS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  // Visiting order: NEWEST pair first.  Pass 1 streamed oldest -> newest two small kernels ago, so the newest ~100 MB of the
  // history are still in the 126 MB L2 when this pass starts with them.
  auto issue = [&](int i) {  // thread 0: i-th visited pair (age len-1-i) -> stage i % LB_RING
    int p = head + (len - 1 - i);
    if (p >= NST_LBFGS_SLOTS) p -= NST_LBFGS_SLOTS;
    const int st = i % LB_RING;
    mbar_arrive_expect_tx(&full_bar[st], pair_bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(ring + st * stage_bytes)),
                 "l"(hblk + static_cast<size_t>(2 * p) * row), "r"(pair_bytes), "r"(smem_u32(&full_bar[st]))
                 : "memory");
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < LB_RING && i < len; ++i) issue(i);
  for (int i = 0; i < len; ++i) {
    int p = head + (len - 1 - i);
    if (p >= NST_LBFGS_SLOTS) p -= NST_LBFGS_SLOTS;
    const int st = i % LB_RING;
    const uint32_t ph = static_cast<uint32_t>(i / LB_RING) & 1u;
    // refill the stage consumed one iteration ago (every warp has had a full iteration to release it)
    if (threadIdx.x == 0 && i >= 1 && i - 1 + LB_RING < len) {
      mbar_wait(&empty_bar[(i - 1) % LB_RING], static_cast<uint32_t>((i - 1) / LB_RING) & 1u);
      issue(i - 1 + LB_RING);
    }
    const float cs = coef[p], cy = coef[NST_LBFGS_SLOTS + p];
    mbar_wait(&full_bar[st], ph);
    const float* Sp = reinterpret_cast<const float*>(ring + st * stage_bytes);
    const float* Yp = Sp + row;
#pragma unroll
    for (int k = 0; k < LB_VEC_PER_THREAD; ++k) {
      if (!ok[k]) continue;
      const float4 a4 = ld4(Sp + hoff[k]);
      const float4 c4 = ld4(Yp + hoff[k]);
      acc[k].x = fmaf(cs, a4.x, fmaf(cy, c4.x, acc[k].x));
      acc[k].y = fmaf(cs, a4.y, fmaf(cy, c4.y, acc[k].y));
      acc[k].z = fmaf(cs, a4.z, fmaf(cy, c4.z, acc[k].z));
      acc[k].w = fmaf(cs, a4.w, fmaf(cy, c4.w, acc[k].w));
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[st]);
  }
  float mtd = 0.f;
#pragma unroll
  for (int i = 0; i < LB_VEC_PER_THREAD; ++i) {
    if (!ok[i]) continue;
    st4(b.d + off[i], acc[i]);
    float4 x = ld4(b.x + off[i]);
    // lbfgs.py:492 (_add_grad) followed by the closure's clamp_(0, 1), run_style_transfer.py:108-109
    x.x = fminf(fmaxf(fmaf(t, acc[i].x, x.x), 0.f), 1.f);
    x.y = fminf(fmaxf(fmaf(t, acc[i].y, x.y), 0.f), 1.f);
    x.z = fminf(fmaxf(fmaf(t, acc[i].z, x.z), 0.f), 1.f);
    x.w = fminf(fmaxf(fmaf(t, acc[i].w, x.w), 0.f), 1.f);
    st4(b.x + off[i], x);
    mtd = fmaxf(mtd, fmaxf(fmaxf(fabsf(acc[i].x * t), fabsf(acc[i].y * t)),
                           fmaxf(fabsf(acc[i].z * t), fabsf(acc[i].w * t))));
  }
  mtd = warp_max(mtd);
  if (lane == 0) wmax[warp] = mtd;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
#pragma unroll
    for (int w = 0; w < LB_THREADS / 32; ++w) m = fmaxf(m, wmax[w]);
    b.td_part[blockIdx.x] = m;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
void lbfgs_plan(LbfgsBuffers& b, int num_sms) {
  const int nv = b.n_pad >> 2;
  int r = (nv + LB_MAX_VEC_PER_BLOCK * num_sms - 1) / (LB_MAX_VEC_PER_BLOCK * num_sms);
  if (r < 1) r = 1;
  int nblocks = num_sms * r;
  if (nblocks > nv) nblocks = nv > 0 ? nv : 1;
  b.nblocks = nblocks;
  b.vec_per_blk = (nv + nblocks - 1) / nblocks;
}

// dynamic shared memory of the streaming passes: barriers + LB_RING stages of one (s, y) pair of a block's chunk
size_t lbfgs_ring_bytes(const LbfgsBuffers& b) {
  const size_t pair_bytes = static_cast<size_t>(b.vec_per_blk) * 32;
  return 128 + LB_RING * ((pair_bytes + 127) & ~static_cast<size_t>(127));
}

size_t lbfgs_hist_floats(const LbfgsBuffers& b) {
  return static_cast<size_t>(b.nblocks) * (2 * NST_LBFGS_SLOTS) * static_cast<size_t>(b.vec_per_blk) * 4;
}

cudaError_t launch_lbfgs_step_begin(const LbfgsBuffers& b, cudaStream_t s) {
  lbfgs_step_begin_kernel<<<1, 1, 0, s>>>(b.ctl);
  return cudaGetLastError();
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args... args) {
  static const bool pdl = getenv("NST_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  static const int prio = [] {
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    return getenv("NST_KERNEL_PRIO") != nullptr ? least : 1000;
  }();
  if (prio != 1000) {
    attr[na].id = cudaLaunchAttributePriority;
    attr[na].val.priority = prio;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
cudaError_t launch_lbfgs_pass1(const LbfgsBuffers& b, cudaStream_t s) {
  return launch_pdl(lbfgs_pass1_kernel, b.nblocks, LB_THREADS, lbfgs_ring_bytes(b), s, b);
}
cudaError_t launch_lbfgs_reduce(const LbfgsBuffers& b, cudaStream_t s) {
  const int warps_per_blk = 8;
  return launch_pdl(lbfgs_pass1_reduce_kernel, (LB_PART_STRIDE + warps_per_blk - 1) / warps_per_blk, 32 * warps_per_blk, 0, s, b);
}
cudaError_t lbfgs_init() {
  cudaError_t e = cudaFuncSetAttribute(lbfgs_control_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LB_CTL_SMEM);
  const int ring_max = 128 + LB_RING * LB_MAX_VEC_PER_BLOCK * 32;
  if (e == cudaSuccess) e = cudaFuncSetAttribute(lbfgs_pass2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_max);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(lbfgs_pass1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_max);
  return e;
}
cudaError_t launch_lbfgs_control(const LbfgsBuffers& b, int mode, cudaStream_t s) {
  return launch_pdl(lbfgs_control_kernel, 1, LB_CTL_THREADS, LB_CTL_SMEM, s, b, mode);
}
cudaError_t launch_lbfgs_pass2(const LbfgsBuffers& b, cudaStream_t s) {
  return launch_pdl(lbfgs_pass2_kernel, b.nblocks, LB_THREADS, lbfgs_ring_bytes(b), s, b);
}

cudaError_t launch_lbfgs_iteration(const LbfgsBuffers& b, int mode, cudaStream_t s) {
  cudaError_t e = launch_lbfgs_pass1(b, s);
  if (e == cudaSuccess) e = launch_lbfgs_reduce(b, s);
  if (e == cudaSuccess) e = launch_lbfgs_control(b, mode, s);
  if (e == cudaSuccess) e = launch_lbfgs_pass2(b, s);
  return e;
}

}  // namespace nst
