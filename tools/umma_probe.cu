// Probe: how does tcgen05.mma address a 128B-swizzled K-major operand whose start address is NOT 1024-byte aligned,
// and whose 8-row groups are NOT 1024 bytes apart?  (Decides whether a convolution can issue its nine filter taps
// from ONE halo patch in shared memory by shifting the operand descriptor instead of re-loading the patch per tap.)
//
// Shared memory holds X[rows][64] fp16 in the layout TMA writes with CU_TENSOR_MAP_SWIZZLE_128B: row r at byte r*128,
// 16-byte chunk c stored at chunk position c ^ (r & 7).  B is a 64x64 identity (K-major, same swizzle), so
// D[m][n] = A[m][n]: D shows which shared-memory element every (m, k) of the A operand was fetched from.
// Two runs encode the source row and the source column in the values.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu && tools/umma_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../text-based-image-style-transfer_b200/csrc/common.cuh"

using namespace nst;

static constexpr int ROWS = 320;

struct Case {
  int shift_rows;   // start address = base + shift_rows * 128
  int sbo_bytes;    // stride between 8-row groups
  int base_offset;  // descriptor bits [49,52)
};

__global__ void __launch_bounds__(128) probe_kernel(const __half* __restrict__ X /*[ROWS][64] logical*/, Case cs,
                                                    float* __restrict__ D /*[128][64]*/) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sX = smem;                      // ROWS * 128 bytes
  uint8_t* sB = smem + ROWS * 128;         // 64 * 128 bytes (ROWS*128 is a multiple of 1024)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 64 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  // X in TMA SW128 layout
  for (int i = tid; i < ROWS * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(X + r * 64 + c * 8);
    *reinterpret_cast<uint4*>(sX + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  // B = identity [n][k]
  for (int i = tid; i < 64 * 8; i += 128) {
    const int n = i >> 3, c = i & 7;
    __half h[8];
    for (int e = 0; e < 8; ++e) h[e] = __float2half((c * 8 + e) == n ? 1.f : 0.f);
    *reinterpret_cast<uint4*>(sB + n * 128 + ((c ^ (n & 7)) << 4)) = *reinterpret_cast<uint4*>(h);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (tid < 32) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  // make the generic-proxy shared-memory writes visible to the async proxy (tensor core reads)
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_f16(128, 64, 0, 0, 0);
    const uint32_t a_addr = smem_u32(sX) + cs.shift_rows * 128;
    const uint32_t b_addr = smem_u32(sB);
    for (int k = 0; k < 4; ++k) {
      uint64_t da = umma_desc_sw128(a_addr + k * 32, 16, cs.sbo_bytes);
      da |= static_cast<uint64_t>(cs.base_offset & 7) << 49;
      const uint64_t db = umma_desc_sw128(b_addr + k * 32, 16, 1024);
      umma_f16(tmem, da, db, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 64);
}

// ---- issue-rate microbenchmark: back-to-back tcgen05.mma (M = 128, K = 16) from shared-memory operands, no TMA traffic
__global__ void __launch_bounds__(128) rate_kernel(int N, int iters, long long* out, int shift_rows, int sbo, int vary) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sA = smem;               // 192 x 128 B (room for a shifted / strided operand)
  uint8_t* sB = smem + 24576;       // up to 256 x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  for (int i = tid; i < (24576 + 32768) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (tid < 32) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // the issuing thread is chosen with elect.sync inside a warp-uniform branch: the compiler then knows that exactly one
  // lane executes the uniform-datapath UTCHMMA and does not wrap each one in an ELECT / BRA.U.ANY serialisation loop
  if (tid < 32 && elect_one()) {
    const uint32_t idesc = umma_idesc_f16(128, N, 0, 0, 0);
    const uint32_t a_addr0 = smem_u32(sA) + shift_rows * 128, b_addr = smem_u32(sB);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // vary: walk the start row like the nine filter taps do (0,1,2,10,11,12,20,21,22)
      const int tap = vary ? it % 9 : 0;
      const uint32_t a_addr = a_addr0 + ((tap / 3) * 10 + tap % 3) * 128;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = umma_desc_sw128(a_addr + k * 32, 16, sbo);
        const uint64_t db = umma_desc_sw128(b_addr + k * 32, 16, 1024);
        umma_f16(tmem, da, db, idesc, 1u);
      }
    }
    const long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 256);
}

static void rate_bench() {
  long long* d;
  long long h[2];
  cudaMalloc(&d, 2 * sizeof(long long));
  const int smem = 24576 + 32768 + 1024 + 64;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  const int Ns[] = {16, 64, 128, 256};
  for (int N : Ns) {
    const int variants[5][3] = {{0, 1024, 0}, {0, 1280, 0}, {1, 1024, 0}, {11, 1280, 0}, {0, 1280, 1}};
    for (int v = 0; v < 5; ++v) {
      rate_kernel<<<148, 128, smem>>>(N, iters, d, variants[v][0], variants[v][1], variants[v][2]);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("rate N=%d: %s\n", N, cudaGetErrorString(e));
        return;
      }
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("MMA rate M=128 N=%3d K=16 A: start row %2d, group stride %4d B, tap walk %d: issue %.1f, complete %.1f cycles/MMA (tensor floor %d)\n",
             N, variants[v][0], variants[v][1], variants[v][2], (double)h[0] / (4.0 * iters), (double)h[1] / (4.0 * iters), 128 * N / 256);
    }
  }
  cudaFree(d);
}

int main() {
  rate_bench();
  __half* hX = (__half*)malloc(ROWS * 64 * sizeof(__half));
  __half* dX;
  float* dD;
  float hD[2][128 * 64];
  cudaMalloc(&dX, ROWS * 64 * sizeof(__half));
  cudaMalloc(&dD, 128 * 64 * sizeof(float));
  const int smem = ROWS * 128 + 64 * 128 + 1024 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 10, 11, 17, 18, 19, 20, 37};
  const int sbos[] = {1024, 1280, 2304};  // contiguous, 10-row groups (8-wide tile + halo), 18-row groups
  int n_ok = 0, n_all = 0;
  for (int sbo : sbos) {
    for (int sh : shifts) {
      for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
        Case cs{sh, sbo, bo_mode ? (sh & 7) : 0};
        if (bo_mode && cs.base_offset == 0) continue;
        for (int run = 0; run < 2; ++run) {
          for (int r = 0; r < ROWS; ++r)
            for (int k = 0; k < 64; ++k) hX[r * 64 + k] = __float2half(run == 0 ? (float)r : (float)k);
          cudaMemcpy(dX, hX, ROWS * 64 * sizeof(__half), cudaMemcpyHostToDevice);
          cudaMemset(dD, 0xff, 128 * 64 * sizeof(float));
          probe_kernel<<<1, 128, smem>>>(dX, cs, dD);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("shift %d sbo %d bo %d: CUDA error %s\n", sh, sbo, cs.base_offset, cudaGetErrorString(e));
            return 1;
          }
          cudaMemcpy(hD[run], dD, sizeof(hD[run]), cudaMemcpyDeviceToHost);
        }
        // expected under the "absolute-address swizzle" model: row m of the operand = X row group(m) + m%8 + shift
        int bad = 0, first_bad = -1;
        for (int m = 0; m < 128; ++m) {
          const int want_row = (m / 8) * (sbo / 128) + (m % 8) + sh;
          for (int n = 0; n < 64; ++n) {
            if ((int)hD[0][m * 64 + n] != want_row || (int)hD[1][m * 64 + n] != n) {
              ++bad;
              if (first_bad < 0) first_bad = m * 64 + n;
            }
          }
        }
        ++n_all;
        if (bad == 0) {
          ++n_ok;
          printf("shift %2d sbo %4d base_offset %d : OK (operand row m = X[(m/8)*%d + m%%8 + %d])\n", sh, sbo, cs.base_offset,
                 sbo / 128, sh);
        } else {
          printf("shift %2d sbo %4d base_offset %d : MISMATCH %d/8192; rows 0..11 read (row,col0..3): ", sh, sbo, cs.base_offset, bad);
          for (int m = 0; m < 12; ++m)
            printf("[%d: r%d c%d,%d,%d,%d] ", m, (int)hD[0][m * 64], (int)hD[1][m * 64], (int)hD[1][m * 64 + 8],
                   (int)hD[1][m * 64 + 16], (int)hD[1][m * 64 + 24]);
          printf("\n");
        }
      }
    }
  }
  printf("%d of %d descriptor variants follow the absolute-address swizzle model\n", n_ok, n_all);
  return 0;
}
