// K-tc-busy: the tensor-core compute command `T` of the concurrency benchmark.
//
// The reference's compute command is a dependent-FMA chain on the general-purpose
// ALUs (concurency/bench.hpp:23-31).  On B200 the machine's compute capacity is the
// 5th-generation tensor cores, so the suite also offers a compute command that keeps
// the *tensor pipe* busy while copies run next to it: a per-CTA bf16 GEMM tile
//        D[128 x 256] (fp32, TMEM) += A[128 x 64] . B[256 x 64]^T        x tripcount
// * A and B tiles are fetched ONCE by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B)
//   into shared memory, completion on an mbarrier;
// * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (4 x K=16 per pass over
//   the 64-deep tile, `tripcount` passes), accumulating in tensor memory;
// * tcgen05.commit signals an mbarrier; four epilogue warps read the accumulator back
//   with tcgen05.ld (32x32b.x32) and store it, so the result is checkable:
//   D = tripcount * (A . B^T), exact in fp32 for the small-integer operands used.
// The loop touches no global memory, so its duration is proportional to tripcount —
// the same property the FMA chain has — but the busy unit is the tensor core.
//
// Warp roles (6 warps): 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator +
// epilogue, 3..5 = epilogue (warp w reads TMEM lanes 32*(w%4) .. +31).
#include "api.h"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <mutex>

#include "../common/cuda_check.h"
#include "../common/ptx.cuh"
#include "umma.cuh"

namespace hpcp {

namespace {

constexpr int kTileM = 128;
constexpr int kTileN = 256;
constexpr int kTileK = 64;   // one 128-byte swizzle atom of bf16
constexpr int kUmmaK = 16;   // K per tcgen05.mma for 16-bit operands
constexpr int kTmemCols = 256;
constexpr int kThreads = 192;
constexpr uint32_t kABytes = kTileM * kTileK * 2;  // 16 KiB
constexpr uint32_t kBBytes = kTileN * kTileK * 2;  // 32 KiB

// Descriptors, TMA tensor loads and tcgen05 wrappers: umma.cuh (shared with the GEMM kernels).
using umma::cluster_cta_rank;
using umma::cluster_sync_all;
using umma::make_idesc;
using umma::make_smem_desc;
using umma::tma_load_2d;
using umma::tma_load_2d_multicast;
using umma::tmem_ld_32x32b_x32;
using umma::umma_bf16;
using umma::umma_commit;

// Dynamic smem (1024-aligned): A tile | B tile ; static smem: barriers + TMEM base.
// kCluster == 2: launched as thread-block clusters of two CTAs that share the B tile — each CTA
// fetches half of B (128 of its 256 rows) and TMA-multicasts it into both CTAs' shared memory.
template <int kCluster>
__global__ void __launch_bounds__(kThreads)
    tc_busy_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   float* __restrict__ out, uint32_t tripcount) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar;
  __shared__ __align__(8) uint64_t mma_done_bar;
  __shared__ uint32_t tmem_base_s;

  // SWIZZLE_128B operand tiles must start on a 1024-byte boundary of the shared window.
  unsigned char* smem_a = smem + ((1024u - (ptx::smem_u32(smem) & 1023u)) & 1023u);
  unsigned char* smem_b = smem_a + kABytes;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    ptx::mbar_init(&full_bar, 1);
    ptx::mbar_init(&mma_done_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {  // one warp allocates (and later frees) the accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     ptx::smem_u32(&tmem_base_s)),
                 "r"(static_cast<uint32_t>(kTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  // The peer CTA's mbarrier must be initialised before anything is multicast into it.
  if (kCluster > 1) cluster_sync_all();

  if (warp == 0) {
    if (lane == 0) {  // TMA producer: both operand tiles, once
      ptx::mbar_arrive_expect_tx(&full_bar, kABytes + kBBytes);
      tma_load_2d(smem_a, &map_a, 0, 0, &full_bar);
      if (kCluster > 1) {
        const uint32_t r = cluster_cta_rank();  // my half of B, delivered to both CTAs
        tma_load_2d_multicast(smem_b + r * (kBBytes / 2), &map_b, 0, static_cast<int>(r) * (kTileN / 2),
                              &full_bar, static_cast<uint16_t>(0x3));
      } else {
        tma_load_2d(smem_b, &map_b, 0, 0, &full_bar);
      }
    }
  } else if (warp == 1) {
    ptx::mbar_wait(&full_bar, 0);  // operands landed in smem
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (lane == 0) {  // single-thread MMA issue
      const uint32_t idesc = make_idesc(kTileM, kTileN);
      const uint64_t desc_a0 = make_smem_desc(ptx::smem_u32(smem_a));
      const uint64_t desc_b0 = make_smem_desc(ptx::smem_u32(smem_b));
      for (uint32_t it = 0; it < tripcount; ++it) {
#pragma unroll
        for (int k = 0; k < kTileK / kUmmaK; ++k) {
          // advance the start address by k * 32 bytes inside the 128-byte swizzle atom
          const uint64_t adv = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
          umma_bf16(tmem_d, desc_a0 + adv, desc_b0 + adv, idesc, (it | static_cast<uint32_t>(k)) != 0u);
        }
      }
      umma_commit(&mma_done_bar);  // arrives when every MMA above has completed
    }
    __syncwarp();
  }
  if (warp >= 2) {  // epilogue: TMEM -> registers -> global
    ptx::mbar_wait(&mma_done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int lane_group = warp & 3;               // TMEM lanes this warp may touch
    const int row = lane_group * 32 + lane;        // accumulator row == TMEM lane
    float* out_row = out + (static_cast<size_t>(blockIdx.x) * kTileM + row) * kTileN;
    for (int col = 0; col < kTileN; col += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_d + (static_cast<uint32_t>(lane_group * 32) << 16) + col, r);
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(out_row + col + j) =
            make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                        __uint_as_float(r[j + 3]));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d),
                 "r"(static_cast<uint32_t>(kTmemCols))
                 : "memory");
  // Neither CTA of a pair may retire while the other could still be receiving its multicast.
  if (kCluster > 1) cluster_sync_all();
}

__global__ void tc_fill_operands_kernel(__nv_bfloat16* a, __nv_bfloat16* b) {
  // Small integers: every product and every partial sum is exact in bf16 / fp32.
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kTileN * kTileK; i += gridDim.x * blockDim.x) {
    const int r = i / kTileK, k = i % kTileK;
    if (r < kTileM) a[i] = __float2bfloat16(static_cast<float>((r + k) % 3 - 1));
    b[i] = __float2bfloat16(static_cast<float>((r * 2 + k) % 5 - 2));
  }
}

CUtensorMap make_operand_map(const void* base, int rows, int box_rows) {
  CUtensorMap map;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kTileK), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(kTileK) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kTileK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult r = umma::gemm_tensor_map_encoder()(
      &map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
      elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HPCP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string(r));
  return map;
}

}  // namespace

size_t tc_busy_operand_bytes() { return static_cast<size_t>(kTileM + kTileN) * kTileK * 2; }
size_t tc_busy_out_elems_per_cta() { return static_cast<size_t>(kTileM) * kTileN; }

void launch_tc_fill_operands(void* operands, cudaStream_t stream) {
  __nv_bfloat16* a = static_cast<__nv_bfloat16*>(operands);
  tc_fill_operands_kernel<<<32, 256, 0, stream>>>(a, a + kTileM * kTileK);
  HPCP_CUDA(cudaGetLastError());
}

void launch_tc_busy(const void* operands, float* out, int ctas, uint32_t tripcount,
                    cudaStream_t stream, int cluster) {
  HPCP_REQUIRE(ctas >= 1, "tc_busy: need at least one CTA");
  HPCP_REQUIRE((reinterpret_cast<uintptr_t>(operands) & 127) == 0, "tc_busy: operands must be 128-byte aligned");
  HPCP_REQUIRE(cluster == 1 || cluster == 2, "tc_busy: cluster size must be 1 or 2");
  if (cluster == 2 && ctas % 2 != 0) cluster = 1;  // pairs only
  const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(operands);
  const CUtensorMap map_a = make_operand_map(a, kTileM, kTileM);
  const size_t smem = kABytes + kBBytes + 1024;  // slack for the 1024-byte alignment
  if (cluster == 1) {
    const CUtensorMap map_b = make_operand_map(a + kTileM * kTileK, kTileN, kTileN);
    HPCP_ENABLE_SMEM(tc_busy_kernel<1>, smem);
    tc_busy_kernel<1><<<ctas, kThreads, smem, stream>>>(map_a, map_b, out, tripcount);
    HPCP_CUDA(cudaGetLastError());
    return;
  }
  // Thread-block clusters of 2: each CTA loads and multicasts one half (128 rows) of B.
  const CUtensorMap map_b = make_operand_map(a + kTileM * kTileK, kTileN, kTileN / 2);
  HPCP_ENABLE_SMEM(tc_busy_kernel<2>, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(ctas));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  HPCP_CUDA(cudaLaunchKernelEx(&cfg, tc_busy_kernel<2>, map_a, map_b, out, tripcount));
}

}  // namespace hpcp
