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

// ---- descriptors (bit layouts: PTX ISA "tcgen05 matrix / instruction descriptor") ----
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row x 128-byte atoms, atoms
// stacked every 1024 bytes (stride byte offset); leading byte offset unused for swizzled
// K-major (encoded 1); descriptor version 1 (Blackwell); layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                            // LBO            [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                    // SBO            [32,46)
  d |= static_cast<uint64_t>(1) << 46;                            // version        [46,48)
  d |= static_cast<uint64_t>(2) << 61;                            // SWIZZLE_128B   [61,64)
  return d;
}
// Instruction descriptor for kind::f16: D = f32, A = B = bf16, both K-major, M x N tile.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) |                      // c_format = F32
         (1u << 7) | (1u << 10) |         // a_format = b_format = BF16
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];" ::"r"(ptx::smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(ptx::smem_u32(bar))
      : "memory");
}

// Same load, multicast to every CTA of the cluster named in `cta_mask`: the tile lands at the same
// shared-memory offset in each destination CTA and each destination's mbarrier (same offset) gets
// the complete_tx — one L2 read feeds both SMs of the pair.
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const CUtensorMap* map, int x, int y,
                                                      uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(ptx::smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(ptx::smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   ptx::smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

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

PFN_cuTensorMapEncodeTiled tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    (void)cudaGetLastError();
  });
  HPCP_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available (driver too old?)");
  return fn;
}

CUtensorMap make_operand_map(const void* base, int rows, int box_rows) {
  CUtensorMap map;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kTileK), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(kTileK) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kTileK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult r = tensor_map_encoder()(
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
