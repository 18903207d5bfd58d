// K-gemm-put: bf16 GEMM on the 5th-gen tensor cores whose epilogue PUTs the result tile into a
// peer GPU over NVLink — "compute step followed by a transfer" as ONE kernel.
//
//     C[M,N] (fp32) = A[M,K] (bf16, K-major) . B[N,K]^T (bf16, K-major)
//     C is written to c_local and/or to c_peer (a peer-mapped pointer), then the arrival epoch is
//     published on the peer (same release/acquire protocol as every other kernel of the suite).
//
// Nothing in the reference is GEMM-shaped; this kernel is the tensor-core member of
// the suite's "fuse the transfer into the producing kernel" family, next to K-fused-triad-put
// (stream triad -> put) and K-ring (accumulate -> forward).  Stock comparison: cuBLAS GEMM followed
// by cudaMemcpyPeerAsync (two operations, the C tile goes to HBM and is read back before it moves).
//
// Structure (persistent, one CTA per SM, 256 threads, canonical Blackwell warp specialisation):
//   warp 0      TMA producer  : cp.async.bulk.tensor.2d (SWIZZLE_128B) A[128x64] + B[256x64] per stage,
//                               4-stage smem ring, full/empty mbarriers
//   warp 1      MMA issuer    : one elected thread, tcgen05.mma.cta_group::1.kind::f16 (M128 N256 K16),
//                               tcgen05.commit frees the smem stage / publishes the accumulator
//   warp 2      TMEM allocator: 512 columns = two 128x256 fp32 accumulators (epilogue of tile i overlaps
//                               the main loop of tile i+1)
//   warps 4..7  epilogue      : tcgen05.ld 32x32b.x32 -> padded smem transpose -> 128-byte row segments
//                               -> st.global to c_local and/or the peer (coalesced NVLink stores)
#include "api.h"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "../common/cuda_check.h"
#include "../common/signal.cuh"
#include "umma.cuh"

namespace hpcp {

namespace {

using namespace umma;  // tile geometry, descriptors, TMA / tcgen05 wrappers, rasterisation, tensor-map encoder

constexpr int kStages = 4;
constexpr uint32_t kStageBytes = kABytes + kBBytes;     // 48 KiB
constexpr size_t kSmemBytes = static_cast<size_t>(kStages) * kStageBytes + kEpiWarps * kEpiWarpBytes + 1024;

struct GemmDev {
  void* c_local;    // may be null; fp32 [M,N], or bf16 [M,N] when out_bf16
  void* c_peer;     // may be null
  int out_bf16;
  int m, n, k;
  int tiles_m, tiles_n;
  SyncOps sync;
};

// The kernel is the suite's persistent tile loop (umma.cuh: gemm_persistent) with this policy: grouped
// rasterisation over the whole C, stores (local and/or peer) in the epilogue, one arrival epoch at the end.
struct PutPolicy {
  static constexpr bool kHasAuxWarp = false;
  const GemmDev& g;
  __device__ __forceinline__ void coords(int tile, int* m_blk, int* n_blk) const {
    tile_coords(tile, g.tiles_m, g.tiles_n, m_blk, n_blk);
  }
  __device__ __forceinline__ void a_rows_ready(int) const {}
  __device__ __forceinline__ void epilogue(uint32_t taddr, float* stage_buf, int m0, int n0, int ew, int lane) const {
    epilogue_store_tile(g, taddr, stage_buf, m0, n0, ew, lane);
  }
  __device__ __forceinline__ void aux_warp(int, unsigned char*) const {}
  // Put epilogue: the last CTA publishes the arrival epoch on the peer.
  __device__ __forceinline__ void finish() const {
    if (g.sync.ticket != nullptr)
      last_cta_publish(g.sync.ticket, g.sync.ticket_base + gridDim.x, g.sync.signal_flag, g.sync.signal_epoch);
  }
};

template <int kCluster>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_put_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ GemmDev g) {
  gemm_persistent<kCluster, kStages>(map_a, map_b, g.tiles_m, g.tiles_n, g.k, PutPolicy{g});
}

// ---------------------------------------------------------------- 2-SM UMMA variant ----
// The CTA pair of a cluster computes ONE 256x256 tile with tcgen05.mma.cta_group::2 (M=256, N=256, K=16):
// CTA r owns rows [128r, 128r+128) of the tile (its A rows, its TMEM accumulator, its epilogue) and stages
// only HALF of the B tile, so a pipeline stage is 32 KiB per CTA instead of 48 (6 stages instead of 4) and
// each SM reads 8 instead of 12 KiB of operands from its shared memory per MMA.  Only the leader (rank 0)
// issues MMAs and owns the pair-wide barriers:
//   full[s]       (leader)  2 arrivals: leader's expect_tx(2 x 32 KiB) + the peer's remote arrive; the TMA
//                           loads of BOTH CTAs complete_tx on it
//   empty[s]      (each)    1 arrival: tcgen05.commit multicast to the pair
//   tmem_full[a]  (each)    1 arrival: tcgen05.commit multicast to the pair
//   tmem_empty[a] (leader)  2 x 4 arrivals: the epilogue warps of both CTAs
constexpr int kStages2 = 6;
constexpr uint32_t kStageBytes2 = kABytes + kBBytes / 2;  // 32 KiB per CTA
constexpr size_t kSmemBytes2 = static_cast<size_t>(kStages2) * kStageBytes2 + kEpiWarps * kEpiWarpBytes + 1024;

__global__ void __launch_bounds__(kThreads, 1)
    gemm_put_2sm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const __grid_constant__ GemmDev g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages2];
  __shared__ __align__(8) uint64_t empty_bar[kStages2];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;

  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* epi_smem = smem + static_cast<size_t>(kStages2) * kStageBytes2;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = g.k / kBK;
  const int crank = static_cast<int>(cluster_cta_rank());
  const bool leader = crank == 0;
  // Work items are 256x256 tiles = vertically adjacent pairs (2p, 2p+1) of the grouped rasterisation.
  const int first_item = static_cast<int>(blockIdx.x) / 2;
  const int item_stride = static_cast<int>(gridDim.x) / 2;
  const int num_items = g.tiles_m * g.tiles_n / 2;

  cluster_sync_all();  // both CTAs of the pair are resident before the pair-wide TMEM allocation
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < kStages2; ++s) {
      ptx::mbar_init(&full_bar[s], 2);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], 2 * kEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {  // the same warp of both CTAs allocates collectively
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     ptx::smem_u32(&tmem_base_s)),
                 "r"(static_cast<uint32_t>(kTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the partner's barriers and TMEM exist before anything remote touches them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int item = first_item; item < num_items; item += item_stride) {
      int m_blk, n_blk;
      tile_coords(item * 2 + crank, g.tiles_m, g.tiles_n, &m_blk, &n_blk);
      const int m0 = m_blk * kBM;
      const int n0 = n_blk * kBN + crank * (kBN / 2);  // my half of the shared B tile
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_or_trap(&empty_bar[stage], phase ^ 1, false, kWaitEmpty, stage);  // slot released by the pair's MMA
        if (lane == 0) {
          unsigned char* sa = smem + static_cast<size_t>(stage) * kStageBytes2;
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes2);
          tma_load_2d_2sm(sa, &map_a, kb * kBK, m0, &full_bar[stage]);
          tma_load_2d_2sm(sa + kABytes, &map_b, kb * kBK, n0, &full_bar[stage]);
          if (!leader) mbar_arrive_cluster(&full_bar[stage], 0);
        }
        __syncwarp();
        if (++stage == kStages2) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader CTA only) =====================
    const uint32_t idesc = make_idesc(2 * kBM, kBN);
    int stage = 0;
    uint32_t phase = 0;
    int local_tile = 0;
    for (int item = first_item; item < num_items; item += item_stride, ++local_tile) {
      const int acc = local_tile & 1;
      const uint32_t acc_phase = static_cast<uint32_t>(local_tile >> 1) & 1;
      mbar_wait_or_trap(&tmem_empty_bar[acc], acc_phase ^ 1, true, kWaitTmemEmpty, acc);  // both epilogues drained it
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kBN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_or_trap(&full_bar[stage], phase, true, kWaitFull, stage);  // both CTAs' TMA bytes landed
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(smem + static_cast<size_t>(stage) * kStageBytes2);
          const uint64_t desc_a = make_smem_desc(sa);
          const uint64_t desc_b = make_smem_desc(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint64_t adv = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
            umma_bf16_2sm(tmem_d, desc_a + adv, desc_b + adv, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage], static_cast<uint16_t>(0x3));  // frees the slot in both CTAs
          if (kb == num_kb - 1) umma_commit_2sm(&tmem_full_bar[acc], static_cast<uint16_t>(0x3));
        }
        __syncwarp();
        if (++stage == kStages2) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int ew = warp - 4;
    float* stage_buf = reinterpret_cast<float*>(epi_smem + static_cast<size_t>(ew) * kEpiWarpBytes);
    int local_tile = 0;
    for (int item = first_item; item < num_items; item += item_stride, ++local_tile) {
      const int acc = local_tile & 1;
      const uint32_t acc_phase = static_cast<uint32_t>(local_tile >> 1) & 1;
      int m_blk, n_blk;
      tile_coords(item * 2 + crank, g.tiles_m, g.tiles_n, &m_blk, &n_blk);
      mbar_wait_or_trap(&tmem_full_bar[acc], acc_phase, false, kWaitTmemFull, acc);
      tc_fence_after();
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * kBN) + (static_cast<uint32_t>(ew * 32) << 16);
      epilogue_store_tile(g, taddr, stage_buf, m_blk * kBM, n_blk * kBN, ew, lane);
      tc_fence_before();
      if (lane == 0) {  // the leader's MMA warp may overwrite this accumulator in both CTAs
        if (leader)
          ptx::mbar_arrive(&tmem_empty_bar[acc]);
        else
          mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may free TMEM or retire while its partner can still touch it
  if (warp == 2)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(kTmemCols))
                 : "memory");
  if (g.sync.ticket != nullptr)
    last_cta_publish(g.sync.ticket, g.sync.ticket_base + gridDim.x, g.sync.signal_flag, g.sync.signal_epoch);
}

}  // namespace

int launch_gemm_put(const void* a_bf16, const void* b_bf16, void* c_local, void* c_peer, int m, int n,
                    int k, bool out_bf16, const SyncOps& sync, int ctas, int device, cudaStream_t stream,
                    int cluster) {
  HPCP_REQUIRE(m > 0 && n > 0 && k > 0 && m % kBM == 0 && n % kBN == 0 && k % kBK == 0,
               "gemm_put: M, N, K must be multiples of 128, 256, 64");
  HPCP_REQUIRE(c_local != nullptr || c_peer != nullptr, "gemm_put: no output");
  HPCP_REQUIRE((reinterpret_cast<uintptr_t>(a_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(b_bf16) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(c_local) & 15) == 0 && (reinterpret_cast<uintptr_t>(c_peer) & 15) == 0,
               "gemm_put: pointers must be 16-byte aligned");
  HPCP_REQUIRE(sync.signal_flag == nullptr || sync.ticket != nullptr, "gemm_put: a signal needs a ticket counter");
  // Cluster mode needs an even number of tile rows per raster group (pairs share n_blk).
  HPCP_REQUIRE(cluster >= 0 && cluster <= 3, "gemm_put: cluster must be 0 (auto), 1, 2 or 3 (2-SM UMMA)");
  const bool pairable = (m / kBM) % 2 == 0 && ((m / kBM) % kGroupM) % 2 == 0;
  HPCP_REQUIRE(cluster < 2 || pairable, "gemm_put: cluster=2/3 needs an even number of 128-row tiles per raster group");
  const bool use_cluster = cluster != 1 && pairable;
  const bool two_sm = cluster == 3;
  const CUtensorMap map_a = make_kmajor_map(a_bf16, m, k, kBM);
  const CUtensorMap map_b = make_kmajor_map(b_bf16, n, k, use_cluster ? kBN / 2 : kBN);
  GemmDev g{};
  g.c_local = c_local;
  g.c_peer = c_peer;
  g.out_bf16 = out_bf16 ? 1 : 0;
  g.m = m;
  g.n = n;
  g.k = k;
  g.tiles_m = m / kBM;
  g.tiles_n = n / kBN;
  g.sync = sync;
  const int tiles = g.tiles_m * g.tiles_n;
  const int sms = device_sm_count(device);
  int grid = std::min(tiles, ctas > 0 ? ctas : sms);
  if (!use_cluster || grid < 2) {
    const CUtensorMap map_b_full = use_cluster ? make_kmajor_map(b_bf16, n, k, kBN) : map_b;
    HPCP_ENABLE_SMEM(gemm_put_kernel<1>, kSmemBytes);
    gemm_put_kernel<1><<<grid, kThreads, kSmemBytes, stream>>>(map_a, map_b_full, g);
    HPCP_CUDA(cudaGetLastError());
    return grid;
  }
  HPCP_REQUIRE(!two_sm || grid >= 2, "gemm_put: cluster=3 needs at least two CTAs");
  grid &= ~1;  // whole pairs
  if (two_sm)
    HPCP_ENABLE_SMEM(gemm_put_2sm_kernel, kSmemBytes2);
  else
    HPCP_ENABLE_SMEM(gemm_put_kernel<2>, kSmemBytes);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = two_sm ? kSmemBytes2 : kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (two_sm)
    HPCP_CUDA(cudaLaunchKernelEx(&cfg, gemm_put_2sm_kernel, map_a, map_b, g));
  else
    HPCP_CUDA(cudaLaunchKernelEx(&cfg, gemm_put_kernel<2>, map_a, map_b, g));
  return grid;
}

}  // namespace hpcp
