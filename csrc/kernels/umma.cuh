// UMMA building blocks shared by the tensor-core kernels of the suite (gemm_put.cu, gemm_collective.cu):
// tile geometry, shared-memory / instruction descriptors, TMA tensor loads, tcgen05.mma / commit / ld wrappers
// (cta_group::1 and cta_group::2 forms), the grouped tile rasterisation and the host-side tensor-map encoder.
// Everything device-side is __forceinline__, so including this header does not change a kernel's SASS.
#pragma once

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <mutex>
#include <string>

#include "../common/cuda_check.h"
#include "../common/ptx.cuh"
#include "tile_order.h"

namespace hpcp {
namespace umma {

constexpr int kBM = 128, kBN = 256, kBK = 64, kUmmaK = 16;
constexpr int kThreads = 256;
constexpr int kEpiWarps = 4;
constexpr int kTmemCols = 512;                          // 2 accumulators x 256 columns
constexpr uint32_t kABytes = kBM * kBK * 2;             // 16 KiB
constexpr uint32_t kBBytes = kBN * kBK * 2;             // 32 KiB
constexpr int kStageRowWords = 36;                      // 32 payload floats + 4 pad (keeps float4 alignment)
constexpr uint32_t kEpiWarpBytes = 32 * kStageRowWords * 4;  // 4608 B per epilogue warp

// ---- descriptors (bit layouts: PTX ISA "tcgen05 matrix / instruction descriptor") ----
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row x 128-byte atoms, atoms stacked every 1024 bytes
// (stride byte offset); leading byte offset unused for swizzled K-major (encoded 1); descriptor version 1
// (Blackwell); layout type 2 = SWIZZLE_128B.
__host__ __device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);  // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                      // LBO            [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // SBO            [32,46): 8 rows x 128 B
  d |= static_cast<uint64_t>(1) << 46;                      // version        [46,48)
  d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B   [61,64)
  return d;
}
// Instruction descriptor for kind::f16: D = f32 (bit 4), A = B = bf16 (bits 7, 10), both K-major, N >> 3 at
// [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);  // D=f32, A=B=bf16, K-major both
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];" ::"r"(ptx::smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(ptx::smem_u32(bar))
      : "memory");
}
// Multicast variant: the box lands at the same smem offset in every CTA of `cta_mask`, and each of
// those CTAs' mbarrier (same offset) receives the complete_tx.
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const CUtensorMap* map, int x, int y,
                                                      uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(ptx::smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(ptx::smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// tcgen05.commit that arrives on the barrier at this offset in every CTA of `cta_mask`.
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          ptx::smem_u32(bar)),
      "h"(cta_mask)
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
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   ptx::smem_u32(bar))
               : "memory");
}
// ---- cta_group::2 ("2-SM UMMA") forms: a CTA pair executes ONE M=256 MMA; each CTA holds its own 128 rows
// of A and HALF of the B tile, the tensor cores of the two SMs exchange the B halves themselves.
// Issued by both CTAs; the transaction bytes are counted on the LEADER's barrier (same offset in CTA 0 of
// the pair, address obtained with mapa).
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, int x, int y,
                                                uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b32 lb;\n\t"
      "mapa.shared::cluster.u32 lb, %4, 0;\n\t"
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [lb];\n\t}" ::"r"(ptx::smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(ptx::smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          ptx::smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// Arrive on the barrier at the same offset in CTA `cta` of the cluster.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(ptx::smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// The 2-SM kernel has not run on hardware yet: its waits are bounded so that a protocol error reports WHICH
// barrier never completed (printf + trap -> the launch fails with an error) instead of hanging the GPU.
enum { kWaitEmpty = 0, kWaitFull = 1, kWaitTmemEmpty = 2, kWaitTmemFull = 3 };
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity, bool cluster_scope, int what,
                                                  int index) {
  for (unsigned long long spins = 0;; ++spins) {
    uint32_t ok;
    if (cluster_scope)  // cluster-scope acquire: costs an L1 invalidate (CCTL.IVALL) per successful wait — not used by
                        // the tile loops (a cta-scope wait observes remote arrivals too; what they order is TMA / TMEM
                        // traffic, fenced by tcgen05.fence / the async proxy, not generic-proxy data in L1)
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(ptx::smem_u32(bar)), "r"(parity)
          : "memory");
    else
      ok = ptx::mbar_try_wait(bar, parity) ? 1u : 0u;
    if (ok) return;
    if (spins > (1ull << 26)) {  // every try_wait poll already blocks for a HW time slice: seconds in total
      if ((threadIdx.x & 31) == 0)
        printf("gemm_put_2sm: barrier timeout cta=%d rank=%d warp=%d what=%d index=%d parity=%u\n",
               static_cast<int>(blockIdx.x), static_cast<int>(cluster_cta_rank()), static_cast<int>(threadIdx.x >> 5),
               what, index, parity);
      __trap();
    }
  }
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
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// Epilogue walk of one 128x256 accumulator (one epilogue warp = 32 accumulator rows): TMEM -> registers -> padded
// smem transpose -> 16-byte pieces of full 128-byte row segments.  One round moves a 128-byte segment per row: 32
// fp32 columns, or 64 bf16 columns (two TMEM loads, converted before staging), so whatever emit() does with a piece
// (store, peer store, reduction) always happens in full 128-byte segments: 8 consecutive lanes = one segment, 4 rows
// per instruction.  emit(row, col, byte, v): 16 bytes `v` of accumulator row `row` (0..31 inside this warp's rows)
// that start `byte` bytes into the round's segment, whose first column is `col`.
// pre(x): applied to every fp32 accumulator value before it is converted / staged (a fused activation).
struct EpilogueIdentity {
  __device__ __forceinline__ float operator()(float x) const { return x; }
};
template <class Emit, class Pre = EpilogueIdentity>
__device__ __forceinline__ void epilogue_segments(bool out_bf16, uint32_t taddr, float* stage_buf, int lane, Emit emit,
                                                  Pre pre = Pre{}) {
  const int cols_per_round = out_bf16 ? 64 : 32;
  unsigned char* stage_row = reinterpret_cast<unsigned char*>(stage_buf) + lane * (kStageRowWords * 4);
  for (int col = 0; col < kBN; col += cols_per_round) {
    if (out_bf16) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + col + half * 32, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(pre(__uint_as_float(r[j])));
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 pk;
          __nv_bfloat162 t;
          t = __floats2bfloat162_rn(__uint_as_float(r[j]), __uint_as_float(r[j + 1]));
          pk.x = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          pk.y = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5]));
          pk.z = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(__uint_as_float(r[j + 6]), __uint_as_float(r[j + 7]));
          pk.w = *reinterpret_cast<uint32_t*>(&t);
          *reinterpret_cast<uint4*>(stage_row + half * 64 + j * 2) = pk;
        }
      }
    } else {
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr + col, r);
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(pre(__uint_as_float(r[j])));
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<uint4*>(stage_row + j * 4) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = it * 4 + (lane >> 3);
      const int c16 = lane & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(stage_buf) +
                                                      row * (kStageRowWords * 4) + c16 * 16);
      emit(row, col, c16 * 16, v);
    }
    __syncwarp();
  }
}

// The plain epilogue: the same walk with the stores written out (kept in this explicit form on purpose: it is the
// code of the GPU-validated gemm_put kernels, and the lambda form above does not compile to byte-identical SASS).
// G: any struct with c_local, c_peer (either may be null), out_bf16 and n (the row length of C).
template <class G>
__device__ __forceinline__ void epilogue_store_tile(const G& g, uint32_t taddr, float* stage_buf, int m0,
                                                    int n0, int ew, int lane) {
  // One round moves a 128-byte row segment per accumulator row: 32 fp32 columns, or 64 bf16
  // columns (two TMEM loads, converted before staging), so the NVLink / HBM stores below are
  // always full 128-byte segments.
  const int cols_per_round = g.out_bf16 ? 64 : 32;
  const size_t elem = g.out_bf16 ? 2 : 4;
  unsigned char* stage_row = reinterpret_cast<unsigned char*>(stage_buf) + lane * (kStageRowWords * 4);
  for (int col = 0; col < kBN; col += cols_per_round) {
    if (g.out_bf16) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + col + half * 32, r);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 pk;
          __nv_bfloat162 t;
          t = __floats2bfloat162_rn(__uint_as_float(r[j]), __uint_as_float(r[j + 1]));
          pk.x = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          pk.y = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5]));
          pk.z = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(__uint_as_float(r[j + 6]), __uint_as_float(r[j + 7]));
          pk.w = *reinterpret_cast<uint32_t*>(&t);
          *reinterpret_cast<uint4*>(stage_row + half * 64 + j * 2) = pk;
        }
      }
    } else {
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr + col, r);
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<uint4*>(stage_row + j * 4) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
    }
    __syncwarp();
    // 8 lanes x 16 B per row, 4 rows per instruction.
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = it * 4 + (lane >> 3);
      const int c16 = lane & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(stage_buf) +
                                                      row * (kStageRowWords * 4) + c16 * 16);
      const size_t off = (static_cast<size_t>(m0 + ew * 32 + row) * g.n + n0 + col) * elem + c16 * 16;
      if (g.c_peer != nullptr)
        ptx::st_stream_v4(reinterpret_cast<uint4*>(static_cast<unsigned char*>(g.c_peer) + off), v);
      if (g.c_local != nullptr)
        *reinterpret_cast<uint4*>(static_cast<unsigned char*>(g.c_local) + off) = v;
    }
    __syncwarp();
  }
}

// smem [32 x 128 B] tile (SWIZZLE_128B layout, 1024-byte aligned) added into a 2-D fp32 tensor at element (x, y) by
// the TMA unit: one instruction per 4 KiB instead of 256 REDG requests from the LSU.
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, int x, int y, const void* smem_src) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map),
               "r"(x), "r"(y), "r"(ptx::smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int x, int y, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x),
               "r"(y), "r"(ptx::smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Epilogue walk for the TMA unit: the warp's 32 accumulator rows leave in [32 rows x 128 bytes] pieces (32 fp32 or
// 64 bf16 columns) staged in dense, 128-byte-swizzled shared memory — lane = row, the 16-byte chunk c of row r goes
// to chunk c ^ (r % 8), conflict-free for the writer and the layout a SWIZZLE_128B tensor map expects — and lane 0
// calls issue(tile, col) (one or more cp.async.bulk.tensor / cp.reduce.async.bulk.tensor instructions on a tensor
// map with a 32-row x 128-byte box), which this function commits as one bulk group.  Two 4 KiB tiles per warp,
// carved out of the (1024-byte aligned) epilogue area: the TMA reads one while the warp fills the other.
// The caller's finish() must wait for the groups (lane 0 of warps 4..7: bulk_wait<0>) before it signals anything.
constexpr uint32_t kTmaTileBytes = 32 * 128;
constexpr uint32_t kTmaEpiSmemBytes = kEpiWarps * 2 * kTmaTileBytes;  // 32 KiB instead of the 18 KiB transpose staging
template <class Issue>
__device__ __forceinline__ void epilogue_tma_tiles(bool out_bf16, uint32_t taddr, float* stage_buf, int ew, int lane,
                                                   Issue issue) {
  unsigned char* epi_base = reinterpret_cast<unsigned char*>(stage_buf) - static_cast<size_t>(ew) * kEpiWarpBytes;
  unsigned char* tiles = epi_base + static_cast<size_t>(ew) * 2 * kTmaTileBytes;
  const int cols_per_round = out_bf16 ? 64 : 32;
  int round = 0;
#pragma unroll 1
  for (int col = 0; col < kBN; col += cols_per_round, ++round) {
    unsigned char* tile = tiles + (round & 1) * kTmaTileBytes;
    // The group issued two rounds ago used this tile: it must have finished READING it (groups are per thread, lane
    // 0 issued them all; at most the newest one — the other tile — may still be pending).
    if (lane == 0) ptx::bulk_wait_read<1>();
    __syncwarp();
    unsigned char* row = tile + lane * 128;
    if (out_bf16) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + col + half * 32, r);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 pk;
          __nv_bfloat162 t;
          t = __floats2bfloat162_rn(__uint_as_float(r[j]), __uint_as_float(r[j + 1]));
          pk.x = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          pk.y = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5]));
          pk.z = *reinterpret_cast<uint32_t*>(&t);
          t = __floats2bfloat162_rn(__uint_as_float(r[j + 6]), __uint_as_float(r[j + 7]));
          pk.w = *reinterpret_cast<uint32_t*>(&t);
          const int c = half * 4 + j / 8;  // 16-byte chunk of the 128-byte row
          *reinterpret_cast<uint4*>(row + ((c ^ (lane & 7)) << 4)) = pk;
        }
      }
    } else {
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr + col, r);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(row + ((c ^ (lane & 7)) << 4)) =
            make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
    }
    fence_proxy_async_smem();  // my generic-proxy writes before the async-proxy read of the tile
    __syncwarp();
    if (lane == 0) {
      issue(tile, col);
      ptx::bulk_commit();
    }
  }
}
// finish() part of the TMA epilogues: every bulk group of this CTA must be PERFORMED before anything is signalled
// (and must have stopped reading shared memory before the CTA retires).
__device__ __forceinline__ void epilogue_tma_drain() {
  if (threadIdx.x >= 128 && (threadIdx.x & 31) == 0) {
    ptx::bulk_wait<0>();
    asm volatile("fence.proxy.async;" ::: "memory");
  }
}

// ---------------------------------------------------------------- persistent GEMM main loop ----
// The warp-specialised persistent tile loop every GEMM kernel of the suite runs (256 threads, one CTA per SM):
//   warp 0      TMA producer  : A[128x64] + B[256x64] per stage, kStagesT-deep smem ring, full/empty mbarriers
//   warp 1      MMA issuer    : one elected thread, tcgen05.mma.cta_group::1 (M128 N256 K16); tcgen05.commit
//                               frees the smem stage / publishes the accumulator
//   warp 2      TMEM allocator: 512 columns = two 128x256 fp32 accumulators (epilogue of tile i overlaps the
//                               main loop of tile i+1)
//   warp 3      policy.aux_warp(): idle for a plain GEMM, the gather engine of all-gather -> GEMM
//   warps 4..7  epilogue      : policy.epilogue() per accumulator (tcgen05.ld -> stores / reductions / puts)
// kCluster == 2: clusters of two CTAs working on vertically adjacent tiles (same n_blk).  Both need the same B
// tile, so each CTA fetches half of it (128 rows) and TMA-multicasts it into both CTAs' shared memory:
// L2->SM operand traffic drops from 48 to 32 KiB per CTA per k-block.  A stage may only be refilled when BOTH
// CTAs' MMAs have consumed it: the empty barriers count 2 arrivals and every tcgen05.commit is multicast.
//
// What differs between the kernels is the Policy:
//   void coords(int tile, int* m_blk, int* n_blk) const    tile index -> tile coordinates (rasterisation)
//   void a_rows_ready(int m_blk) const                      producer warp, before the first A load of a tile
//   void epilogue(uint32_t taddr, float* stage_buf, int m0, int n0, int ew, int lane) const
//   void aux_warp(int lane, unsigned char* aux_smem) const  warp 3, runs concurrently with the tile loop
//   void finish() const                                     all threads, after the loop and the TMEM release
//   static constexpr bool kHasAuxWarp                       false: warp 3 idles (the branch is compiled out)
// aux_smem points behind the epilogue staging buffers; its size is whatever the launcher added to
// gemm_smem_bytes<kStagesT>(aux_bytes).
template <int kStagesT>
constexpr size_t gemm_smem_bytes(uint32_t aux_bytes) {
  return static_cast<size_t>(kStagesT) * (kABytes + kBBytes) + kEpiWarps * kEpiWarpBytes + aux_bytes + 1024;
}

template <int kCluster, int kStagesT, class Policy>
__device__ __forceinline__ void gemm_persistent(const CUtensorMap& map_a, const CUtensorMap& map_b, int tiles_m,
                                                int tiles_n, int k, const Policy& policy) {
  constexpr uint32_t kStageBytesT = kABytes + kBBytes;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStagesT];
  __shared__ __align__(8) uint64_t empty_bar[kStagesT];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;

  unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* epi_smem = smem + static_cast<size_t>(kStagesT) * kStageBytesT;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = k / kBK;
  const int crank = kCluster > 1 ? static_cast<int>(cluster_cta_rank()) : 0;
  // Work items are tile pairs in cluster mode: pair p -> tiles 2p, 2p+1 (consecutive tiles of the
  // grouped rasterisation share n_blk when the group height is even).
  const int first_item = static_cast<int>(blockIdx.x) / kCluster;
  const int item_stride = static_cast<int>(gridDim.x) / kCluster;
  const int num_items = num_tiles / kCluster;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < kStagesT; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], kCluster);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], kEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     ptx::smem_u32(&tmem_base_s)),
                 "r"(static_cast<uint32_t>(kTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (kCluster > 1) cluster_sync_all();  // the peer's barriers exist before anything remote touches them

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp walks the loop (convergent barriers at kernel end); lane 0 issues the TMA.
    int stage = 0;
    uint32_t phase = 0;
    for (int item = first_item; item < num_items; item += item_stride) {
      const int tile = item * kCluster + crank;
      int m_blk, n_blk;
      policy.coords(tile, &m_blk, &n_blk);
      const int m0 = m_blk * kBM;
      const int n0 = n_blk * kBN;
      policy.a_rows_ready(m_blk);
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);  // slot released by the MMA warp(s)
        if (lane == 0) {
          unsigned char* sa = smem + static_cast<size_t>(stage) * kStageBytesT;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], kStageBytesT);
          tma_load_2d(sa, &map_a, kb * kBK, m0, &full_bar[stage]);
          if (kCluster > 1)  // my half of the shared B tile, delivered to both CTAs of the pair
            tma_load_2d_multicast(sa + kABytes + crank * (kBBytes / 2), &map_b, kb * kBK,
                                  n0 + crank * (kBN / 2), &full_bar[stage], static_cast<uint16_t>(0x3));
          else
            tma_load_2d(sa + kABytes, &map_b, kb * kBK, n0, &full_bar[stage]);
        }
        __syncwarp();
        if (++stage == kStagesT) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc(kBM, kBN);
    int stage = 0;
    uint32_t phase = 0;
    int local_tile = 0;
    for (int item = first_item; item < num_items; item += item_stride, ++local_tile) {
      const int acc = local_tile & 1;
      const uint32_t acc_phase = static_cast<uint32_t>(local_tile >> 1) & 1;
      ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);  // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kBN);
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);  // TMA bytes landed
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(smem + static_cast<size_t>(stage) * kStageBytesT);
          const uint64_t desc_a = make_smem_desc(sa);
          const uint64_t desc_b = make_smem_desc(sa + kABytes);
#pragma unroll
          for (int kk = 0; kk < kBK / kUmmaK; ++kk) {
            const uint64_t adv = static_cast<uint64_t>((kk * kUmmaK * 2) >> 4);
            umma_bf16(tmem_d, desc_a + adv, desc_b + adv, idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          if (kCluster > 1)
            umma_commit_multicast(&empty_bar[stage], static_cast<uint16_t>(0x3));  // frees the slot in both CTAs
          else
            umma_commit(&empty_bar[stage]);                            // frees the smem slot
          if (kb == num_kb - 1) umma_commit(&tmem_full_bar[acc]);      // accumulator complete
        }
        __syncwarp();
        if (++stage == kStagesT) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (Policy::kHasAuxWarp && warp == 3) {
    policy.aux_warp(lane, epi_smem + static_cast<size_t>(kEpiWarps) * kEpiWarpBytes);
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;            // 0..3 == warp % 4 -> TMEM lane group
    float* stage_buf = reinterpret_cast<float*>(epi_smem + static_cast<size_t>(ew) * kEpiWarpBytes);
    int local_tile = 0;
    for (int item = first_item; item < num_items; item += item_stride, ++local_tile) {
      const int tile = item * kCluster + crank;
      const int acc = local_tile & 1;
      const uint32_t acc_phase = static_cast<uint32_t>(local_tile >> 1) & 1;
      int m_blk, n_blk;
      policy.coords(tile, &m_blk, &n_blk);
      const int m0 = m_blk * kBM;
      const int n0 = n_blk * kBN;
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * kBN) + (static_cast<uint32_t>(ew * 32) << 16);
      policy.epilogue(taddr, stage_buf, m0, n0, ew, lane);
      tc_fence_before();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);  // accumulator may be overwritten
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(kTmemCols))
                 : "memory");
  // A CTA of a pair must not retire while its partner can still multicast into it.
  if (kCluster > 1) cluster_sync_all();
  policy.finish();
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
constexpr size_t gemm_2sm_smem_bytes(uint32_t aux_bytes) {
  return static_cast<size_t>(kStages2) * kStageBytes2 + kEpiWarps * kEpiWarpBytes + aux_bytes + 1024;
}

// Same Policy contract as gemm_persistent (coords() is called with tile = 2 * item + cluster rank: the two tiles of
// an item must be vertically adjacent with the same n_blk, the leader's on top).
template <class Policy>
__device__ __forceinline__ void gemm_persistent_2sm(const CUtensorMap& map_a, const CUtensorMap& map_b, int tiles_m,
                                                    int tiles_n, int k, const Policy& policy) {
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
  const int num_kb = k / kBK;
  const int crank = static_cast<int>(cluster_cta_rank());
  const bool leader = crank == 0;
  // Work items are 256x256 tiles = vertically adjacent pairs (2p, 2p+1) of the grouped rasterisation.
  const int first_item = static_cast<int>(blockIdx.x) / 2;
  const int item_stride = static_cast<int>(gridDim.x) / 2;
  const int num_items = tiles_m * tiles_n / 2;

  cluster_sync_all();  // both CTAs of the pair are resident before the pair-wide TMEM allocation
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < kStages2; ++s) {
      // ONE arrival per phase: the leader's arrive.expect_tx for the bytes of BOTH CTAs.  The partner's TMA loads signal
      // this barrier directly (cta_group::2 complete_tx); a per-stage remote arrive from the partner costs a
      // cluster-scope release (MEMBAR.ALL.GPU) per k-block and paced the whole loop at 30 % tensor activity
      // (profiles/r2_call6_1gpu/prof_gemm_2sm).  Early complete_tx bytes only drive the tx-count negative until the
      // leader's expect_tx arrives; the phase cannot complete before that arrival.
      ptx::mbar_init(&full_bar[s], 1);
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
      policy.coords(item * 2 + crank, &m_blk, &n_blk);
      const int m0 = m_blk * kBM;
      const int n0 = n_blk * kBN + crank * (kBN / 2);  // my half of the shared B tile
      policy.a_rows_ready(m_blk);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_or_trap(&empty_bar[stage], phase ^ 1, false, kWaitEmpty, stage);  // slot released by the pair's MMA
        if (lane == 0) {
          unsigned char* sa = smem + static_cast<size_t>(stage) * kStageBytes2;
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes2);
          tma_load_2d_2sm(sa, &map_a, kb * kBK, m0, &full_bar[stage]);
          tma_load_2d_2sm(sa + kABytes, &map_b, kb * kBK, n0, &full_bar[stage]);
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
      mbar_wait_or_trap(&tmem_empty_bar[acc], acc_phase ^ 1, false, kWaitTmemEmpty, acc);  // both epilogues drained it
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kBN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_or_trap(&full_bar[stage], phase, false, kWaitFull, stage);  // both CTAs' TMA bytes landed
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
  } else if (Policy::kHasAuxWarp && warp == 3) {
    policy.aux_warp(lane, epi_smem + static_cast<size_t>(kEpiWarps) * kEpiWarpBytes);
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int ew = warp - 4;
    float* stage_buf = reinterpret_cast<float*>(epi_smem + static_cast<size_t>(ew) * kEpiWarpBytes);
    int local_tile = 0;
    for (int item = first_item; item < num_items; item += item_stride, ++local_tile) {
      const int acc = local_tile & 1;
      const uint32_t acc_phase = static_cast<uint32_t>(local_tile >> 1) & 1;
      int m_blk, n_blk;
      policy.coords(item * 2 + crank, &m_blk, &n_blk);
      mbar_wait_or_trap(&tmem_full_bar[acc], acc_phase, false, kWaitTmemFull, acc);
      tc_fence_after();
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * kBN) + (static_cast<uint32_t>(ew * 32) << 16);
      policy.epilogue(taddr, stage_buf, m_blk * kBM, n_blk * kBN, ew, lane);
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
  policy.finish();
}

inline PFN_cuTensorMapEncodeTiled gemm_tensor_map_encoder() {
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
  HPCP_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available");
  return fn;
}

// 2-D fp32 / bf16 [rows, cols] row-major tensor with a [32 rows x 128 bytes] SWIZZLE_128B box: the destination of
// the TMA epilogues (cp.async.bulk.tensor.2d store, cp.reduce.async.bulk.tensor.2d ... .add).
inline CUtensorMap make_c_tile_map(const void* base, int rows, int cols, bool bf16) {
  CUtensorMap map;
  const size_t elem = bf16 ? 2 : 4;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * elem};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elem), 32};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult r = gemm_tensor_map_encoder()(
      &map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims,
      strides, box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HPCP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (C tile map) failed with code " + std::to_string(r));
  return map;
}

// 2-D bf16 [rows, k] K-major tensor with a [box_rows x 64] SWIZZLE_128B box.
inline CUtensorMap make_kmajor_map(const void* base, int rows, int k, int box_rows) {
  CUtensorMap map;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(k) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult r = gemm_tensor_map_encoder()(
      &map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, elem_strides,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HPCP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string(r));
  return map;
}

}  // namespace umma
}  // namespace hpcp
