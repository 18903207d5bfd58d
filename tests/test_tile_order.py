"""CPU checks of the tile and gather-piece orderings of the tensor-core kernels (csrc/kernels/tile_order.h):
the same integer functions the device code calls, evaluated on the host through the extension."""
import itertools

from hypothesis import given, settings, strategies as st

import hpc_patterns_b200


@given(tiles_m=st.integers(1, 70), tiles_n=st.integers(1, 40))
@settings(max_examples=60, deadline=None)
def test_grouped_rasterisation_is_a_bijection(tiles_m, tiles_n):
    C = hpc_patterns_b200.native()
    seen = {C.gemm_tile_coords(t, tiles_m, tiles_n) for t in range(tiles_m * tiles_n)}
    assert seen == set(itertools.product(range(tiles_m), range(tiles_n)))


@given(tiles_m=st.integers(1, 40).map(lambda v: 2 * v).filter(lambda v: (v % 8) % 2 == 0), tiles_n=st.integers(1, 20))
@settings(max_examples=40, deadline=None)
def test_cta_pairs_get_vertically_adjacent_tiles(tiles_m, tiles_n):
    """Cluster mode pairs tiles (2p, 2p+1): same n_blk (they share the B tile), m_blk and m_blk + 1."""
    C = hpc_patterns_b200.native()
    for p in range(tiles_m * tiles_n // 2):
        (m0, n0), (m1, n1) = C.gemm_tile_coords(2 * p, tiles_m, tiles_n), C.gemm_tile_coords(2 * p + 1, tiles_m, tiles_n)
        assert n0 == n1 and m1 == m0 + 1 and m0 % 2 == 0


@given(world=st.integers(1, 8), shard_tiles_m=st.integers(1, 12), tiles_n=st.integers(1, 9), first=st.integers(0, 1),
       data=st.data())
@settings(max_examples=60, deadline=None)
def test_shard_major_order(world, shard_tiles_m, tiles_n, first, data):
    """Every rank visits every tile once, shard after shard, starting with rank+first; at any position in the
    order the `world` ranks are working on `world` different shards (every link busy, no hot spot)."""
    C = hpc_patterns_b200.native()
    rank = data.draw(st.integers(0, world - 1))
    per_shard = shard_tiles_m * tiles_n
    order = [C.gemm_shard_coords(t, rank, world, first, shard_tiles_m, tiles_n) for t in range(world * per_shard)]
    assert set(order) == set(itertools.product(range(world * shard_tiles_m), range(tiles_n)))
    owners = [m // shard_tiles_m for m, _ in order]
    assert owners == [(rank + first + t // per_shard) % world for t in range(world * per_shard)]
    t = data.draw(st.integers(0, world * per_shard - 1))
    at_t = {C.gemm_shard_coords(t, r, world, first, shard_tiles_m, tiles_n)[0] // shard_tiles_m for r in range(world)}
    assert len(at_t) == world


@given(world=st.integers(2, 8), shard_tiles_m=st.integers(1, 4), k64=st.integers(1, 8),
       chunk=st.sampled_from([512, 1024, 2048, 4096]), grid=st.integers(1, 148), data=st.data())
@settings(max_examples=40, deadline=None)
def test_gather_pieces_cover_every_remote_byte_once(world, shard_tiles_m, k64, chunk, grid, data):
    C = hpc_patterns_b200.native()
    rank = data.draw(st.integers(0, world - 1))
    k = 64 * k64
    block_bytes = 128 * k * 2
    if block_bytes % chunk:
        return
    cpb = C.allgather_gemm_chunks_per_block(k, chunk)
    assert cpb == block_bytes // chunk
    total = (world - 1) * shard_tiles_m * cpb
    covered = {}
    per_block = {}
    last_peer_index = -1
    for c in range(total):
        peer, m_blk, src_off, dst_off = C.gemm_gather_piece(c, rank, world, shard_tiles_m, cpb, chunk, block_bytes)
        assert peer != rank and 0 <= peer < world
        assert m_blk // shard_tiles_m == peer
        assert dst_off == peer * shard_tiles_m * block_bytes + src_off      # same place inside the peer's rows
        assert 0 <= src_off and src_off + chunk <= shard_tiles_m * block_bytes
        assert m_blk == dst_off // block_bytes                               # counted on the block it lands in
        assert dst_off not in covered
        covered[dst_off] = c
        per_block[m_blk] = per_block.get(m_blk, 0) + 1
        idx = (peer - rank - 1) % world                                      # peers in the order rank+1, rank+2, ...
        assert idx >= last_peer_index
        last_peer_index = idx
    assert len(covered) * chunk == (world - 1) * shard_tiles_m * block_bytes
    assert set(per_block.values()) == {cpb}                                  # every counter reaches the target
    # dealt round-robin: the pieces of any CTA are c = cta, cta + grid, ... -> shares differ by at most one
    shares = [len(range(cta, total, grid)) for cta in range(grid)]
    assert max(shares) - min(shares) <= 1
