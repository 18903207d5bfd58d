"""The hand-built tcgen05 instruction descriptor (csrc/kernels/umma.cuh: make_idesc) against the one CUTLASS builds
for the same MMA (headers vendored in the image under flashinfer/data/cutlass/include; skipped when absent).
M = 128 is what the GPU-validated kernels use; M = 256 is the 2-SM UMMA form that has not run yet."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r"""
#include "kernels/umma.cuh"   // first: it needs the driver-API typedefs before CUTLASS pulls in its own CUDA headers
#include <cstdio>
#include <cute/tensor.hpp>
#include <cute/arch/mma_sm100_desc.hpp>
#include <cute/atom/mma_traits_sm100.hpp>
using namespace cute;
template <int M, int N> void row() {
  auto d = UMMA::make_instr_desc<bfloat16_t, bfloat16_t, float, M, N, UMMA::Major::K, UMMA::Major::K>();
  printf("%d %d %08x %08x\n", M, N, uint32_t(d), hpcp::umma::make_idesc(M, N));
}
int main() {
  row<128, 256>(); row<256, 256>(); row<128, 128>(); row<256, 64>();
  // The 128-byte swizzle the TMA epilogues write by hand (umma.cuh: epilogue_tma_tiles: chunk c of row r at
  // c ^ (r % 8)) against CUTLASS's Swizzle<3,4,3>, the layout a SWIZZLE_128B tensor map reads / writes.
  int bad = 0;
  for (int r = 0; r < 32; ++r)
    for (int c = 0; c < 8; ++c)
      bad += int(Swizzle<3, 4, 3>{}(r * 128 + c * 16)) != r * 128 + ((c ^ (r & 7)) << 4);
  printf("swizzle_mismatches %d\n", bad);
  // Shared-memory matrix descriptor (K-major, SWIZZLE_128B, 8-row x 128-byte atoms every 1024 bytes) against
  // CUTLASS's bit-field definition of the same descriptor.
  int bad_desc = 0;
  for (uint32_t addr : {0x0u, 0x400u, 0xC000u, 0x2A400u, 0x38C00u}) {
    UMMA::SmemDescriptor d;
    d.start_address_ = uint16_t((addr & 0x3FFFF) >> 4);
    d.leading_byte_offset_ = 1;
    d.stride_byte_offset_ = 1024 >> 4;
    d.version_ = 1;
    d.base_offset_ = 0;
    d.lbo_mode_ = 0;
    d.layout_type_ = uint8_t(UMMA::LayoutType::SWIZZLE_128B);
    bad_desc += uint64_t(d) != hpcp::umma::make_smem_desc(addr);
  }
  printf("smem_desc_mismatches %d\n", bad_desc);
  return 0;
}
"""


def _cutlass_include():
    import importlib.util

    spec = importlib.util.find_spec("flashinfer")          # located, not imported
    if spec is None or not spec.submodule_search_locations:
        return None
    d = os.path.join(list(spec.submodule_search_locations)[0], "data", "cutlass", "include")
    return d if os.path.exists(os.path.join(d, "cute", "arch", "mma_sm100_desc.hpp")) else None


def test_instruction_descriptor_matches_cutlass(tmp_path):
    inc = _cutlass_include()
    if inc is None:
        pytest.skip("no vendored CUTLASS headers in this image")
    src = tmp_path / "idesc.cu"
    src.write_text(PROGRAM)
    exe = tmp_path / "idesc"
    p = subprocess.run(["nvcc", "-ccbin", "/usr/bin/g++", "-std=c++17", "--expt-relaxed-constexpr", "-I", inc, "-I",
                        os.path.join(ROOT, "csrc"), "-gencode", "arch=compute_100a,code=sm_100a", str(src), "-o",
                        str(exe)], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60).stdout
    rows = [ln.split() for ln in out.splitlines() if re.match(r"\d+ \d+ ", ln)]
    assert len(rows) == 4
    for m, n, theirs, ours in rows:
        assert theirs == ours, (m, n, theirs, ours)
    assert "swizzle_mismatches 0" in out and "smem_desc_mismatches 0" in out
