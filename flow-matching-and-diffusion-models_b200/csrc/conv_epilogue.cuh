// Register-level pieces of the conv epilogues, shared by the per-tile kernel (conv_igemm.cu) and the rolling-row
// kernel (conv_rolling.cuh).  A thread owns one output pixel (TMEM lane) and works on 32 consecutive fp32 columns at a
// time; the staging tile in shared memory is laid out exactly as TMA's 128-byte swizzle expects it (rows of 64
// channels, 16-byte chunk index XOR (row & 7)), so one TMA store per 64-channel slab writes it out.
#pragma once

namespace fm {

// v[0..31] += bias[0..31] (fp32, shared memory, 16-byte aligned)
__device__ __forceinline__ void epi_add_bias(float (&v)[32], const float* sb) {
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 b = *reinterpret_cast<const float4*>(sb + j4 * 4);
    const float2 s0 = fadd2(make_float2(v[j4 * 4 + 0], v[j4 * 4 + 1]), make_float2(b.x, b.y));
    const float2 s1 = fadd2(make_float2(v[j4 * 4 + 2], v[j4 * 4 + 3]), make_float2(b.z, b.w));
    v[j4 * 4 + 0] = s0.x; v[j4 * 4 + 1] = s0.y; v[j4 * 4 + 2] = s1.x; v[j4 * 4 + 3] = s1.y;
  }
}

// v[0..31] += the bf16 residual staged at this row of the swizzled tile (chunks chunk0 .. chunk0+3 of the 128-byte row)
__device__ __forceinline__ void epi_add_residual(float (&v)[32], const uint8_t* rowp, int chunk0, int row) {
#pragma unroll
  for (int j8 = 0; j8 < 4; ++j8) {
    const uint4 rr = *reinterpret_cast<const uint4*>(rowp + (((chunk0 + j8) ^ (row & 7)) * 16));
    const uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 s = fadd2(make_float2(v[j8 * 8 + 2 * k], v[j8 * 8 + 2 * k + 1]), unpack_bf16x2(w[k]));
      v[j8 * 8 + 2 * k] = s.x;
      v[j8 * 8 + 2 * k + 1] = s.y;
    }
  }
}

// GroupNorm partial statistics of the warp's 32 rows: per channel quad (4 channels) the sum and the sum of squares,
// reduced with a recursive-halving butterfly (16 values -> 16 + 1 shuffles, fixed order => deterministic).  Returns this
// lane's value; `vidx` says which: quad (vidx & 7) of the 8 quads in these 32 columns, sum (vidx < 8) or sum of squares.
// Lanes 2k and 2k+1 hold the same value; the caller lets the even lane write it.
__device__ __forceinline__ float epi_quad_stats(const float (&v)[32], bool valid, int lane, int& vidx) {
  float red[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float a0 = valid ? v[4 * j + 0] : 0.f, a1 = valid ? v[4 * j + 1] : 0.f;
    const float a2 = valid ? v[4 * j + 2] : 0.f, a3 = valid ? v[4 * j + 3] : 0.f;
    red[j] = (a0 + a1) + (a2 + a3);
    red[8 + j] = fmaf(a0, a0, a1 * a1) + fmaf(a2, a2, a3 * a3);
  }
#pragma unroll
  for (int width = 8, mask = 16; width >= 1; width >>= 1, mask >>= 1) {
    const bool upper = (lane & mask) != 0;
#pragma unroll
    for (int i = 0; i < width; ++i) {
      const float keep = upper ? red[i + width] : red[i];
      const float give = upper ? red[i] : red[i + width];
      red[i] = keep + __shfl_xor_sync(0xffffffffu, give, mask);
    }
  }
  red[0] += __shfl_xor_sync(0xffffffffu, red[0], 1);
  vidx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  return red[0];
}

// convert to bf16 and store the 32 columns into this row of the swizzled staging tile
__device__ __forceinline__ void epi_pack_store(const float (&v)[32], uint8_t* rowp, int chunk0, int row) {
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int chunk = (chunk0 + j) ^ (row & 7);
    *reinterpret_cast<uint4*>(rowp + chunk * 16) = make_uint4(pk[4 * j + 0], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  }
}

}  // namespace fm
