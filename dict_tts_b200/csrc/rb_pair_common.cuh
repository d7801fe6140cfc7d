// Shared by the fused ResBlock-pair kernels (rb_pair.cu: C = 32 / 64, rb_pair128.cu: C = 128): tile geometry and the
// row-tile -> (item, first row) decode of a uniform or ragged launch.
#pragma once
#include "tc_conv.cuh"

namespace dtts {
namespace {

constexpr int kPairThreads = 96 + 8 * 32;
constexpr int kPairRows = 256;                 // rows per tile (two 128-row MMA sub-tiles)

struct PairTile { int b, q0, lim; };
struct PairCursor { int b = 0; uint32_t base = 0; };
// row tile rt (tiles of all items back to back) -> (item, first output row, row limit of the item)
__device__ __forceinline__ PairTile pair_decode(const RbPairParams& p, const int* pref, const int* limv, uint32_t rt,
                                                PairCursor& cur) {
  PairTile c;
  if (pref) {
    int b = cur.b;
    while (b + 1 < p.B && (uint32_t)pref[b + 1] <= rt) ++b;
    cur.b = b;
    c.b = b;
    c.q0 = (int)(rt - (uint32_t)pref[b]) * p.S;
    c.lim = limv[b];
  } else {
    const uint32_t nt = (uint32_t)p.ntiles;
    while (rt >= cur.base + nt) { cur.base += nt; ++cur.b; }
    c.b = cur.b;
    c.q0 = (int)(rt - cur.base) * p.S;
    c.lim = p.T;
  }
  return c;
}

}  // namespace

// C = 128 (CTA pairs, rb_pair128.cu)
int rb_pair128_supported(const TcConvW& c1, const TcConvW& c2, int dil, int a_planes);
cudaError_t launch_rb_pair128(RbPairParams p, cudaStream_t stream);

}  // namespace dtts
