// A whole ResBlock1 (three conv1 -> leaky -> conv2 -> + residual pairs, dilations d0, d1, d2) with k = 3 of the two narrow
// HiFi-GAN stages (C = 32, 64) in ONE launch.  The three fused pair launches of rb_pair32/64_kernel move 3 x 12 B per
// element through HBM and run at or near its roof; here the fp32 residual stream of a row stays in the registers of the
// epilogue thread that owns the row and the operand planes between the convolutions stay in shared memory, so the block
// reads 2 + 4 B and writes 4 B per element.  Reference: ResBlock1.forward, modules/hifigan/hifigan.py:51-58.
//
// Tile = a window of 256 rows [t0, t0 + 256), t0 = q0 - H with H = sum(d_p + 1) = 12 the receptive field of the block
// per side: every convolution computes all 256 rows, a row is exact when it is at least (rows of receptive field so
// far) away from a window edge, and only the core rows [H, 256 - H) of the last pair are stored; tiles advance by
// S = 256 - 2 H = 232 rows (10 % recompute).  Rows outside [0, T) are written as zeros into the operand tile (the zero
// padding each convolution sees).  The six convolutions are the same implicit GEMMs as the pair kernels (hi | lo weight
// planes stacked along N' = 2 C, 32-channel K chunks of two MMA K-steps per tap, the same MMA order, the same roundings),
// so the result is BIT-IDENTICAL to the three pair launches -- and therefore to the six tc_conv launches -- it replaces.
//
// ONE operand tile per slot, updated in place: conv1 reads it as the pair's input, E1 overwrites it with
// fp16(leaky(conv1 + b1)), conv2 reads that, E2 overwrites it with the next pair's input fp16(leaky(y)); its 5 margin rows
// on both sides keep whatever the input load put there, which only ever reaches rows that are not stored.
//
// Two tiles are in flight per CTA (slots 0 / 1), each walks its six phases conv1(0) conv2(0) conv1(1) ... in order; the
// MMA thread alternates between the slots, so the tensor pipe works on one tile while the other tile's epilogue group
// turns accumulators into the next operand:
//   warp 0       the input window of the next tile of a slot by bulk copy; C = 32: all weights once (72 KB, resident)
//   warp 1       C = 64: the weights of one convolution (48 KB) per phase of a tile pair through a 3-stage ring -- both
//                slots consume a stage before it is released
//   warp 2       TMEM (one 2 x N'-column accumulator per slot); one thread issues the MMAs
//   warps 3-18   two groups of eight, one per slot; thread = window row: E1 (bias, leaky, fp16 -> tile), E2 (bias,
//                + residual registers; leaky, fp16 -> tile for the next pair, or the fp32 stream / output planes after
//                the last pair).  With ONE group serving both slots the C = 32 kernel was epilogue bound (0.75 ms
//                against 0.68: E(A) then E(B) per phase next to 2 x 600 cycles of MMAs)
#include "tc_conv.cuh"
#include "tc16.cuh"
#include "tc_ptx.cuh"
#include "rb_pair_common.cuh"

#include <mutex>

namespace dtts {

namespace {

constexpr int kBK = 3;                             // kernel size
constexpr int kBMaxDil = 5;                        // margin rows of the operand tile on both sides
constexpr int kBRP = kPairRows + 2 * kBMaxDil;     // rows of an operand tile (window row w at buffer row w + kBMaxDil)
constexpr int kBThreads = 96 + 16 * 32;            // warps 0-2: producers / MMA; warps 3-18: two epilogue groups, one per slot
constexpr int kBWStages = 3;                       // C = 64: weight ring
// mbarriers, per slot: input landed / slot may be re-filled / accumulator full / operand written (or accumulator drained);
// weights: C = 32 kBWFull once, C = 64 kBWFull + stage / kBWEmpty + stage
constexpr int kBInFull = 0, kBSlotFree = 2, kBAccFull = 4, kBReady = 6, kBWFull = 8, kBWEmpty = kBWFull + kBWStages,
              kBNumBars = kBWEmpty + kBWStages;
constexpr int kBTmemOff = kBNumBars * 8;
constexpr int kBBiasOff = 128;                     // 6 x C floats

template <int C>
struct BlockGeom {
  static constexpr int NM = 2 * C;                              // MMA N (hi | lo stacked)
  static constexpr int NCH = C / 32;                            // 32-channel K chunks
  static constexpr int TapBytes = NM * 32 * 2;                  // one (chunk, tap) blob
  static constexpr int ConvBytes = NCH * kBK * TapBytes;        // one convolution: 12 KB (C = 32) / 48 KB (C = 64)
  static constexpr int TileBytes = (C / 8) * kBRP * 16;         // operand tile of one slot
  static constexpr bool Resident = C == 32;                     // all six convolutions' weights stay in shared memory
  static constexpr int WBytes = Resident ? 6 * ConvBytes : kBWStages * ConvBytes;
  static constexpr int PrefOff = kBBiasOff + 6 * C * 4;
  static constexpr int Header = (PrefOff + (2 * TC_MAX_RAGGED_ITEMS + 8) * 4 + 127) / 128 * 128;
  static constexpr int Smem = Header + WBytes + 2 * TileBytes;
  static constexpr int CW = C == 32 ? 16 : 8;                   // accumulator columns an epilogue thread holds at a time
};

template <int C>
__global__ void __launch_bounds__(kBThreads, 1) rb_block_kernel(const RbBlockParams p) {
  using G_ = BlockGeom<C>;
  constexpr int NM = G_::NM, NCH = G_::NCH, CW = G_::CW;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.halo;
  const uint32_t bar0 = smem_u32(smem);
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + kBTmemOff);
  float* bias_s = reinterpret_cast<float*>(smem + kBBiasOff);
  int* pref_s = reinterpret_cast<int*>(smem + G_::PrefOff);
  int* lim_s = pref_s + TC_MAX_RAGGED_ITEMS + 1;
  const uint32_t w_base = smem_u32(smem + G_::Header);                  // C = 32: conv1(0) conv2(0) ...; C = 64: ring stages
  const uint32_t p_base = w_base + (uint32_t)G_::WBytes;                // operand tile of slot 0, slot 1

  griddep_launch();
  RbPairParams pp{};                                 // the tile decode of the pair kernels (rb_pair_common.cuh)
  pp.B = p.B; pp.S = p.S; pp.ntiles = p.ntiles; pp.T = p.T;
  if (p.lens && warp == 3) {                         // per-item row limits and the exclusive prefix of their tile counts
    griddep_wait();
    int carry = 0;
    for (int b0 = 0; b0 < p.B; b0 += 32) {
      const int b = b0 + lane;
      int lim = 0;
      if (b < p.B) {
        const long v = (long)__ldg(p.lens + b) * p.len_mul + p.len_add;
        lim = v < 0 ? 0 : (v > p.T ? p.T : (int)v);
        lim_s[b] = lim;
      }
      int nt = (lim + p.S - 1) / p.S, inc = nt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
      }
      if (b < p.B) pref_s[b] = carry + inc - nt;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) pref_s[p.B] = carry;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(kBInFull + s), 1);
      mbar_init(bar(kBSlotFree + s), 1);
      mbar_init(bar(kBAccFull + s), 1);
      mbar_init(bar(kBReady + s), 8);
    }
    for (int s = 0; s < kBWStages; ++s) { mbar_init(bar(kBWFull + s), 1); mbar_init(bar(kBWEmpty + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem + kBTmemOff)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x - 96; i >= 0 && i < 6 * C; i += kBThreads - 96) {
    const int c = i / C;                                                // conv c: pair c / 2, conv1 or conv2
    bias_s[i] = __ldg(p.bias[c] + (i - c * C));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (G_::Resident && warp == 0 && elect_one()) {    // constants: may be fetched before the predecessor grid has completed
    mbar_arrive_expect_tx(bar(kBWFull), 6u * G_::ConvBytes);
    for (int c = 0; c < 6; ++c) bulk_g2s(w_base + (uint32_t)c * G_::ConvBytes, p.w[c], G_::ConvBytes, bar(kBWFull));
  }
  if (warp != 1) griddep_wait();                     // (warp 1 only reads weights)
  const uint32_t tmem_base = *tmem_ptr_s;
  const int* pref = p.lens ? pref_s : nullptr;
  const int nrt = p.lens ? pref_s[p.B] : p.ntiles * p.B;
  const int G = gridDim.x, cta = blockIdx.x;
  const int n_it = nrt > cta ? (nrt - cta + G - 1) / G : 0;           // tiles of this CTA: cta, cta + G, ...; tile i -> slot i & 1

  if (warp == 0) {
    // ------------------------------------------------ producer: the input window (+ margins) of tile i into slot i & 1
    __syncwarp();
    PairCursor cur;
    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1;
      const PairTile tc = pair_decode(pp, pref, lim_s, (uint32_t)(cta + it * G), cur);
      mbar_wait(bar(kBSlotFree + s), ((it >> 1) & 1) ^ 1);             // the previous tile of this slot has issued its last MMAs
      if (elect_one()) {
        mbar_arrive_expect_tx(bar(kBInFull + s), G_::TileBytes);
        const size_t row0 = (size_t)(p.a_pad + tc.q0 - H - kBMaxDil);
        const tc16* src = p.a_hi + (size_t)tc.b * p.a_bs;
        for (int sl = 0; sl < C / 8; ++sl)
          bulk_g2s(p_base + (uint32_t)s * G_::TileBytes + (uint32_t)sl * kBRP * 16u, src + ((size_t)sl * p.a_rows + row0) * 8,
                   (uint32_t)kBRP * 16u, bar(kBInFull + s));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------ C = 64: weight producer, one convolution per (tile pair, phase)
    if (!G_::Resident) {
      int st = 0;
      uint32_t ph_w = 1;
      for (int i0 = 0; i0 < n_it; i0 += 2) {
        for (int ph = 0; ph < 6; ++ph) {
          mbar_wait(bar(kBWEmpty + st), ph_w);
          if (elect_one()) {
            mbar_arrive_expect_tx(bar(kBWFull + st), G_::ConvBytes);
            bulk_g2s(w_base + (uint32_t)st * G_::ConvBytes, p.w[ph], G_::ConvBytes, bar(kBWFull + st));
          }
          __syncwarp();
          if (++st == kBWStages) { st = 0; ph_w ^= 1u; }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------ MMA issuer: phase ph of slot 0, phase ph of slot 1, ph = 0 .. 5
    if (elect_one()) {
      const uint32_t hiw = (128u >> 4) | (1u << 14);                  // SBO = 128 B, descriptor version 1
      const uint32_t f16b = p.fmt ? 1u : 0u;
      const uint32_t idesc = (1u << 4) | (f16b << 7) | (f16b << 10) | ((uint32_t)(NM >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t p_low0 = ((p_base >> 4) & 0x3FFFu) | ((uint32_t)kBRP << 16);
      const uint32_t w_low0 = ((w_base >> 4) & 0x3FFFu) | ((uint32_t)NM << 16);
      const uint32_t a_kstep = 2u * kBRP, a_chunk = 4u * kBRP, b_kstep = 2u * NM;
      if (G_::Resident) mbar_wait(bar(kBWFull), 0);
      uint32_t rdy_par[2] = {0u, 0u};                                   // phase parity of kBReady per slot
      int st = 0;
      uint32_t ph_w = 0;
      for (int i0 = 0; i0 < n_it; i0 += 2) {
        const int ns = n_it - i0 >= 2 ? 2 : 1;
        for (int ph = 0; ph < 6; ++ph) {
          uint32_t b_conv;
          if (G_::Resident) {
            b_conv = w_low0 + (uint32_t)ph * (G_::ConvBytes >> 4);
          } else {
            mbar_wait(bar(kBWFull + st), ph_w);                         // this phase's convolution has landed
            b_conv = w_low0 + (uint32_t)st * (G_::ConvBytes >> 4);
          }
          for (int s = 0; s < ns; ++s) {
            const int it = i0 + s;
            if (ph == 0) {
              mbar_wait(bar(kBInFull + s), (it >> 1) & 1);
              if (it >= 2) { mbar_wait(bar(kBReady + s), rdy_par[s]); rdy_par[s] ^= 1u; }   // accumulator drained
            } else {
              mbar_wait(bar(kBReady + s), rdy_par[s]); rdy_par[s] ^= 1u;                     // operand tile written
            }
            tc_fence_after();
            const uint32_t d_base = tmem_base + (uint32_t)(s * 2 * NM);
            // conv1 of pair ph / 2: taps at -d, 0, +d around the row; conv2: taps at -1, 0, +1
            const uint32_t d = (ph & 1) == 0 ? (uint32_t)p.dil[ph >> 1] : 1u;
            const uint32_t a0 = p_low0 + (uint32_t)s * (G_::TileBytes >> 4) + (uint32_t)kBMaxDil - d;
            uint32_t b_lo = b_conv;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              uint32_t a_tap = a0 + (uint32_t)c * a_chunk;
              for (int j = 0; j < kBK; ++j, a_tap += d, b_lo += (G_::TapBytes >> 4)) {
#pragma unroll
                for (int m = 0; m < 2; ++m) {
#pragma unroll
                  for (int ks = 0; ks < 2; ++ks)
                    umma_bf16(d_base + (uint32_t)(m * NM), desc64(a_tap + (uint32_t)(m * 128) + ks * a_kstep, hiw),
                              desc64(b_lo + ks * b_kstep, hiw), idesc, (c | j | ks) != 0 ? 1u : 0u);
                }
              }
            }
            if (ph == 5) umma_commit(bar(kBSlotFree + s));             // the tile of this slot is not read again
            umma_commit(bar(kBAccFull + s));
          }
          if (!G_::Resident) {
            umma_commit(bar(kBWEmpty + st));                            // both slots have issued this convolution
            if (++st == kBWStages) { st = 0; ph_w ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 3) {
    // ------------------------------------------------ epilogue: 8 warps per slot; thread = one window row of its slot's tiles
    const int quad = warp & 3;                       // TMEM lanes [32 * quad, 32 * quad + 32)
    const int s = (warp - 3) >> 3;                   // slot
    const int m = ((warp - 3) >> 2) & 1;             // 128-row sub-tile
    const int r = m * 128 + quad * 32 + lane;        // window row
    const int fmt = p.fmt;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * 2 * NM + m * NM);
    const uint32_t p_row = p_base + (uint32_t)s * G_::TileBytes + (uint32_t)(kBMaxDil + r) * 16u;
    PairCursor cur;
    uint32_t acc_par = 0u;
    float y[C];                                       // fp32 residual stream of this thread's row
    for (int it = s; it < n_it; it += 2) {
      const PairTile tc = pair_decode(pp, pref, lim_s, (uint32_t)(cta + it * G), cur);
      const int t = tc.q0 - H + r;                                     // sequence position of this row
      const bool inside = t >= 0 && t < p.T;
      // the residual of this row (x of the ResBlock): in flight while conv1 of the first pair runs
      if (p.res && inside) {
        const float4* rp = reinterpret_cast<const float4*>(p.res + (size_t)tc.b * p.o32_bs) + t;
#pragma unroll
        for (int q = 0; q < C / 4; ++q) {
          const float4 v = rp[(size_t)q * p.T];
          y[4 * q] = v.x; y[4 * q + 1] = v.y; y[4 * q + 2] = v.z; y[4 * q + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < C; ++q) y[q] = 0.f;
      }
#pragma unroll
      for (int ph = 0; ph < 6; ++ph) {                // unrolled: E1 / E2 / the final stores are straight-line code
        mbar_wait(bar(kBAccFull + s), acc_par); acc_par ^= 1u;
        tc_fence_after();
        const bool e1 = (ph & 1) == 0, last = ph == 5;
        const bool store_out = last && r >= H && r < kPairRows - H && inside && t < tc.lim;
        float* op = (store_out && p.o32) ? p.o32 + (size_t)tc.b * p.o32_bs + (size_t)t * 4 : nullptr;
#pragma unroll
        for (int hh = 0; hh < C / CW; ++hh) {         // CW channels at a time (608 threads: 107 registers each)
          uint32_t a[CW], l[CW];
          __syncwarp();
          if constexpr (CW == 16) {
            tmem_ld16_nowait(lane_addr + (uint32_t)(CW * hh), a);
            tmem_ld16_nowait(lane_addr + (uint32_t)(C + CW * hh), l);
          } else {
            tmem_ld8_nowait(lane_addr + (uint32_t)(CW * hh), a);
            tmem_ld8_nowait(lane_addr + (uint32_t)(C + CW * hh), l);
          }
          tmem_ld_wait();
          float* af = reinterpret_cast<float*>(a);
          const float* lf = reinterpret_cast<const float*>(l);
#pragma unroll
          for (int q = 0; q < CW; q += 2) add2(af[q], af[q + 1], lf[q], lf[q + 1]);       // hi + lo weight plane
          const float4* bv = reinterpret_cast<const float4*>(bias_s + ph * C + CW * hh);
#pragma unroll
          for (int q = 0; q < CW / 4; ++q) {
            const float4 bq = bv[q];
            fma2(af[4 * q], af[4 * q + 1], 1.f, bq.x, bq.y);
            fma2(af[4 * q + 2], af[4 * q + 3], 1.f, bq.z, bq.w);
          }
          if (!e1) {
            // E2: y = conv2 + b2 + y (the pair's residual update, fp32)
#pragma unroll
            for (int q = 0; q < CW; q += 2) add2(af[q], af[q + 1], y[CW * hh + q], y[CW * hh + q + 1]);
#pragma unroll
            for (int q = 0; q < CW; ++q) y[CW * hh + q] = af[q];
          }
          if (!last) {
            // E1: tile = fp16(leaky(conv1 + b1)); E2: tile = the next pair's input fp16(leaky(y)); zero outside the sequence
            const uint32_t dst = p_row + (uint32_t)(CW / 8 * hh) * kBRP * 16u;
#pragma unroll
            for (int sl = 0; sl < CW / 8; ++sl) {
              uint32_t hw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float l0, l1;
                leaky2(af[8 * sl + 2 * e], af[8 * sl + 2 * e + 1], p.slope, l0, l1);
                hw[e] = inside ? pack2(l0, l1, fmt) : 0u;
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)sl * kBRP * 16u), "r"(hw[0]),
                           "r"(hw[1]), "r"(hw[2]), "r"(hw[3])
                           : "memory");
            }
          } else if (store_out) {
            // the block's output (core rows only): post * y (+ out), fp32 stream [C/4][T][4] and the next operand planes
            float* v = af;
            if (p.post != 1.f) {
#pragma unroll
              for (int q = 0; q < CW; q += 2) mul2(v[q], v[q + 1], p.post, p.post);
            }
            if (op) {
              float4* o4 = reinterpret_cast<float4*>(op) + (size_t)(CW / 4 * hh) * p.T;
              if (p.accumulate) {
                float4 old[CW / 4];
#pragma unroll
                for (int q = 0; q < CW / 4; ++q) old[q] = o4[(size_t)q * p.T];
#pragma unroll
                for (int q = 0; q < CW / 4; ++q) {
                  v[4 * q] += old[q].x; v[4 * q + 1] += old[q].y; v[4 * q + 2] += old[q].z; v[4 * q + 3] += old[q].w;
                }
              }
#pragma unroll
              for (int q = 0; q < CW / 4; ++q) o4[(size_t)q * p.T] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
            if (p.o_hi) {
              const size_t prow = (size_t)tc.b * p.op_bs + ((size_t)p.op_pad + t) * 8;
#pragma unroll
              for (int sl = 0; sl < CW / 8; ++sl) {
                uint32_t hw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float l0, l1;
                  leaky2(v[8 * sl + 2 * e], v[8 * sl + 2 * e + 1], p.slope, l0, l1);
                  hw[e] = pack2(l0, l1, fmt);
                }
                *reinterpret_cast<uint4*>(p.o_hi + prow + (size_t)(CW / 8 * hh + sl) * p.op_rows * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              }
            }
          }
        }
        tc_fence_before();
        if (!last) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kBReady + s));                  // operand written / accumulator drained
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

int rb_block_supported(const TcConvW* const c1[3], const TcConvW* const c2[3], const int dil[3], int a_planes) {
  if (a_planes != 1 || !c1[0]) return 0;
  const int C = c1[0]->C_in;
  if (C != 32 && C != 64) return 0;
  int halo = 0;
  for (int i = 0; i < 3; ++i) {
    const TcConvW* cs[2] = {c1[i], c2[i]};
    for (const TcConvW* c : cs) {
      if (!c || c->C_in != C || c->C_out != C || c->ktaps != kBK || !c->stack || c->planes != 1 || c->N != C ||
          c->KC != 32 || c->il_u || c->pair || c->lo8 || c->fmt != c1[0]->fmt)
        return 0;
    }
    if (dil[i] < 1 || dil[i] > kBMaxDil) return 0;
    halo += dil[i] + 1;
  }
  if (halo + kBMaxDil > TC_PADF || 2 * halo >= kPairRows / 2) return 0;
  return (C == 32 ? BlockGeom<32>::Smem : BlockGeom<64>::Smem) <= 227 * 1024;
}

cudaError_t launch_rb_block(RbBlockParams p, cudaStream_t stream) {
  if (p.B <= 0 || p.T <= 0) return cudaSuccess;
  if (p.lens && p.B > TC_MAX_RAGGED_ITEMS) return cudaErrorInvalidValue;
  if (p.C != 32 && p.C != 64) return cudaErrorInvalidValue;
  p.halo = 0;
  for (int i = 0; i < 3; ++i) p.halo += p.dil[i] + 1;
  p.S = kPairRows - 2 * p.halo;
  p.ntiles = cdiv(p.T, p.S);
  // the last tile stages rows up to q0 - halo - 5 + 266 of the input planes: they must exist
  if (p.a_pad - p.halo - kBMaxDil < 0 || p.a_pad + (p.ntiles - 1) * p.S - p.halo - kBMaxDil + kBRP > p.a_rows)
    return cudaErrorInvalidValue;
  static unsigned long long done32 = 0, done64 = 0;
  const cudaError_t attr_err = p.C == 32 ? ensure_max_dyn_smem(rb_block_kernel<32>, 227 * 1024, &done32)
                                         : ensure_max_dyn_smem(rb_block_kernel<64>, 227 * 1024, &done64);
  if (attr_err != cudaSuccess) return attr_err;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  const long tiles = (long)p.ntiles * p.B;
  const int grid = (int)(tiles < sms ? tiles : sms);
  const size_t smem = p.C == 32 ? BlockGeom<32>::Smem : BlockGeom<64>::Smem;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kBThreads);
  cfg.dynamicSmemBytes = smem < 116 * 1024 ? 116 * 1024 : smem;     // the CTA owns all of TMEM: alone on its SM
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = tc_pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return p.C == 32 ? cudaLaunchKernelEx(&cfg, rb_block_kernel<32>, p) : cudaLaunchKernelEx(&cfg, rb_block_kernel<64>, p);
}

}  // namespace dtts
