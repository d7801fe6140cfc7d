// A whole ResBlock1 (three conv1 -> leaky -> conv2 -> + residual pairs, dilations d0, d1, d2) of the last HiFi-GAN stage
// (C = 32) with k = 3 in ONE launch.  The three fused pair launches of rb_pair32_kernel move 3 x 12 B per element through
// HBM and run at its roof (0.86 of the copy peak); here the fp32 residual stream of a row stays in the registers of the
// epilogue thread that owns the row and the operand planes between the pairs stay in shared memory, so the block reads
// 2 + 4 B and writes 4 B per element.  Reference: ResBlock1.forward, modules/hifigan/hifigan.py:51-58.
//
// Tile = a window of 256 rows [t0, t0 + 256), t0 = q0 - H with H = sum(d_p + 1) = 12 the receptive field of the block
// per side: every convolution computes all 256 rows, a row is exact when it is at least (rows of receptive field so
// far) away from a window edge, and only the core rows [H, 256 - H) of the last pair are stored; tiles advance by
// S = 256 - 2 H = 232 rows (10 % recompute).  Rows outside [0, T) are written as zeros into the operand tiles (the zero
// padding each convolution sees).  The six convolutions are the same implicit GEMMs as rb_pair32_kernel (hi | lo weight
// planes stacked along N = 64, K = 32 = two MMA K-steps per tap, the same MMA order, the same roundings), so the result is
// BIT-IDENTICAL to the three pair launches -- and therefore to the six tc_conv launches -- it replaces.
//
// Two tiles are in flight per CTA (slots 0 / 1), each walks its six phases conv1(0) conv2(0) conv1(1) ... in order; the
// MMA thread alternates between the slots and the eight epilogue warps follow it, so the tensor pipe works on one tile
// while the epilogue turns the other tile's accumulators into the next operand:
//   warp 0       all weights once (6 x 3 taps x 4 KB, resident); the input window of the next tile of a slot by bulk copy
//   warp 2       TMEM (one 2 x 64-column accumulator per slot); one thread issues the MMAs
//   warps 3-18   two groups of eight, one per slot; thread = window row: E1 (bias, leaky, fp16 -> T tile), E2 (bias,
//                + residual register; leaky, fp16 -> P tile for the next pair, or the fp32 stream / output planes after
//                the last pair).  With ONE group serving both slots the kernel was epilogue bound (0.75 ms: E(A) then
//                E(B) per phase against 2 x 600 cycles of MMAs)
#include "tc_conv.cuh"
#include "tc16.cuh"
#include "tc_ptx.cuh"
#include "rb_pair_common.cuh"

#include <mutex>

namespace dtts {

namespace {

constexpr int kBC = 32, kBNM = 64;                 // channels; MMA N (hi | lo stacked)
constexpr int kBK = 3;                             // kernel size
constexpr int kBTapBytes = kBNM * kBC * 2;         // 4 KB
constexpr int kBConvBytes = kBK * kBTapBytes;      // one convolution's weights: 12 KB
constexpr int kBMaxDil = 5;                        // margin rows of the P tile on both sides
constexpr int kBRP = kPairRows + 2 * kBMaxDil;     // rows of a P tile (window row w at buffer row w + kBMaxDil)
constexpr int kBRT = kPairRows + 2;                // rows of a T tile (window row w at buffer row w + 1)
constexpr int kBPBytes = (kBC / 8) * kBRP * 16, kBTBytes = (kBC / 8) * kBRT * 16;
// mbarriers, per slot: input landed / slot may be re-filled / accumulator full / operand written (or accumulator drained)
constexpr int kBInFull = 0, kBSlotFree = 2, kBAccFull = 4, kBReady = 6, kBWFull = 8, kBNumBars = 9;
constexpr int kBTmemOff = kBNumBars * 8;
constexpr int kBBiasOff = 128;                                            // 6 x 32 floats
constexpr int kBPrefOff = kBBiasOff + 6 * kBC * 4;
constexpr int kBHeader = (kBPrefOff + (2 * TC_MAX_RAGGED_ITEMS + 8) * 4 + 127) / 128 * 128;

constexpr int kBThreads = 96 + 16 * 32;            // warps 0-2: producer / - / MMA; warps 3-18: two epilogue groups, one per slot

__global__ void __launch_bounds__(kBThreads, 1) rb_block32_kernel(const RbBlockParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.halo;
  const uint32_t bar0 = smem_u32(smem);
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + kBTmemOff);
  float* bias_s = reinterpret_cast<float*>(smem + kBBiasOff);
  int* pref_s = reinterpret_cast<int*>(smem + kBPrefOff);
  int* lim_s = pref_s + TC_MAX_RAGGED_ITEMS + 1;
  const uint32_t w_base = smem_u32(smem + kBHeader);                     // conv1(0) conv2(0) conv1(1) ... 12 KB each
  const uint32_t p_base = w_base + 6u * kBConvBytes;                     // P tile of slot 0, slot 1
  const uint32_t t_base = p_base + 2u * kBPBytes;                        // T tile of slot 0, slot 1

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
    mbar_init(bar(kBWFull), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem + kBTmemOff)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 96 && threadIdx.x < 96 + 6 * kBC) {
    const int i = threadIdx.x - 96, c = i / kBC;                        // conv c: bias of pair c / 2, conv1 or conv2
    bias_s[i] = __ldg(p.bias[c] + (i - c * kBC));
  }
  // the T tiles' margin rows are never written: zero the T tiles once (their content only ever reaches rows that are
  // not stored, but it should be finite)
  for (uint32_t i = threadIdx.x; i < 2u * kBTBytes / 16u; i += kBThreads)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(t_base + i * 16u), "r"(0u) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0 && elect_one()) {                    // constants: may be fetched before the predecessor grid has completed
    mbar_arrive_expect_tx(bar(kBWFull), 6u * kBConvBytes);
    for (int c = 0; c < 6; ++c) bulk_g2s(w_base + (uint32_t)c * kBConvBytes, p.w[c], kBConvBytes, bar(kBWFull));
  }
  griddep_wait();
  const uint32_t tmem_base = *tmem_ptr_s;
  const int* pref = p.lens ? pref_s : nullptr;
  const int nrt = p.lens ? pref_s[p.B] : p.ntiles * p.B;
  const int G = gridDim.x, cta = blockIdx.x;
  const int n_it = nrt > cta ? (nrt - cta + G - 1) / G : 0;           // tiles of this CTA: cta, cta + G, ...; tile i -> slot i & 1

  if (warp == 0) {
    // ------------------------------------------------ producer: the input window (+ margins) of tile i into P[i & 1]
    __syncwarp();
    PairCursor cur;
    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1;
      const PairTile tc = pair_decode(pp, pref, lim_s, (uint32_t)(cta + it * G), cur);
      mbar_wait(bar(kBSlotFree + s), ((it >> 1) & 1) ^ 1);             // the previous tile of this slot has issued its last MMAs
      if (elect_one()) {
        mbar_arrive_expect_tx(bar(kBInFull + s), kBPBytes);
        const size_t row0 = (size_t)(p.a_pad + tc.q0 - H - kBMaxDil);
        const tc16* src = p.a_hi + (size_t)tc.b * p.a_bs;
        for (int sl = 0; sl < kBC / 8; ++sl)
          bulk_g2s(p_base + (uint32_t)s * kBPBytes + (uint32_t)sl * kBRP * 16u, src + ((size_t)sl * p.a_rows + row0) * 8,
                   (uint32_t)kBRP * 16u, bar(kBInFull + s));
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ------------------------------------------------ MMA issuer: phase ph of slot 0, phase ph of slot 1, ph = 0 .. 5
    if (elect_one()) {
      const uint32_t hiw = (128u >> 4) | (1u << 14);                  // SBO = 128 B, descriptor version 1
      const uint32_t f16b = p.fmt ? 1u : 0u;
      const uint32_t idesc = (1u << 4) | (f16b << 7) | (f16b << 10) | ((uint32_t)(kBNM >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t p_low0 = ((p_base >> 4) & 0x3FFFu) | ((uint32_t)kBRP << 16);
      const uint32_t t_low0 = ((t_base >> 4) & 0x3FFFu) | ((uint32_t)kBRT << 16);
      const uint32_t w_low0 = ((w_base >> 4) & 0x3FFFu) | ((uint32_t)kBNM << 16);
      const uint32_t p_kstep = 2u * kBRP, t_kstep = 2u * kBRT, b_kstep = 2u * kBNM;
      mbar_wait(bar(kBWFull), 0);
      uint32_t rdy_par[2] = {0u, 0u};                                   // phase parity of kBReady per slot
      for (int i0 = 0; i0 < n_it; i0 += 2) {
        const int ns = n_it - i0 >= 2 ? 2 : 1;
        for (int ph = 0; ph < 6; ++ph) {
          for (int s = 0; s < ns; ++s) {
            const int it = i0 + s;
            if (ph == 0) {
              mbar_wait(bar(kBInFull + s), (it >> 1) & 1);
              if (it >= 2) { mbar_wait(bar(kBReady + s), rdy_par[s]); rdy_par[s] ^= 1u; }   // accumulator drained
            } else {
              mbar_wait(bar(kBReady + s), rdy_par[s]); rdy_par[s] ^= 1u;                     // operand tile written
            }
            tc_fence_after();
            const uint32_t d_base = tmem_base + (uint32_t)(s * 128);
            const uint32_t b_lo0 = w_low0 + (uint32_t)ph * (kBConvBytes >> 4);
            uint32_t a_tap, a_step, a_kstep;
            if ((ph & 1) == 0) {                                        // conv1 of pair ph / 2: taps at -d, 0, +d around the row
              const uint32_t d = (uint32_t)p.dil[ph >> 1];
              a_tap = p_low0 + (uint32_t)s * (kBPBytes >> 4) + (uint32_t)kBMaxDil - d;
              a_step = d; a_kstep = p_kstep;
            } else {                                                    // conv2: taps at -1, 0, +1
              a_tap = t_low0 + (uint32_t)s * (kBTBytes >> 4);
              a_step = 1u; a_kstep = t_kstep;
            }
            uint32_t b_lo = b_lo0;
            for (int j = 0; j < kBK; ++j, a_tap += a_step, b_lo += (kBTapBytes >> 4)) {
#pragma unroll
              for (int m = 0; m < 2; ++m) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                  umma_bf16(d_base + (uint32_t)(m * kBNM), desc64(a_tap + (uint32_t)(m * 128) + ks * a_kstep, hiw),
                            desc64(b_lo + ks * b_kstep, hiw), idesc, (j | ks) != 0 ? 1u : 0u);
              }
            }
            if (ph == 5) umma_commit(bar(kBSlotFree + s));             // P / T of this slot are not read again
            umma_commit(bar(kBAccFull + s));
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
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * 128 + m * kBNM);
    const uint32_t t_row = t_base + (uint32_t)s * kBTBytes + (uint32_t)(1 + r) * 16u;
    const uint32_t p_row = p_base + (uint32_t)s * kBPBytes + (uint32_t)(kBMaxDil + r) * 16u;
    PairCursor cur;
    uint32_t acc_par = 0u;
    float y[kBC];                                     // fp32 residual stream of this thread's row
    for (int it = s; it < n_it; it += 2) {
      const PairTile tc = pair_decode(pp, pref, lim_s, (uint32_t)(cta + it * G), cur);
      const int t = tc.q0 - H + r;                                     // sequence position of this row
      const bool inside = t >= 0 && t < p.T;
      // the residual of this row (x of the ResBlock): in flight while conv1 of the first pair runs
      if (p.res && inside) {
        const float4* rp = reinterpret_cast<const float4*>(p.res + (size_t)tc.b * p.o32_bs) + t;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = rp[(size_t)q * p.T];
          y[4 * q] = v.x; y[4 * q + 1] = v.y; y[4 * q + 2] = v.z; y[4 * q + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < kBC; ++q) y[q] = 0.f;
      }
#pragma unroll
      for (int ph = 0; ph < 6; ++ph) {                // unrolled: E1 / E2 / the final stores are straight-line code
        mbar_wait(bar(kBAccFull + s), acc_par); acc_par ^= 1u;
        tc_fence_after();
        const bool e1 = (ph & 1) == 0, last = ph == 5;
        const bool store_out = last && r >= H && r < kPairRows - H && inside && t < tc.lim;
        float* op = (store_out && p.o32) ? p.o32 + (size_t)tc.b * p.o32_bs + (size_t)t * 4 : nullptr;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {              // 16 channels at a time (608 threads: 107 registers each)
          uint32_t a[16], l[16];
          __syncwarp();
          tmem_ld16_nowait(lane_addr + (uint32_t)(16 * hh), a);
          tmem_ld16_nowait(lane_addr + (uint32_t)(kBC + 16 * hh), l);
          tmem_ld_wait();
          float* af = reinterpret_cast<float*>(a);
          const float* lf = reinterpret_cast<const float*>(l);
#pragma unroll
          for (int q = 0; q < 16; q += 2) add2(af[q], af[q + 1], lf[q], lf[q + 1]);       // hi + lo weight plane
          const float4* bv = reinterpret_cast<const float4*>(bias_s + ph * kBC + 16 * hh);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = bv[q];
            fma2(af[4 * q], af[4 * q + 1], 1.f, bq.x, bq.y);
            fma2(af[4 * q + 2], af[4 * q + 3], 1.f, bq.z, bq.w);
          }
          if (!e1) {
            // E2: y = conv2 + b2 + y (the pair's residual update, fp32)
#pragma unroll
            for (int q = 0; q < 16; q += 2) add2(af[q], af[q + 1], y[16 * hh + q], y[16 * hh + q + 1]);
#pragma unroll
            for (int q = 0; q < 16; ++q) y[16 * hh + q] = af[q];
          }
          if (!last) {
            // E1: T tile = fp16(leaky(conv1 + b1)); E2: P tile of the next pair = fp16(leaky(y)); zero outside the sequence
            const uint32_t dst = (e1 ? t_row : p_row) + (uint32_t)(2 * hh) * (uint32_t)(e1 ? kBRT : kBRP) * 16u;
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
              uint32_t hw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float l0, l1;
                leaky2(af[8 * sl + 2 * e], af[8 * sl + 2 * e + 1], p.slope, l0, l1);
                hw[e] = inside ? pack2(l0, l1, fmt) : 0u;
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)sl * (uint32_t)(e1 ? kBRT : kBRP) * 16u),
                           "r"(hw[0]), "r"(hw[1]), "r"(hw[2]), "r"(hw[3])
                           : "memory");
            }
          } else if (store_out) {
            // the block's output (core rows only): post * y (+ out), fp32 stream [C/4][T][4] and the next operand planes
            float* v = af;
            if (p.post != 1.f) {
#pragma unroll
              for (int q = 0; q < 16; q += 2) mul2(v[q], v[q + 1], p.post, p.post);
            }
            if (op) {
              float4* o4 = reinterpret_cast<float4*>(op) + (size_t)(4 * hh) * p.T;
              if (p.accumulate) {
                float4 old[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) old[q] = o4[(size_t)q * p.T];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  v[4 * q] += old[q].x; v[4 * q + 1] += old[q].y; v[4 * q + 2] += old[q].z; v[4 * q + 3] += old[q].w;
                }
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) o4[(size_t)q * p.T] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
            if (p.o_hi) {
              const size_t prow = (size_t)tc.b * p.op_bs + ((size_t)p.op_pad + t) * 8;
#pragma unroll
              for (int sl = 0; sl < 2; ++sl) {
                uint32_t hw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float l0, l1;
                  leaky2(v[8 * sl + 2 * e], v[8 * sl + 2 * e + 1], p.slope, l0, l1);
                  hw[e] = pack2(l0, l1, fmt);
                }
                *reinterpret_cast<uint4*>(p.o_hi + prow + (size_t)(2 * hh + sl) * p.op_rows * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
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

size_t block32_smem_bytes() { return (size_t)kBHeader + 6 * kBConvBytes + 2 * (size_t)kBPBytes + 2 * (size_t)kBTBytes; }

}  // namespace

int rb_block32_supported(const TcConvW* const c1[3], const TcConvW* const c2[3], const int dil[3], int a_planes) {
  if (a_planes != 1) return 0;
  int halo = 0;
  for (int i = 0; i < 3; ++i) {
    const TcConvW* cs[2] = {c1[i], c2[i]};
    for (const TcConvW* c : cs) {
      if (!c || c->C_in != kBC || c->C_out != kBC || c->ktaps != kBK || !c->stack || c->planes != 1 || c->N != kBC ||
          c->KC != 32 || c->il_u || c->pair || c->lo8 || c->fmt != c1[0]->fmt)
        return 0;
    }
    if (dil[i] < 1 || dil[i] > kBMaxDil) return 0;
    halo += dil[i] + 1;
  }
  if (halo + kBMaxDil > TC_PADF || 2 * halo >= kPairRows / 2) return 0;
  return block32_smem_bytes() <= (size_t)227 * 1024;
}

cudaError_t launch_rb_block32(RbBlockParams p, cudaStream_t stream) {
  if (p.B <= 0 || p.T <= 0) return cudaSuccess;
  if (p.lens && p.B > TC_MAX_RAGGED_ITEMS) return cudaErrorInvalidValue;
  p.halo = 0;
  for (int i = 0; i < 3; ++i) p.halo += p.dil[i] + 1;
  p.S = kPairRows - 2 * p.halo;
  p.ntiles = cdiv(p.T, p.S);
  // the last tile stages rows up to q0 - halo - 5 + 266 of the input planes: they must exist
  if (p.a_pad - p.halo - kBMaxDil < 0 || p.a_pad + (p.ntiles - 1) * p.S - p.halo - kBMaxDil + kBRP > p.a_rows)
    return cudaErrorInvalidValue;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(rb_block32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) return attr_err;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  const long tiles = (long)p.ntiles * p.B;
  const int grid = (int)(tiles < sms ? tiles : sms);
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kBThreads);
  cfg.dynamicSmemBytes = block32_smem_bytes();     // > half of the SM's shared memory: the CTA (all of TMEM) is alone on its SM
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = tc_pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, rb_block32_kernel, p);
}

}  // namespace dtts
