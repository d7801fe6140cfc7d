// Fused ResBlock convolution pair for the narrow HiFi-GAN stages (C = 32, the last stage): conv1 (dilated) -> leaky-ReLU ->
// conv2 (dilation 1) -> + residual in ONE launch; the intermediate activation never leaves the SM.
//
//   t1[t]  = fp16(leaky(conv1(x)[t] + b1))           (zero outside [0, T): conv2's zero padding)
//   out[t] = post * (conv2(t1)[t] + b2 + res[t])     (+ out[t] if accumulate), planes_out = fp16(leaky(out))
//
// Unfused (tc_conv.cu) the pair moves 16 B per element through HBM (2 in + 2 out | 2 in + 4 residual + 4 + 2 out) and its
// second launch runs at the HBM roof; fused it moves 12 B and the 2 + 2 B of the intermediate become shared-memory
// traffic.  Reference: ResBlock1.forward, modules/hifigan/hifigan.py:51-58.
//
// Both convolutions are the same implicit GEMM as tc_conv_kernel (M = 128 time rows per MMA, hi | lo weight planes
// stacked along N = 64, K = 32 channels = 2 MMA K-steps per tap, fp32 accumulators in TMEM) and issue their MMAs in the
// same order, and the intermediate is rounded exactly like the operand planes the unfused pair writes -- so the fused
// result is BIT-IDENTICAL to the two launches it replaces (tests/test_gpu_tensorcore.py).  The output planes must be a
// different buffer than the input planes (a tile stages its neighbours' rows as halo); the fp32 stream may be updated in
// place (rows map one to one).
//
// Tile = 256 rows [t0, t0 + 256), t0 = q0 - h2 (h2 = (k-1)/2): conv1 produces all 256 rows into a shared-memory tile in
// the UMMA K-major layout, conv2 reads its taps from that tile by moving the descriptor start address; its rows
// [h2, 256 - h2) are complete, so tiles advance by S = 256 - 2*h2 rows (<= 4 % recompute).  All weights of both
// convolutions (2 * k * 4 KB) stay resident in shared memory for the life of the persistent CTA.
// Pipeline (one CTA per SM, 352 threads):
//   warp 0    input tiles (tile + conv1 halo) by cp.async.bulk, 3 stages; the weights once
//   warp 2    TMEM: 2 x conv1 accumulator sets + 2 x conv2 sets (4 x 128 columns); one thread issues
//             M1(i) then M2(i-1): conv1 of the next tile fills the tensor pipe while the epilogue warps turn tile i's
//             conv1 accumulators into the conv2 operand
//   warps 3-10  E1(i): TMEM -> bias, leaky, fp16 -> shared tile (fence.proxy.async) ; E2(i-1): TMEM -> bias, residual,
//             1/3 mean, fp32 stream + next layer's operand planes
#include "tc_conv.cuh"
#include "tc16.cuh"
#include "tc_ptx.cuh"
#include "rb_pair_common.cuh"

#include <mutex>

namespace dtts {

namespace {

constexpr int kPairC = 32, kPairNM = 64;       // channels; MMA N (hi | lo stacked)
constexpr int kPairAStages = 3;
constexpr int kTapBytes = kPairNM * kPairC * 2;   // one tap of one convolution: 4 KB
// mbarriers
constexpr int kPAFull = 0, kPAEmpty = kPAFull + kPairAStages, kPWFull = kPAEmpty + kPairAStages, kPAcc1Full = kPWFull + 1,
              kPAcc1Empty = kPAcc1Full + 2, kPAcc2Full = kPAcc1Empty + 2, kPAcc2Empty = kPAcc2Full + 2,
              kPTFull = kPAcc2Empty + 2, kPTEmpty = kPTFull + 2, kPNumBars = kPTEmpty + 2;
constexpr int kPTmemOff = kPNumBars * 8;
constexpr int kPBiasOff = (kPTmemOff + 4 + 63) / 64 * 64;                 // b1[32], b2[32], conv_post weights [32][7]
constexpr int kPPrefOff = kPBiasOff + (2 * kPairC + kPairC * 7) * 4;       // ragged tables
constexpr int kPHeader = (kPPrefOff + (2 * TC_MAX_RAGGED_ITEMS + 8) * 4 + 127) / 128 * 128;

__global__ void __launch_bounds__(kPairThreads, 1) rb_pair32_kernel(const RbPairParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = p.k, h2 = (p.k - 1) / 2, hd = h2 * p.dil;
  const int RA = kPairRows + 2 * hd;           // staged input rows per tile
  const int RT = kPairRows + 2 * h2;           // rows of the intermediate tile (h2 margin rows on both sides)
  const uint32_t bar0 = smem_u32(smem);
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + kPTmemOff);
  float* bias_s = reinterpret_cast<float*>(smem + kPBiasOff);
  int* pref_s = reinterpret_cast<int*>(smem + kPPrefOff);
  int* lim_s = pref_s + TC_MAX_RAGGED_ITEMS + 1;
  const uint32_t w_bytes = (uint32_t)k * kTapBytes;                  // one convolution's weights
  const uint32_t a_stage_bytes = (uint32_t)(kPairC / 8) * RA * 16u;
  const uint32_t t_buf_bytes = (uint32_t)(kPairC / 8) * RT * 16u;
  const uint32_t w1_base = smem_u32(smem + kPHeader);
  const uint32_t w2_base = w1_base + w_bytes;
  const uint32_t a_base = w2_base + w_bytes;
  const uint32_t t_base = a_base + kPairAStages * a_stage_bytes;

  griddep_launch();                            // programmatic dependent launch, as tc_conv_kernel (common.cuh)
  if (p.lens && warp == 3) {                   // per-item row limits and the exclusive prefix of their tile counts
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
    for (int s = 0; s < kPairAStages; ++s) { mbar_init(bar(kPAFull + s), 1); mbar_init(bar(kPAEmpty + s), 1); }
    mbar_init(bar(kPWFull), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(kPAcc1Full + s), 1); mbar_init(bar(kPAcc1Empty + s), 8);
      mbar_init(bar(kPAcc2Full + s), 1); mbar_init(bar(kPAcc2Empty + s), 8);
      mbar_init(bar(kPTFull + s), 8);    mbar_init(bar(kPTEmpty + s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem + kPTmemOff)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 96 && threadIdx.x < 96 + 2 * kPairC) {
    const int i = threadIdx.x - 96;
    bias_s[i] = i < kPairC ? __ldg(p.b1 + i) : __ldg(p.b2 + i - kPairC);
  }
  if (p.post_part && threadIdx.x >= 96 && threadIdx.x < 96 + kPairC * 7) bias_s[2 * kPairC + threadIdx.x - 96] = __ldg(p.post_w + threadIdx.x - 96);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // everything above overlaps the tail of the previous launch; the weights (constants) are fetched by warp 0's first
  // bulk copies right below, activations / residuals / outputs only after the predecessor grid has completed
  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(bar(kPWFull), 2u * w_bytes);
    bulk_g2s(w1_base, p.w1, w_bytes, bar(kPWFull));
    bulk_g2s(w2_base, p.w2, w_bytes, bar(kPWFull));
  }
  griddep_wait();
  const uint32_t tmem_base = *tmem_ptr_s;
  const int* pref = p.lens ? pref_s : nullptr;
  const int nrt = p.lens ? pref_s[p.B] : p.ntiles * p.B;
  const int G = gridDim.x, cta = blockIdx.x;
  const int n_it = nrt > cta ? (nrt - cta + G - 1) / G : 0;           // tiles of this CTA: cta, cta + G, ...

  if (warp == 0) {
    // ------------------------------------------------ producer: the input tiles (the weights were requested above)
    __syncwarp();
    int s = 0;
    uint32_t ph = 1;
    PairCursor cur;
    for (int it = 0; it < n_it; ++it) {
      const PairTile tc = pair_decode(p, pref, lim_s, (uint32_t)(cta + it * G), cur);
      mbar_wait(bar(kPAEmpty + s), ph);
      if (elect_one()) {
        mbar_arrive_expect_tx(bar(kPAFull + s), a_stage_bytes);
        const size_t row0 = (size_t)(p.a_pad + tc.q0 - h2 - hd);
        const tc16* src = p.a_hi + (size_t)tc.b * p.a_bs;
        for (int sl = 0; sl < kPairC / 8; ++sl)
          bulk_g2s(a_base + s * a_stage_bytes + sl * (uint32_t)RA * 16u, src + ((size_t)sl * p.a_rows + row0) * 8,
                   (uint32_t)RA * 16u, bar(kPAFull + s));
      }
      __syncwarp();
      if (++s == kPairAStages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 2) {
    // ------------------------------------------------ MMA issuer: M1(i), then M2(i - 1)
    if (elect_one()) {
      const uint32_t hiw = (128u >> 4) | (1u << 14);                  // SBO = 128 B, descriptor version 1
      const uint32_t f16b = p.fmt ? 1u : 0u;
      const uint32_t idesc = (1u << 4) | (f16b << 7) | (f16b << 10) | ((uint32_t)(kPairNM >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t a_low0 = ((a_base >> 4) & 0x3FFFu) | ((uint32_t)RA << 16);
      const uint32_t t_low0 = ((t_base >> 4) & 0x3FFFu) | ((uint32_t)RT << 16);
      const uint32_t w1_low0 = ((w1_base >> 4) & 0x3FFFu) | ((uint32_t)kPairNM << 16);
      const uint32_t w2_low0 = ((w2_base >> 4) & 0x3FFFu) | ((uint32_t)kPairNM << 16);
      const uint32_t a_kstep = 2u * (uint32_t)RA, t_kstep = 2u * (uint32_t)RT, b_kstep = 2u * kPairNM;
      const uint32_t w_tap16 = kTapBytes >> 4, a_stage16 = a_stage_bytes >> 4, t_buf16 = t_buf_bytes >> 4;
      mbar_wait(bar(kPWFull), 0);
      int sa = 0;
      uint32_t pa = 0;
      for (int i = 0; i <= n_it; ++i) {
        if (i < n_it) {                                               // conv1 of tile i
          const int as = i & 1;
          mbar_wait(bar(kPAFull + sa), pa);
          mbar_wait(bar(kPAcc1Empty + as), ((i >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_base = tmem_base + (uint32_t)(as * 128);
          uint32_t a_tap = a_low0 + (uint32_t)sa * a_stage16, b_lo = w1_low0;
          for (int j = 0; j < k; ++j, a_tap += (uint32_t)p.dil, b_lo += w_tap16) {
#pragma unroll
            for (int m = 0; m < 2; ++m) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                umma_bf16(d_base + (uint32_t)(m * kPairNM), desc64(a_tap + (uint32_t)(m * 128) + ks * a_kstep, hiw),
                          desc64(b_lo + ks * b_kstep, hiw), idesc, (j | ks) != 0 ? 1u : 0u);
            }
          }
          umma_commit(bar(kPAEmpty + sa));
          umma_commit(bar(kPAcc1Full + as));
          if (++sa == kPairAStages) { sa = 0; pa ^= 1u; }
        }
        if (i >= 1) {                                                 // conv2 of tile i - 1
          const int t = i - 1, ts = t & 1;
          mbar_wait(bar(kPTFull + ts), (t >> 1) & 1);
          mbar_wait(bar(kPAcc2Empty + ts), ((t >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_base = tmem_base + 256u + (uint32_t)(ts * 128);
          uint32_t a_tap = t_low0 + (uint32_t)ts * t_buf16, b_lo = w2_low0;
          for (int j = 0; j < k; ++j, a_tap += 1u, b_lo += w_tap16) {
#pragma unroll
            for (int m = 0; m < 2; ++m) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                umma_bf16(d_base + (uint32_t)(m * kPairNM), desc64(a_tap + (uint32_t)(m * 128) + ks * t_kstep, hiw),
                          desc64(b_lo + ks * b_kstep, hiw), idesc, (j | ks) != 0 ? 1u : 0u);
            }
          }
          umma_commit(bar(kPTEmpty + ts));
          umma_commit(bar(kPAcc2Full + ts));
        }
      }
    }
    __syncwarp();
  } else if (warp >= 3) {
    // ------------------------------------------------ epilogue: E1(i), then E2(i - 1); one 128-row sub-tile quadrant per warp
    const int quad = warp & 3;                       // TMEM lanes [32*quad, 32*quad + 32)
    const int m = (warp - 3) >> 2;                   // sub-tile
    const int r = m * 128 + quad * 32 + lane;        // row inside the tile
    const int fmt = p.fmt;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(m * kPairNM);
    const float4* b1v = reinterpret_cast<const float4*>(bias_s);
    const float4* b2v = reinterpret_cast<const float4*>(bias_s + kPairC);
    PairCursor cur;
    PairTile prev{0, 0, 0};
    for (int i = 0; i <= n_it; ++i) {
      // residual of tile i - 1 (consumed by E2 below): in flight while E1 runs
      float4 rc[8];
      bool ok2 = false;
      int t2 = 0;
      if (i >= 1) {
        t2 = prev.q0 - h2 + r;
        ok2 = r >= h2 && r < kPairRows - h2 && t2 < prev.lim;
        if (p.res && ok2) {
          const float4* rp = reinterpret_cast<const float4*>(p.res + (size_t)prev.b * p.o32_bs) + t2;
#pragma unroll
          for (int q = 0; q < 8; ++q) rc[q] = rp[(size_t)q * p.T];
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) rc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      PairTile tc = prev;
      if (i < n_it) {
        tc = pair_decode(p, pref, lim_s, (uint32_t)(cta + i * G), cur);
        const int as = i & 1;
        mbar_wait(bar(kPAcc1Full + as), (i >> 1) & 1);
        mbar_wait(bar(kPTEmpty + as), ((i >> 1) & 1) ^ 1);             // conv2 of tile i - 2 has read this buffer
        tc_fence_after();
        uint32_t a[32], l[32];
        __syncwarp();
        tmem_ld32_nowait(lane_addr + (uint32_t)(as * 128), a);
        tmem_ld32_nowait(lane_addr + (uint32_t)(as * 128 + kPairC), l);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kPAcc1Empty + as));             // accumulators are in registers
        float* af = reinterpret_cast<float*>(a);
        const float* lf = reinterpret_cast<const float*>(l);
#pragma unroll
        for (int q = 0; q < 32; q += 2) add2(af[q], af[q + 1], lf[q], lf[q + 1]);     // hi + lo weight plane
        const int t = tc.q0 - h2 + r;
        const bool inside = t >= 0 && t < p.T;                         // conv2 sees zeros outside the sequence
        const uint32_t dst = t_base + (uint32_t)as * t_buf_bytes + (uint32_t)(h2 + r) * 16u;
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
          uint32_t hw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 bq = b1v[2 * sl + (e >> 1)];
            float v0 = af[8 * sl + 2 * e], v1 = af[8 * sl + 2 * e + 1];
            fma2(v0, v1, 1.f, (e & 1) ? bq.z : bq.x, (e & 1) ? bq.w : bq.y);
            float l0, l1;
            leaky2(v0, v1, p.slope, l0, l1);
            hw[e] = inside ? pack2(l0, l1, fmt) : 0u;
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)sl * (uint32_t)RT * 16u), "r"(hw[0]),
                       "r"(hw[1]), "r"(hw[2]), "r"(hw[3])
                       : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kPTFull + as));
      }
      if (i >= 1) {
        const int t = i - 1, ts = t & 1;
        mbar_wait(bar(kPAcc2Full + ts), (t >> 1) & 1);
        tc_fence_after();
        uint32_t a[32], l[32];
        __syncwarp();
        tmem_ld32_nowait(lane_addr + 256u + (uint32_t)(ts * 128), a);
        tmem_ld32_nowait(lane_addr + 256u + (uint32_t)(ts * 128 + kPairC), l);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kPAcc2Empty + ts));
        if (ok2) {
          float* v = reinterpret_cast<float*>(a);
          const float* lf = reinterpret_cast<const float*>(l);
#pragma unroll
          for (int q = 0; q < 32; q += 2) add2(v[q], v[q + 1], lf[q], lf[q + 1]);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bq = b2v[q];
            fma2(v[4 * q], v[4 * q + 1], 1.f, bq.x, bq.y);
            fma2(v[4 * q + 2], v[4 * q + 3], 1.f, bq.z, bq.w);
          }
          if (p.res) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              add2(v[4 * q], v[4 * q + 1], rc[q].x, rc[q].y);
              add2(v[4 * q + 2], v[4 * q + 3], rc[q].z, rc[q].w);
            }
          }
          if (p.post != 1.f) {
#pragma unroll
            for (int q = 0; q < 32; q += 2) mul2(v[q], v[q + 1], p.post, p.post);
          }
          if (p.o32) {
            float4* op = reinterpret_cast<float4*>(p.o32 + (size_t)prev.b * p.o32_bs) + t2;
            if (p.accumulate) {
              float4 old[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) old[q] = op[(size_t)q * p.T];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                v[4 * q] += old[q].x; v[4 * q + 1] += old[q].y; v[4 * q + 2] += old[q].z; v[4 * q + 3] += old[q].w;
              }
            }
            if (!p.post_part) {
#pragma unroll
              for (int q = 0; q < 8; ++q) op[(size_t)q * p.T] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
          }
          if (p.post_part) {
            // conv_post folded in: the seven per-tap partial dot products of this row (tc_conv_post_finish adds the
            // shifted partials and applies tanh); the final fp32 stream is never written
            const float* pw = bias_s + 2 * kPairC;
            float pj[7];
#pragma unroll
            for (int j = 0; j < 7; ++j) pj[j] = 0.f;
#pragma unroll
            for (int c = 0; c < kPairC; ++c) {
              const float lv = fmaxf(v[c], v[c] * p.post_slope);
#pragma unroll
              for (int j = 0; j < 7; ++j) pj[j] = fmaf(pw[c * 7 + j], lv, pj[j]);
            }
            float* pp = p.post_part + (size_t)prev.b * 7 * p.T + t2;
#pragma unroll
            for (int j = 0; j < 7; ++j) pp[(size_t)j * p.T] = pj[j];
          }
          if (p.o_hi) {
            const size_t prow = (size_t)prev.b * p.op_bs + ((size_t)p.op_pad + t2) * 8;
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
              uint32_t hw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float l0, l1;
                leaky2(v[8 * sl + 2 * e], v[8 * sl + 2 * e + 1], p.slope, l0, l1);
                hw[e] = pack2(l0, l1, fmt);
              }
              *reinterpret_cast<uint4*>(p.o_hi + prow + (size_t)sl * p.op_rows * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            }
          }
        }
      }
      prev = tc;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The same fusion for C = 64 (the third HiFi-GAN stage), where the unfused pair is HBM bound too (16 B per element
// against a tensor floor of half its time) but the weights of one convolution (k * 16 KB) no longer fit in shared memory
// next to the tiles: they are STREAMED per tile through a ring of ~32 KB stages by warp 1, in exactly the order the MMA
// thread consumes them -- conv1(0) | conv2(0) conv1(1) | conv2(1) conv1(2) | ... -- as tc_conv_kernel streams them per
// tile today (all CTAs read the same blobs, they stay in L2).  K = 64 = two 32-channel chunks: conv1 waits for one
// staged input chunk at a time, conv2 reads both chunks of the intermediate tile.  One accumulator set per
// convolution (2 x 128 columns x 2 sub-tiles = all 512 TMEM columns) but TWO intermediate tiles, and the issue order
//     M1(0) | M1(1) M2(0) | M1(2) M2(1) | ...        epilogue:  E1(0) | E1(1) E2(0) | E1(2) E2(1) | ...
// so that E1(i+1) (conv1 accumulators -> operand tile) runs under M2(i) and E2(i) (stores) under M1(i+2): with the
// first, simpler order M2(i) M1(i+1) every tile paid E1 on the critical path (k = 11: 11 us per tile for 7 us of MMAs).
// Same MMA order (chunk, tap, sub-tile, K step) and same rounding as the two tc_conv launches: bit-identical.
constexpr int kP64C = 64, kP64NM = 128, kP64KC = 32;
constexpr int kP64TapBytes = kP64NM * kP64KC * 2;               // one (chunk, tap) blob: 8 KB
constexpr int kP64AStages = 3, kP64WStages = 3, kP64MaxTG = 4;   // (p.a_stages <= kP64AStages input stages are used)
constexpr int kQAFull = 0, kQAEmpty = kQAFull + kP64AStages, kQWFull = kQAEmpty + kP64AStages,
              kQWEmpty = kQWFull + kP64WStages, kQAcc1Full = kQWEmpty + kP64WStages, kQAcc1Empty = kQAcc1Full + 1,
              kQAcc2Full = kQAcc1Empty + 1, kQAcc2Empty = kQAcc2Full + 1, kQTFull = kQAcc2Empty + 1,
              kQTEmpty = kQTFull + 2, kQNumBars = kQTEmpty + 2;
constexpr int kQTmemOff = kQNumBars * 8;
constexpr int kQBiasOff = (kQTmemOff + 4 + 63) / 64 * 64;                 // b1[64], b2[64]
constexpr int kQPrefOff = kQBiasOff + 2 * kP64C * 4;
constexpr int kQHeader = (kQPrefOff + (2 * TC_MAX_RAGGED_ITEMS + 8) * 4 + 127) / 128 * 128;

__global__ void __launch_bounds__(kPairThreads, 1) rb_pair64_kernel(const RbPairParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = p.k, h2 = (p.k - 1) / 2, hd = h2 * p.dil;
  const int RA = kPairRows + 2 * hd, RT = kPairRows + 2 * h2;
  const int TG = p.TG;                                                 // taps per weight stage
  // csize = 2: the two CTAs of a cluster walk their tiles in lockstep and share the weight stream -- each fetches HALF of
  // every weight stage and multicasts it into both shared memories (cp.async.bulk ... multicast::cluster), a stage slot is
  // re-filled once the MMA threads of BOTH CTAs have released it (tcgen05.commit ... multicast onto both w_empty
  // barriers): the L2 -> SM weight traffic (5.4 TB/s at k = 11) halves.  Measured: no gain -- the kernel is bound by
  // shared-memory bandwidth (operand fetch of back-to-back N' = 128 MMAs takes all 128 B/clk), so the launcher only uses
  // it with DTTS_TC_PAIR64_CLUSTER=1 (tools/gpu_round.sh runs the parity tests that way too).
  const int csize = p.csize;
  const uint32_t rank = csize > 1 ? cluster_ctarank() : 0u;
  const uint32_t bar0 = smem_u32(smem);
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + kQTmemOff);
  float* bias_s = reinterpret_cast<float*>(smem + kQBiasOff);
  int* pref_s = reinterpret_cast<int*>(smem + kQPrefOff);
  int* lim_s = pref_s + TC_MAX_RAGGED_ITEMS + 1;
  const uint32_t w_stage_bytes = (uint32_t)TG * kP64TapBytes;
  const uint32_t a_stage_bytes = 4u * (uint32_t)RA * 16u;             // one 32-channel chunk of a tile
  const uint32_t t_chunk_bytes = 4u * (uint32_t)RT * 16u;
  const int a_stages = p.a_stages;
  const uint32_t t_buf_bytes = 2u * t_chunk_bytes;                    // one intermediate tile: both chunks
  const uint32_t w_base = smem_u32(smem + kQHeader);
  const uint32_t a_base = w_base + kP64WStages * w_stage_bytes;
  const uint32_t t_base = a_base + (uint32_t)a_stages * a_stage_bytes;

  griddep_launch();
  if (p.lens && warp == 3) {
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
    for (int s = 0; s < kP64AStages; ++s) { mbar_init(bar(kQAFull + s), 1); mbar_init(bar(kQAEmpty + s), 1); }
    for (int s = 0; s < kP64WStages; ++s) { mbar_init(bar(kQWFull + s), 1); mbar_init(bar(kQWEmpty + s), (uint32_t)csize); }
    mbar_init(bar(kQAcc1Full), 1); mbar_init(bar(kQAcc1Empty), 8);
    mbar_init(bar(kQAcc2Full), 1); mbar_init(bar(kQAcc2Empty), 8);
    for (int s = 0; s < 2; ++s) { mbar_init(bar(kQTFull + s), 8); mbar_init(bar(kQTEmpty + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem + kQTmemOff)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 96 && threadIdx.x < 96 + 2 * kP64C) {
    const int i = threadIdx.x - 96;
    bias_s[i] = i < kP64C ? __ldg(p.b1 + i) : __ldg(p.b2 + i - kP64C);
  }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();             // the peer's barriers exist before any multicast copy / remote arrive
  tc_fence_after();
  if (warp != 1) griddep_wait();                 // warp 1 only reads the (constant) weights: it may run ahead
  const uint32_t tmem_base = *tmem_ptr_s;
  const int* pref = p.lens ? pref_s : nullptr;
  const int nrt = p.lens ? pref_s[p.B] : p.ntiles * p.B;
  const int G = gridDim.x, cta = blockIdx.x;
  const int n_it = nrt > cta ? (nrt - cta + G - 1) / G : 0;
  // tiles of the cluster's first CTA: the weight stream runs for that many tiles in both CTAs (the second one may have
  // one tile less: it then consumes the last tile's weight stages without issuing MMAs)
  const int cta0 = cta - (int)rank;
  const int n_w = nrt > cta0 ? (nrt - cta0 + G - 1) / G : 0;

  if (warp == 0) {
    // ------------------------------------------------ input producer: two 32-channel chunks per tile
    int s = 0;
    uint32_t ph = 1;
    PairCursor cur;
    for (int it = 0; it < n_it; ++it) {
      const PairTile tc = pair_decode(p, pref, lim_s, (uint32_t)(cta + it * G), cur);
      const size_t row0 = (size_t)(p.a_pad + tc.q0 - h2 - hd);
      const tc16* src = p.a_hi + (size_t)tc.b * p.a_bs;
      for (int c = 0; c < 2; ++c) {
        mbar_wait(bar(kQAEmpty + s), ph);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar(kQAFull + s), a_stage_bytes);
          for (int sl = 0; sl < 4; ++sl)
            bulk_g2s(a_base + s * a_stage_bytes + sl * (uint32_t)RA * 16u,
                     src + ((size_t)(c * 4 + sl) * p.a_rows + row0) * 8, (uint32_t)RA * 16u, bar(kQAFull + s));
        }
        __syncwarp();
        if (++s == a_stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ weight producer: conv1(0) | conv1(1) conv2(0) | conv1(2) conv2(1) ...
    int s = 0;
    uint32_t ph = 1;
    auto stream_conv = [&](const tc16* w) {
      for (int c = 0; c < 2; ++c)
        for (int j0 = 0; j0 < k; j0 += TG) {
          const uint32_t ntap = (uint32_t)min(TG, k - j0);
          mbar_wait(bar(kQWEmpty + s), ph);                         // released by the MMA threads of all CTAs of the cluster
          if (elect_one()) {
            mbar_arrive_expect_tx(bar(kQWFull + s), ntap * kP64TapBytes);   // the whole stage lands here, one slice per CTA
            const tc16* src = w + (size_t)(c * k + j0) * (kP64TapBytes / 2);
            if (csize > 1) {
              const uint32_t half = ntap * kP64TapBytes / 2u;
              bulk_g2s_mc(w_base + s * w_stage_bytes + rank * half, src + (size_t)rank * (half / 2), half, bar(kQWFull + s),
                          (uint16_t)3);
            } else {
              bulk_g2s(w_base + s * w_stage_bytes, src, ntap * kP64TapBytes, bar(kQWFull + s));
            }
          }
          __syncwarp();
          if (++s == kP64WStages) { s = 0; ph ^= 1u; }
        }
    };
    if (n_w > 0) stream_conv(p.w1);
    for (int it = 0; it < n_w; ++it) {
      if (it + 1 < n_w) stream_conv(p.w1);
      stream_conv(p.w2);
    }
  } else if (warp == 2) {
    // ------------------------------------------------ MMA issuer: M1(0), then M1(i + 1), M2(i)
    if (elect_one()) {
      const uint32_t hiw = (128u >> 4) | (1u << 14);
      const uint32_t f16b = p.fmt ? 1u : 0u;
      const uint32_t idesc = (1u << 4) | (f16b << 7) | (f16b << 10) | ((uint32_t)(kP64NM >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t a_low0 = ((a_base >> 4) & 0x3FFFu) | ((uint32_t)RA << 16);
      const uint32_t t_low0 = ((t_base >> 4) & 0x3FFFu) | ((uint32_t)RT << 16);
      const uint32_t w_low0 = ((w_base >> 4) & 0x3FFFu) | ((uint32_t)kP64NM << 16);
      const uint32_t a_kstep = 2u * (uint32_t)RA, t_kstep = 2u * (uint32_t)RT, b_kstep = 2u * kP64NM;
      const uint32_t w_tap16 = kP64TapBytes >> 4, w_stage16 = w_stage_bytes >> 4, a_stage16 = a_stage_bytes >> 4;
      const uint32_t t_chunk16 = t_chunk_bytes >> 4, t_buf16 = t_buf_bytes >> 4;
      int sa = 0, sw = 0;
      uint32_t pa = 0, pw = 0;
      // one convolution of one tile: chunks x tap groups.  from_stage: the staged input chunks; else intermediate tile tb
      auto conv = [&](uint32_t d_base, bool from_stage, uint32_t tap_step, int tb) {
        for (int c = 0; c < 2; ++c) {
          uint32_t a_tap;
          if (from_stage) {
            mbar_wait(bar(kQAFull + sa), pa);
            tc_fence_after();
            a_tap = a_low0 + (uint32_t)sa * a_stage16;
          } else {
            a_tap = t_low0 + (uint32_t)tb * t_buf16 + (uint32_t)c * t_chunk16;
          }
          const uint32_t kst = from_stage ? a_kstep : t_kstep;
          for (int j0 = 0; j0 < k; j0 += TG) {
            mbar_wait(bar(kQWFull + sw), pw);
            tc_fence_after();
            uint32_t b_lo = w_low0 + (uint32_t)sw * w_stage16;
            const int j1 = min(j0 + TG, k);
            for (int j = j0; j < j1; ++j, a_tap += tap_step, b_lo += w_tap16) {
              const uint32_t first = (c | j) != 0 ? 1u : 0u;
#pragma unroll
              for (int m = 0; m < 2; ++m) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                  umma_bf16(d_base + (uint32_t)(m * kP64NM), desc64(a_tap + (uint32_t)(m * 128) + ks * kst, hiw),
                            desc64(b_lo + ks * b_kstep, hiw), idesc, ks == 0 ? first : 1u);
              }
            }
            if (csize > 1) umma_commit_mc(bar(kQWEmpty + sw), (uint16_t)3);   // w_empty here and at the peer's producer
            else umma_commit(bar(kQWEmpty + sw));
            if (++sw == kP64WStages) { sw = 0; pw ^= 1u; }
          }
          if (from_stage) {
            umma_commit(bar(kQAEmpty + sa));
            if (++sa == a_stages) { sa = 0; pa ^= 1u; }
          }
        }
      };
      auto conv1 = [&](int i) {
        mbar_wait(bar(kQAcc1Empty), (i & 1) ^ 1);                     // E1(i - 1) has the accumulators in registers
        tc_fence_after();
        conv(tmem_base, true, (uint32_t)p.dil, 0);
        umma_commit(bar(kQAcc1Full));
      };
      // a tile this CTA does not have (the cluster's odd tile): keep the shared weight pipeline moving, nothing else
      auto skip_conv = [&]() {
        for (int c = 0; c < 2; ++c)
          for (int j0 = 0; j0 < k; j0 += TG) {
            mbar_wait(bar(kQWFull + sw), pw);
            umma_commit_mc(bar(kQWEmpty + sw), (uint16_t)3);
            if (++sw == kP64WStages) { sw = 0; pw ^= 1u; }
          }
      };
      auto conv2 = [&](int i) {
        const int tb = i & 1;
        mbar_wait(bar(kQTFull + tb), (i >> 1) & 1);                   // E1(i) wrote intermediate tile tb
        mbar_wait(bar(kQAcc2Empty), (i & 1) ^ 1);                     // E2(i - 1) has its accumulators in registers
        tc_fence_after();
        conv(tmem_base + 256u, false, 1u, tb);
        umma_commit(bar(kQTEmpty + tb));
        umma_commit(bar(kQAcc2Full));
      };
      auto c1 = [&](int i) { if (i < n_it) conv1(i); else skip_conv(); };
      auto c2 = [&](int i) { if (i < n_it) conv2(i); else skip_conv(); };
      if (n_w > 0) c1(0);
      for (int i = 0; i < n_w; ++i) {
        if (i + 1 < n_w) c1(i + 1);                                   // runs while E1(i) / E2(i - 1) finish
        c2(i);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ epilogue: E1(i), E2(i); warp = (lane quadrant, sub-tile), two
    // 32-channel column chunks each
    const int quad = warp & 3, m = (warp - 3) >> 2;
    const int r = m * 128 + quad * 32 + lane;
    const int fmt = p.fmt;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(m * kP64NM);
    PairCursor cur;
    auto load_res = [&](const PairTile& tc, int cc, bool ok, int t, float4 (&dst)[8]) {
      if (p.res && ok) {
        const float4* rp = reinterpret_cast<const float4*>(p.res + (size_t)tc.b * p.o32_bs) + (size_t)(cc * 8) * p.T + t;
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[q] = rp[(size_t)q * p.T];
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // E1(j): conv1 accumulators of tile j -> + b1, leaky, fp16 -> intermediate tile j & 1
    auto E1 = [&](int j, const PairTile& tj) {
      const int t = tj.q0 - h2 + r;
      const bool inside = t >= 0 && t < p.T;
      const int tb = j & 1;
      mbar_wait(bar(kQAcc1Full), j & 1);
      mbar_wait(bar(kQTEmpty + tb), ((j >> 1) & 1) ^ 1);               // conv2 of tile j - 2 has read this buffer
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t a[32], l[32];
        __syncwarp();
        tmem_ld32_nowait(lane_addr + (uint32_t)(cc * 32), a);
        tmem_ld32_nowait(lane_addr + (uint32_t)(kP64C + cc * 32), l);
        tmem_ld_wait();
        if (cc == 1) {                                                // both chunks are in registers: conv1 of the next tile may go
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(kQAcc1Empty));
        }
        float* af = reinterpret_cast<float*>(a);
        const float* lf = reinterpret_cast<const float*>(l);
#pragma unroll
        for (int q = 0; q < 32; q += 2) add2(af[q], af[q + 1], lf[q], lf[q + 1]);
        const float4* bv = reinterpret_cast<const float4*>(bias_s + cc * 32);
        const uint32_t dst = t_base + (uint32_t)tb * t_buf_bytes + (uint32_t)cc * t_chunk_bytes + (uint32_t)(h2 + r) * 16u;
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
          uint32_t hw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 bq = bv[2 * sl + (e >> 1)];
            float v0 = af[8 * sl + 2 * e], v1 = af[8 * sl + 2 * e + 1];
            fma2(v0, v1, 1.f, (e & 1) ? bq.z : bq.x, (e & 1) ? bq.w : bq.y);
            float l0, l1;
            leaky2(v0, v1, p.slope, l0, l1);
            hw[e] = inside ? pack2(l0, l1, fmt) : 0u;
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)sl * (uint32_t)RT * 16u), "r"(hw[0]),
                       "r"(hw[1]), "r"(hw[2]), "r"(hw[3])
                       : "memory");
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kQTFull + tb));
    };
    PairTile tc{0, 0, 0}, tn{0, 0, 0};
    if (n_it > 0) {
      tc = pair_decode(p, pref, lim_s, (uint32_t)cta, cur);
      E1(0, tc);
    }
    for (int i = 0; i < n_it; ++i) {
      const int t = tc.q0 - h2 + r;
      const bool ok2 = r >= h2 && r < kPairRows - h2 && t < tc.lim;
      float4 rc[8];
      load_res(tc, 0, ok2, t, rc);                                    // in flight while E1(i + 1) runs
      if (i + 1 < n_it) {
        tn = pair_decode(p, pref, lim_s, (uint32_t)(cta + (i + 1) * G), cur);
        E1(i + 1, tn);
      }
      // ---- E2(i): conv2 accumulators -> + b2 + residual, 1/3 mean, fp32 stream + operand planes
      mbar_wait(bar(kQAcc2Full), i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t a[32], l[32];
        __syncwarp();
        tmem_ld32_nowait(lane_addr + 256u + (uint32_t)(cc * 32), a);
        tmem_ld32_nowait(lane_addr + 256u + (uint32_t)(kP64C + cc * 32), l);
        float4 rn[8];
        if (cc == 0) load_res(tc, 1, ok2, t, rn);                     // the second chunk's residual behind the first one's math
        tmem_ld_wait();
        if (cc == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(kQAcc2Empty));               // both chunks are in registers
        }
        if (ok2) {
          float* v = reinterpret_cast<float*>(a);
          const float* lf = reinterpret_cast<const float*>(l);
#pragma unroll
          for (int q = 0; q < 32; q += 2) add2(v[q], v[q + 1], lf[q], lf[q + 1]);
          const float4* bv = reinterpret_cast<const float4*>(bias_s + kP64C + cc * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bq = bv[q];
            fma2(v[4 * q], v[4 * q + 1], 1.f, bq.x, bq.y);
            fma2(v[4 * q + 2], v[4 * q + 3], 1.f, bq.z, bq.w);
          }
          if (p.res) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              add2(v[4 * q], v[4 * q + 1], rc[q].x, rc[q].y);
              add2(v[4 * q + 2], v[4 * q + 3], rc[q].z, rc[q].w);
            }
          }
          if (p.post != 1.f) {
#pragma unroll
            for (int q = 0; q < 32; q += 2) mul2(v[q], v[q + 1], p.post, p.post);
          }
          if (p.o32) {
            float4* op = reinterpret_cast<float4*>(p.o32 + (size_t)tc.b * p.o32_bs) + (size_t)(cc * 8) * p.T + t;
            if (p.accumulate) {
              float4 old[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) old[q] = op[(size_t)q * p.T];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                v[4 * q] += old[q].x; v[4 * q + 1] += old[q].y; v[4 * q + 2] += old[q].z; v[4 * q + 3] += old[q].w;
              }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) op[(size_t)q * p.T] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
          if (p.o_hi) {
            const size_t prow = (size_t)tc.b * p.op_bs + ((size_t)(cc * 4) * p.op_rows + p.op_pad + t) * 8;
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
              uint32_t hw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float l0, l1;
                leaky2(v[8 * sl + 2 * e], v[8 * sl + 2 * e + 1], p.slope, l0, l1);
                hw[e] = pack2(l0, l1, fmt);
              }
              *reinterpret_cast<uint4*>(p.o_hi + prow + (size_t)sl * p.op_rows * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            }
          }
        }
        if (cc == 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q) rc[q] = rn[q];
        }
      }
      tc = tn;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();             // no peer may still multicast into / arrive on this CTA's shared memory
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int pair64_tg(int k) {                           // taps per weight stage: <= 4 (32 KB), balanced (11 -> 4 + 4 + 3)
  const int groups = (k + kP64MaxTG - 1) / kP64MaxTG;
  return (k + groups - 1) / groups;
}
size_t pair64_smem_bytes(int k, int dil, int a_stages) {
  const int h2 = (k - 1) / 2, hd = h2 * dil;
  return (size_t)kQHeader + (size_t)kP64WStages * pair64_tg(k) * kP64TapBytes +
         (size_t)a_stages * 4 * (kPairRows + 2 * hd) * 16 + (size_t)2 * 8 * (kPairRows + 2 * h2) * 16;
}
int pair64_a_stages(int k, int dil) {            // three input stages when they fit next to the two intermediate tiles
  return pair64_smem_bytes(k, dil, kP64AStages) <= (size_t)227 * 1024 ? kP64AStages : 2;
}

size_t pair_smem_bytes(int k, int dil) {
  const int h2 = (k - 1) / 2, hd = h2 * dil;
  return (size_t)kPHeader + 2 * (size_t)k * kTapBytes + (size_t)kPairAStages * 4 * (kPairRows + 2 * hd) * 16 +
         2 * (size_t)4 * (kPairRows + 2 * h2) * 16;
}

}  // namespace

int rb_pair_supported(const TcConvW& c1, const TcConvW& c2, int dil, int a_planes) {
  const int C = c1.C_in;
  if (C == 128) return rb_pair128_supported(c1, c2, dil, a_planes);
  if ((C != kPairC && C != kP64C) || c1.C_out != C || c2.C_in != C || c2.C_out != C) return 0;
  if (c1.ktaps != c2.ktaps || !(c1.ktaps & 1) || c1.ktaps > 11 || dil < 1) return 0;
  if (!c1.stack || !c2.stack || c1.planes != 1 || c2.planes != 1 || c1.N != C || c2.N != C || c1.KC != 32 ||
      c2.KC != 32 || c1.il_u || c2.il_u || c1.pair || c2.pair || c1.lo8 || c2.lo8 || a_planes != 1 || c1.fmt != c2.fmt)
    return 0;
  const int h2 = (c1.ktaps - 1) / 2;
  if (h2 + h2 * dil > TC_PADF) return 0;                              // the conv1 halo of the first tile starts inside the front padding
  if (C == kP64C && !tc_fuse64_enabled()) return 0;
  return (C == kPairC ? pair_smem_bytes(c1.ktaps, dil)
                      : pair64_smem_bytes(c1.ktaps, dil, pair64_a_stages(c1.ktaps, dil))) <= 227 * 1024;
}

cudaError_t launch_rb_pair(RbPairParams p, cudaStream_t stream) {
  if (p.B <= 0 || p.T <= 0) return cudaSuccess;
  if (p.C == 128) return launch_rb_pair128(p, stream);
  if (p.lens && p.B > TC_MAX_RAGGED_ITEMS) return cudaErrorInvalidValue;
  const int h2 = (p.k - 1) / 2, hd = h2 * p.dil;
  p.S = kPairRows - 2 * h2;
  p.ntiles = cdiv(p.T, p.S);
  // the last tile stages rows up to q0 - h2 + 256 + hd of the input planes: they must exist (tc_rows keeps TC_PADB + slack)
  if (p.a_pad - h2 - hd < 0 || p.a_pad + (p.ntiles - 1) * p.S - h2 + kPairRows + hd > p.a_rows) return cudaErrorInvalidValue;
  if (p.C != kPairC && p.C != kP64C) return cudaErrorInvalidValue;
  if (p.post_part && (p.C != kPairC || !p.post_w)) return cudaErrorInvalidValue;
  const bool wide = p.C == kP64C;
  p.TG = wide ? pair64_tg(p.k) : 0;
  p.a_stages = wide ? pair64_a_stages(p.k, p.dil) : kPairAStages;
  const size_t smem = wide ? pair64_smem_bytes(p.k, p.dil, p.a_stages) : pair_smem_bytes(p.k, p.dil);
  if (smem > (size_t)227 * 1024) return cudaErrorInvalidConfiguration;
  static unsigned long long done32 = 0, done64 = 0;
  const cudaError_t attr_err = p.C == kPairC ? ensure_max_dyn_smem(rb_pair32_kernel, 227 * 1024, &done32)
                                             : ensure_max_dyn_smem(rb_pair64_kernel, 227 * 1024, &done64);
  if (attr_err != cudaSuccess) return attr_err;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  const long tiles = (long)p.ntiles * p.B;
  int grid = (int)(tiles < sms ? tiles : sms);
  // the CTA owns all 512 TMEM columns: it must be alone on its SM (shared memory above half of the SM's guarantees it)
  const size_t smem_launch = smem < 116 * 1024 ? 116 * 1024 : smem;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  cfg.blockDim = dim3(kPairThreads);
  cfg.dynamicSmemBytes = smem_launch;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = tc_pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  p.csize = 1;
  if (wide && tc_pair64_cluster_enabled() && tiles >= 2) {
    // clusters of two CTAs sharing the weight stream; as many as can be co-resident (queried once per device)
    static std::mutex mu;
    static int max_clusters[64] = {0};
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64) {
      attr[1].id = cudaLaunchAttributeClusterDimension;
      attr[1].val.clusterDim = {2, 1, 1};
      cfg.numAttrs = 2;
      if (max_clusters[dev] == 0) {
        int n = 0;
        cfg.gridDim = dim3((unsigned)(sms / 2 * 2));
        cfg.dynamicSmemBytes = 227 * 1024;
        cudaError_t q = cudaOccupancyMaxActiveClusters(&n, rb_pair64_kernel, &cfg);
        cfg.dynamicSmemBytes = smem_launch;
        max_clusters[dev] = (q == cudaSuccess && n > 0) ? n : -1;
        if (q != cudaSuccess) (void)cudaGetLastError();
      }
      if (max_clusters[dev] > 0) {
        const long want = tiles / 2;
        const int nc = (int)(want < max_clusters[dev] ? want : max_clusters[dev]);
        grid = 2 * nc;
        p.csize = 2;
      } else {
        cfg.numAttrs = 1;
      }
    }
  }
  cfg.gridDim = dim3((unsigned)grid);
  return wide ? cudaLaunchKernelEx(&cfg, rb_pair64_kernel, p) : cudaLaunchKernelEx(&cfg, rb_pair32_kernel, p);
}

}  // namespace dtts
