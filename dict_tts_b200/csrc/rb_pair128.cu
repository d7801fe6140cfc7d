// Fused ResBlock convolution pair for C = 128 (the second HiFi-GAN stage, 264 of the generator's 614 MFLOP per frame):
// conv1 (dilated) -> leaky-ReLU -> conv2 (dilation 1) -> + residual in ONE launch, on CTA pairs (tcgen05 cta_group::2).
// Reference: ResBlock1.forward, modules/hifigan/hifigan.py:51-58.
//
// Unfused (tc_conv_kernel<true> twice) the pair moves 16 B per element through HBM and its second launch is HBM bound
// (k = 3: 361 us at 5.5 TB/s for a 140 us tensor floor; k = 7: 493 us against 366 us for the same MMAs in conv1); fused it
// moves 12 B and the intermediate activation -- fp16 plus, for the FP8 lo plane, its e5m2 copy -- lives in shared memory.
//
// Same operands and the same MMA order as tc_conv_kernel<true> in its pair mode: M = 256 across the two CTAs of a
// cluster (each CTA owns a 256-row tile = two 128-row sub-tiles and stages its own input), N = 128 with each CTA holding
// half of every weight blob, K = 32 channels per chunk; weights either as two fp16 planes (one MMA each) or fp16 x 2^10 +
// e5m2 lo plane at the FP8 rate (TcMode::lo8), both accumulating into the same TMEM columns.  The intermediate is
// rounded exactly like the operand planes the unfused pair writes, so the result is BIT-IDENTICAL to the two launches
// (tests/test_gpu_tensorcore.py).
//
// Shared memory holds ONE intermediate tile (128 channels x 266 rows x 3 B = 102 KB with the e5m2 copy), two to four input
// chunk stages and a ring of up to twelve one-tap weight stages; TMEM one accumulator set per convolution (2 x 128 columns each).  Order:
//     tensor pipe   C1(0) | C1(1) C2(0) | C1(2) C2(1) | ...
//     epilogue      E1a(0) E1b(0) | E1a(1) . E1b(1) E2(0) | E1a(2) . E1b(2) E2(1) | ...
// E1a(i+1) turns the conv1 accumulators into packed fp16 IN REGISTERS (64 per thread) while C2(i) still reads the
// intermediate tile, E1b(i+1) stores them once C2(i) has completed, E2(i) (residual, fp32 stream, operand planes) runs
// under C1(i+2).
// Warps: 0 input producer (+ the e5m2 copy of every staged chunk), 1 weight producer, 2 MMA issue (leader CTA) / relay of
// "my stage is full" to the leader (peer CTA), 3-10 epilogue (lane quadrant x sub-tile).
#include "tc_conv.cuh"
#include "tc16.cuh"
#include "tc_ptx.cuh"
#include "rb_pair_common.cuh"

#include <cstdlib>
#include <mutex>

namespace dtts {

namespace {

constexpr int kWC = 128;                          // channels
constexpr int kWChunks = 4;                       // K chunks of 32 channels
constexpr int kWMaxAStages = 4, kWMaxWStages = 12;
constexpr int kWHalfN = kWC / 2;                  // rows of the B operand held by one CTA
constexpr int kWPlaneBytes = kWHalfN * 32 * 2;    // one fp16 weight plane of one (chunk, tap), this CTA's half: 4 KB
constexpr int kW8Bytes = kWHalfN * 32;            // its e5m2 lo plane: 2 KB
// mbarriers
constexpr int kWAFull = 0, kWAEmpty = kWAFull + kWMaxAStages, kWA8 = kWAEmpty + kWMaxAStages,
              kWPAFull = kWA8 + kWMaxAStages, kWWFull = kWPAFull + kWMaxAStages, kWWEmpty = kWWFull + kWMaxWStages,
              kWPWFull = kWWEmpty + kWMaxWStages, kWAcc1Full = kWPWFull + kWMaxWStages, kWAcc1Empty = kWAcc1Full + 1,
              kWAcc2Full = kWAcc1Empty + 1, kWAcc2Empty = kWAcc2Full + 1, kWTFull = kWAcc2Empty + 1,
              kWTEmpty = kWTFull + 1, kWNumBars = kWTEmpty + 1;
constexpr int kWTmemOff = kWNumBars * 8;
constexpr int kWBiasOff = (kWTmemOff + 4 + 63) / 64 * 64;                 // b1[128], b2[128]
constexpr int kWPrefOff = kWBiasOff + 2 * kWC * 4;
constexpr int kWHeader = (kWPrefOff + (2 * TC_MAX_RAGGED_ITEMS + 8) * 4 + 127) / 128 * 128;

// Developer trace (-DDTTS_P128_TRACE): clock64 stamps of the pipeline events of CTA 0, tiles 0..kTrTiles-1, read back with
// dtts_debug_p128_trace (tools/p128_trace.py prints the timeline).
#ifdef DTTS_P128_TRACE
constexpr int kTrTiles = 24, kTrEvents = 20;
__device__ long long g_p128_trace[kTrTiles * kTrEvents];
__device__ int g_p128_trace_sel;                  // k * 16 + dilation of the launches that record
#define P128_ON (blockIdx.x == 0 && p.k * 16 + p.dil == g_p128_trace_sel)
#define P128_TR(tile, ev) do { if (P128_ON && (tile) < kTrTiles) g_p128_trace[(tile) * kTrEvents + (ev)] = clock64(); } while (0)
#define P128_ADD(tile, ev, v) do { if (P128_ON && (tile) < kTrTiles) g_p128_trace[(tile) * kTrEvents + (ev)] = (v); } while (0)
#define P128_TIMED(acc, stmt) do { const long long t0_ = clock64(); stmt; acc += clock64() - t0_; } while (0)
#else
#define P128_TR(tile, ev) do { } while (0)
#define P128_ADD(tile, ev, v) do { } while (0)
#define P128_TIMED(acc, stmt) do { stmt; } while (0)
#endif

__global__ void __launch_bounds__(kPairThreads, 1) rb_pair128_kernel(const RbPairParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = p.k, h2 = (p.k - 1) / 2, hd = h2 * p.dil;
  const int RA = kPairRows + 2 * hd, RT = kPairRows + 2 * h2;
  const int TG = p.TG, a_stages = p.a_stages, w_stages = p.w_stages, WPL = p.w_planes;
  const bool lo8 = p.lo8 != 0;
  const uint32_t rank = cluster_ctarank();
  const uint32_t bar0 = smem_u32(smem);
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + kWTmemOff);
  float* bias_s = reinterpret_cast<float*>(smem + kWBiasOff);
  int* pref_s = reinterpret_cast<int*>(smem + kWPrefOff);
  int* lim_s = pref_s + TC_MAX_RAGGED_ITEMS + 1;
  const uint32_t w_tap_bytes = (uint32_t)(kWPlaneBytes * WPL), w8_tap_bytes = lo8 ? (uint32_t)kW8Bytes : 0u;
  const uint32_t w_stage_bytes = (uint32_t)TG * (w_tap_bytes + w8_tap_bytes);      // TG x [fp16 plane(s) | e5m2 plane]
  const uint32_t a_plane_bytes = 4u * (uint32_t)RA * 16u;                            // one 32-channel chunk of a tile
  const uint32_t a_stage_bytes = a_plane_bytes + (lo8 ? 2u * (uint32_t)RA * 16u : 0u);
  const uint32_t t_chunk_bytes = 4u * (uint32_t)RT * 16u, t8_chunk_bytes = 2u * (uint32_t)RT * 16u;
  const uint32_t w_base = smem_u32(smem + kWHeader);
  const uint32_t a_base = w_base + (uint32_t)w_stages * w_stage_bytes;
  const uint32_t t_base = a_base + (uint32_t)a_stages * a_stage_bytes;               // intermediate tile: 4 fp16 chunks,
  const uint32_t t8_base = t_base + kWChunks * t_chunk_bytes;                        // then (lo8) 4 e5m2 chunks

  griddep_launch();
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
    for (int s = 0; s < kWMaxAStages; ++s) {
      mbar_init(bar(kWAFull + s), 1); mbar_init(bar(kWAEmpty + s), 1);
      mbar_init(bar(kWA8 + s), 1);    mbar_init(bar(kWPAFull + s), 1);
    }
    for (int s = 0; s < kWMaxWStages; ++s) {
      mbar_init(bar(kWWFull + s), 1); mbar_init(bar(kWWEmpty + s), 1); mbar_init(bar(kWPWFull + s), 1);
    }
    // the leader's accumulator / intermediate-tile barriers collect the epilogue warps of BOTH CTAs
    mbar_init(bar(kWAcc1Full), 1); mbar_init(bar(kWAcc1Empty), 16);
    mbar_init(bar(kWAcc2Full), 1); mbar_init(bar(kWAcc2Empty), 16);
    mbar_init(bar(kWTFull), 16);   mbar_init(bar(kWTEmpty), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem + kWTmemOff)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 96) {
    const int i = threadIdx.x - 96;               // 256 epilogue threads: b1 | b2
    bias_s[i] = i < kWC ? __ldg(p.b1 + i) : __ldg(p.b2 + i - kWC);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // the peer's barriers exist before any remote arrive / multicast commit
  tc_fence_after();
  if (warp != 1) griddep_wait();                 // warp 1 only reads the (constant) weights: it may run ahead
  const uint32_t tmem_base = *tmem_ptr_s;
  const int* pref = p.lens ? pref_s : nullptr;
  const int nrt = p.lens ? pref_s[p.B] : p.ntiles * p.B;
  // unit u = row tiles (2u, 2u + 1), one per CTA of the pair; cluster c walks units c, c + nclusters, ...  An odd tile
  // count leaves the peer of the last unit a dummy: it recomputes the last tile and stores nothing.
  const int units = (nrt + 1) / 2;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int n_it = units > cluster_id ? (units - cluster_id + nclusters - 1) / nclusters : 0;
  auto tile_rt = [&](int it, bool& dummy) {
    const uint32_t rt = 2u * (uint32_t)(cluster_id + it * nclusters) + rank;
    dummy = rt >= (uint32_t)nrt;
    return dummy ? (uint32_t)(nrt - 1) : rt;
  };

  if (warp == 0) {
    // ------------------------------------------------ input producer: four 32-channel chunks per tile, and (lo8) the
    // e5m2 copy of every staged chunk: e5m2 is fp16 with the mantissa cut to 2 bits = the high byte of every element.
    // A copy is issued whenever its slot is free, a landed chunk is converted as soon as possible (the MMAs wait for it).
    const int total = n_it * kWChunks;
    int n_issue = 0, n_conv = 0;
    int si = 0, sc = 0;
    uint32_t phi = 1, phc = 0;
    PairCursor cur, cur_pf;
    PairTile tc{0, 0, 0};
    auto issue = [&]() {
      const int c = n_issue & 3;
      if (c == 0) {
        bool dummy;
        const uint32_t rt = tile_rt(n_issue >> 2, dummy);
        tc = pair_decode(p, pref, lim_s, rt, cur);
        // The input planes come from HBM (~2 us under load) and only two or three chunks can be staged ahead, so the
        // chunks of the NEXT tile are pulled into L2 now: one cp.async.bulk.prefetch per 8-channel slab, 16 lanes.
        if (p.pf && (n_issue >> 2) + 1 < n_it && lane < 16) {
          bool d2;
          const PairTile tp = pair_decode(p, pref, lim_s, tile_rt((n_issue >> 2) + 1, d2), cur_pf);
          const tc16* src = p.a_hi + (size_t)tp.b * p.a_bs + ((size_t)lane * p.a_rows + (size_t)(p.a_pad + tp.q0 - h2 - hd)) * 8;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)RA * 16u) : "memory");
        }
      }
      if (elect_one()) {
        const size_t row0 = (size_t)(p.a_pad + tc.q0 - h2 - hd);
        const tc16* src = p.a_hi + (size_t)tc.b * p.a_bs;
        mbar_arrive_expect_tx(bar(kWAFull + si), a_plane_bytes);
        for (int sl = 0; sl < 4; ++sl)
          bulk_g2s(a_base + si * a_stage_bytes + sl * (uint32_t)RA * 16u, src + ((size_t)(c * 4 + sl) * p.a_rows + row0) * 8,
                   (uint32_t)RA * 16u, bar(kWAFull + si));
      }
      __syncwarp();
      ++n_issue;
      if (++si == a_stages) { si = 0; phi ^= 1u; }
    };
    auto convert = [&]() {
      mbar_wait(bar(kWAFull + sc), phc);
      const uint32_t src = a_base + sc * a_stage_bytes, dst = src + a_plane_bytes;
      const uint32_t slab = (uint32_t)RA * 16u;
      for (int s8 = 0; s8 < 2; ++s8) {                             // e5m2 slab s8 = high bytes of fp16 slabs 2*s8, 2*s8 + 1
        const uint32_t s0 = src + (uint32_t)(2 * s8) * slab, d0 = dst + (uint32_t)s8 * slab;
#pragma unroll 4
        for (int r = lane; r < RA; r += 32) {
          uint4 lo, hi;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "r"(s0 + (uint32_t)r * 16u));
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "r"(s0 + slab + (uint32_t)r * 16u));
          const uint32_t o0 = __byte_perm(lo.x, lo.y, 0x7531), o1 = __byte_perm(lo.z, lo.w, 0x7531);
          const uint32_t o2 = __byte_perm(hi.x, hi.y, 0x7531), o3 = __byte_perm(hi.z, hi.w, 0x7531);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d0 + (uint32_t)r * 16u), "r"(o0), "r"(o1), "r"(o2),
                       "r"(o3)
                       : "memory");
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kWA8 + sc));
      ++n_conv;
      if (++sc == a_stages) { sc = 0; phc ^= 1u; }
    };
    if (!lo8) {
      while (n_issue < total) {
        mbar_wait(bar(kWAEmpty + si), phi);
        issue();
      }
    } else {
      while (n_conv < total) {
        bool can_issue = false;
        if (n_issue < total) {
          uint32_t ok = lane == 0 ? (mbar_test_wait(bar(kWAEmpty + si), phi) ? 1u : 0u) : 0u;
          can_issue = __shfl_sync(0xffffffffu, ok, 0) != 0;
        }
        if (can_issue) {
          issue();
        } else if (n_conv < n_issue) {
          convert();
        } else {                                                   // nothing landed, no free slot: wait for a slot
          mbar_wait(bar(kWAEmpty + si), phi);
          issue();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ weight producer: this CTA's half of every stage -- ONE bulk copy out of
    // the per-CTA weight stream -- in the order the MMA thread consumes them: conv1(0) | conv1(1) conv2(0) | conv1(2) ...
    int s = 0;
    uint32_t ph = 1;
    const uint32_t tap_bytes = w_tap_bytes + w8_tap_bytes;
    const size_t half_bytes = (size_t)kWChunks * k * tap_bytes;       // one CTA's share of a convolution
    auto stream_conv = [&](const uint8_t* ws, int tr_tile) {
      const uint8_t* src = ws + (size_t)rank * half_bytes;
      for (int c = 0; c < kWChunks; ++c)
        for (int j0 = 0; j0 < k; j0 += TG) {
          const uint32_t bytes = (uint32_t)min(TG, k - j0) * tap_bytes;
          mbar_wait(bar(kWWEmpty + s), ph);
          if (tr_tile >= 0 && c == 2 && j0 == 0 && lane == 0) P128_TR(tr_tile, 17);
          if (elect_one()) {
            mbar_arrive_expect_tx(bar(kWWFull + s), bytes);
            bulk_g2s(w_base + s * w_stage_bytes, src, bytes, bar(kWWFull + s));
          }
          __syncwarp();
#ifdef DTTS_P128_TRACE
          if (tr_tile >= 0 && c == 2 && j0 == 0 && (p.csize & 16)) {   // experiment: the producer watches this copy land
            const long long t0 = clock64();
            mbar_wait(bar(kWWFull + s), ph ^ 1u);
            if (lane == 0) P128_ADD(tr_tile, 13, clock64() - t0);
          }
#endif
          src += bytes;
          if (++s == w_stages) { s = 0; ph ^= 1u; }
        }
    };
    if (n_it > 0) stream_conv(p.w1s, -1);
    for (int it = 0; it < n_it; ++it) {
      if (it + 1 < n_it) stream_conv(p.w1s, -1);
      stream_conv(p.w2s, it);                      // (trace build: stamps the first stage of conv2's third chunk)
    }
  } else if (warp == 2) {
    if (rank != 0) {
      // ------------------------------------------------ peer CTA: forward "my stage is full" to the leader, in pipeline order
      if (elect_one()) {
        int sa = 0, sw = 0;
        uint32_t pa = 0, pw = 0;
        auto relay = [&](bool from_stage) {
          for (int c = 0; c < kWChunks; ++c) {
            if (from_stage) {
              if (lo8) {                          // the e5m2 copy was written by the generic proxy: release at cluster scope
                mbar_wait(bar(kWA8 + sa), pa);
                mbar_arrive_remote_release(bar(kWPAFull + sa), 0);
              } else {
                mbar_wait(bar(kWAFull + sa), pa);
                mbar_arrive_remote(bar(kWPAFull + sa), 0);
              }
              if (++sa == a_stages) { sa = 0; pa ^= 1u; }
            }
            for (int j0 = 0; j0 < k; j0 += TG) {
              mbar_wait(bar(kWWFull + sw), pw);
              mbar_arrive_remote(bar(kWPWFull + sw), 0);
              if (++sw == w_stages) { sw = 0; pw ^= 1u; }
            }
          }
        };
        if (n_it > 0) relay(true);
        for (int i = 0; i < n_it; ++i) {
          if (i + 1 < n_it) relay(true);
          relay(false);
        }
      }
      __syncwarp();
    } else if (elect_one()) {
      // ------------------------------------------------ leader CTA: MMA issue, C1(0), then C1(i + 1), C2(i)
      const uint32_t hiw = (128u >> 4) | (1u << 14);                  // SBO = 128 B, descriptor version 1
      const uint32_t f16b = p.fmt ? 1u : 0u;
      const uint32_t idesc = (1u << 4) | (f16b << 7) | (f16b << 10) | ((uint32_t)(kWC >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t idesc8 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kWC >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t a_low0 = ((a_base >> 4) & 0x3FFFu) | ((uint32_t)RA << 16);
      const uint32_t t_low0 = ((t_base >> 4) & 0x3FFFu) | ((uint32_t)RT << 16);
      const uint32_t t8_low0 = ((t8_base >> 4) & 0x3FFFu) | ((uint32_t)RT << 16);
      const uint32_t w_low0 = ((w_base >> 4) & 0x3FFFu) | ((uint32_t)kWHalfN << 16);
      const uint32_t a_kstep = 2u * (uint32_t)RA, t_kstep = 2u * (uint32_t)RT, b_kstep = 2u * kWHalfN;
      const uint32_t b_plane = kWPlaneBytes >> 4;
      const uint32_t w_tap16 = (w_tap_bytes + w8_tap_bytes) >> 4, w8_off16 = w_tap_bytes >> 4, w_stage16 = w_stage_bytes >> 4;
      const uint32_t a_stage16 = a_stage_bytes >> 4, a8_delta = a_plane_bytes >> 4;
      const uint32_t t_chunk16 = t_chunk_bytes >> 4, t8_chunk16 = t8_chunk_bytes >> 4;
      int sa = 0, sw = 0;
      uint32_t pa = 0, pw = 0;
      long long tr_a = 0, tr_w = 0;                                     // (trace build) cycles spent waiting for input / weights
      int tr_tile = 0;
      auto conv = [&](uint32_t d_base, bool from_stage, uint32_t tap_step) {
        for (int c = 0; c < kWChunks; ++c) {
          uint32_t a_tap, a8;
          if (from_stage) {
            P128_TIMED(tr_a, mbar_wait(bar((lo8 ? kWA8 : kWAFull) + sa), pa));
            P128_TIMED(tr_a, mbar_wait(bar(kWPAFull + sa), pa));        // ... and the peer's chunk
            tc_fence_after();
            a_tap = a_low0 + (uint32_t)sa * a_stage16;
            a8 = a_tap + a8_delta;
          } else {
            a_tap = t_low0 + (uint32_t)c * t_chunk16;
            a8 = t8_low0 + (uint32_t)c * t8_chunk16;
          }
          const uint32_t kst = from_stage ? a_kstep : t_kstep;
          for (int j0 = 0; j0 < k; j0 += TG) {
#ifdef DTTS_P128_TRACE
            if (p.csize & 4) { /* experiment: never wait for weights (results are garbage) */ } else
#endif
            P128_TIMED(tr_w, mbar_wait(bar(kWWFull + sw), pw));
            if (!from_stage && c == 2 && j0 == 0) P128_TR(tr_tile, 18);
#ifdef DTTS_P128_TRACE
            if (p.csize & 12) { /* experiment: never wait for the peer's relay */ } else
#endif
            P128_TIMED(tr_w, mbar_wait(bar(kWPWFull + sw), pw));        // ... and the peer's half of the weights
            if (!from_stage && c == 2 && j0 == 0) P128_TR(tr_tile, 19);
            tc_fence_after();
            uint32_t b_lo = w_low0 + (uint32_t)sw * w_stage16;
            const int j1 = min(j0 + TG, k);
            for (int j = j0; j < j1; ++j, a_tap += tap_step, a8 += tap_step, b_lo += w_tap16) {
              const uint32_t first = (c | j) != 0 ? 1u : 0u;
#pragma unroll
              for (int m = 0; m < 2; ++m) {
                const uint32_t d = d_base + (uint32_t)(m * kWC), am = a_tap + (uint32_t)(m * 128);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  const uint64_t a = desc64(am + ks * kst, hiw), b = desc64(b_lo + ks * b_kstep, hiw);
                  umma_bf16_2cta(d, a, b, idesc, ks == 0 ? first : 1u);
                  if (WPL == 2) umma_bf16_2cta(d, a, desc64(b_lo + ks * b_kstep + b_plane, hiw), idesc, 1u);
                }
                if (lo8) umma_f8_2cta(d, desc64(a8 + (uint32_t)(m * 128), hiw), desc64(b_lo + w8_off16, hiw), idesc8, 1u);
              }
            }
            umma_commit_2cta(bar(kWWEmpty + sw));                       // w_empty of both CTAs
            if (++sw == w_stages) { sw = 0; pw ^= 1u; }
          }
          if (from_stage) {
            umma_commit_2cta(bar(kWAEmpty + sa));                       // a_empty of both CTAs
            if (++sa == a_stages) { sa = 0; pa ^= 1u; }
          }
        }
      };
      auto conv1 = [&](int i) {
        P128_TR(i, 0);
        mbar_wait(bar(kWAcc1Empty), (i & 1) ^ 1);                       // E1a(i - 1) of both CTAs has the accumulators in registers
        tc_fence_after();
        P128_TR(i, 1);
        tr_a = tr_w = 0;
        conv(tmem_base, true, (uint32_t)p.dil);
        umma_commit_2cta(bar(kWAcc1Full));
        P128_TR(i, 2);
        P128_ADD(i, 14, tr_a);
        P128_ADD(i, 15, tr_w);
      };
      auto conv2 = [&](int i) {
        P128_TR(i, 3);
        mbar_wait(bar(kWTFull), i & 1);                                 // E1b(i) of both CTAs wrote the intermediate tiles
        P128_TR(i, 4);
        mbar_wait(bar(kWAcc2Empty), (i & 1) ^ 1);                       // E2(i - 1) has its accumulators in registers
        tc_fence_after();
        P128_TR(i, 5);
        tr_a = tr_w = 0;
        tr_tile = i;
        conv(tmem_base + 256u, false, 1u);
        P128_ADD(i, 16, tr_w);
        umma_commit_2cta(bar(kWTEmpty));
        umma_commit_2cta(bar(kWAcc2Full));
        P128_TR(i, 6);
      };
      if (n_it > 0) conv1(0);
      for (int i = 0; i < n_it; ++i) {
        if (i + 1 < n_it) conv1(i + 1);
        conv2(i);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ epilogue: warp = (TMEM lane quadrant, sub-tile); one row per thread
    const int quad = warp & 3, m = (warp - 3) >> 2;
    const int r = m * 128 + quad * 32 + lane;
    const int fmt = p.fmt;
    const float acc_scale = p.acc_scale;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(m * kWC);
    auto arrive_leader = [&](int b, bool release) {                   // one arrival per warp on a barrier of the leader CTA
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(bar(b));
        else if (release) mbar_arrive_remote_release(bar(b), 0);
        else mbar_arrive_remote(bar(b), 0);
      }
    };
    PairCursor cur;
    uint32_t tp[4][16];                                               // E1a -> E1b: the row's 128 channels as packed fp16
    // E1a(j): conv1 accumulators of tile j -> * acc_scale + b1, leaky, fp16 (zero outside [0, T): conv2's zero padding)
    auto E1a = [&](int j, const PairTile& tj) {
      const int t = tj.q0 - h2 + r;
      const bool inside = t >= 0 && t < p.T;
      mbar_wait(bar(kWAcc1Full), j & 1);
      tc_fence_after();
      if (threadIdx.x == 96) P128_TR(j, 7);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t a[32];
        __syncwarp();
        tmem_ld32_nowait(lane_addr + (uint32_t)(cc * 32), a);
        tmem_ld_wait();
        if (cc == 3) {                                                // all four chunks are in registers: conv1 of the next tile may go
          tc_fence_before();
          arrive_leader(kWAcc1Empty, false);
        }
        float* af = reinterpret_cast<float*>(a);
        const float4* bv = reinterpret_cast<const float4*>(bias_s + cc * 32);
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 bq = bv[2 * sl + (e >> 1)];
            float v0 = af[8 * sl + 2 * e], v1 = af[8 * sl + 2 * e + 1];
            fma2(v0, v1, acc_scale, (e & 1) ? bq.z : bq.x, (e & 1) ? bq.w : bq.y);
            float l0, l1;
            leaky2(v0, v1, p.slope, l0, l1);
            tp[cc][4 * sl + e] = inside ? pack2(l0, l1, fmt) : 0u;
          }
        }
      }
    };
    // E1b(j): packed rows -> intermediate tile (+ its e5m2 copy) once conv2 of tile j - 1 has read the previous one
    auto E1b = [&](int j) {
      if (threadIdx.x == 96) P128_TR(j, 8);
      mbar_wait(bar(kWTEmpty), (j & 1) ^ 1);
      if (threadIdx.x == 96) P128_TR(j, 9);
      const uint32_t row = (uint32_t)(h2 + r) * 16u;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
        for (int sl = 0; sl < 4; ++sl)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t_base + (uint32_t)(cc * 4 + sl) * (uint32_t)RT * 16u + row),
                       "r"(tp[cc][4 * sl]), "r"(tp[cc][4 * sl + 1]), "r"(tp[cc][4 * sl + 2]), "r"(tp[cc][4 * sl + 3])
                       : "memory");
        if (lo8) {
#pragma unroll
          for (int s8 = 0; s8 < 2; ++s8) {
            const uint32_t o0 = __byte_perm(tp[cc][8 * s8], tp[cc][8 * s8 + 1], 0x7531);
            const uint32_t o1 = __byte_perm(tp[cc][8 * s8 + 2], tp[cc][8 * s8 + 3], 0x7531);
            const uint32_t o2 = __byte_perm(tp[cc][8 * s8 + 4], tp[cc][8 * s8 + 5], 0x7531);
            const uint32_t o3 = __byte_perm(tp[cc][8 * s8 + 6], tp[cc][8 * s8 + 7], 0x7531);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t8_base + (uint32_t)(cc * 2 + s8) * (uint32_t)RT * 16u + row),
                         "r"(o0), "r"(o1), "r"(o2), "r"(o3)
                         : "memory");
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      arrive_leader(kWTFull, true);
      if (threadIdx.x == 96) P128_TR(j, 10);
    };
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
    PairTile tc{0, 0, 0}, tn{0, 0, 0};
    bool dummy = false, dummy_n = false;
    // step i = -1 is the prologue E1a(0) E1b(0); every lambda has ONE call site (tp must stay in registers)
    for (int i = -1; i < n_it; ++i) {
      const int t = tc.q0 - h2 + r;
      const bool ok2 = i >= 0 && !dummy && r >= h2 && r < kPairRows - h2 && t < tc.lim;
      const bool more = i + 1 < n_it;
      // The residual of tile i is wanted right after E1b: holding its first chunk in registers across E1a / E1b (64 packed
      // registers live) spills, so it is pulled into L2 here and loaded after E1b.
      if (p.res && ok2) {
        const float4* rp = reinterpret_cast<const float4*>(p.res + (size_t)tc.b * p.o32_bs) + t;
#pragma unroll 8
        for (int q = 0; q < 32; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (size_t)q * p.T));
      }
      if (p.accumulate && ok2) {                                      // ... and what E2 accumulates onto
        const float4* op = reinterpret_cast<const float4*>(p.o32 + (size_t)tc.b * p.o32_bs) + t;
#pragma unroll 8
        for (int q = 0; q < 32; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(op + (size_t)q * p.T));
      }
      if (more) {
        tn = pair_decode(p, pref, lim_s, tile_rt(i + 1, dummy_n), cur);
        E1a(i + 1, tn);                                               // under C2(i)
      }
      if (more) E1b(i + 1);                                           // waits for C2(i)
      // residual: three register buffers, two chunks in flight (one chunk ahead left E2 waiting ~1 us per chunk for L2)
      float4 rr[3][8];
      load_res(tc, 0, ok2, t, rr[0]);                                 // (prefetched into L2 above)
      load_res(tc, 1, ok2, t, rr[1]);
      if (i < 0) {
        tc = tn;
        dummy = dummy_n;
        continue;
      }
      // ---- E2(i): conv2 accumulators -> * acc_scale + b2 + residual, 1/3 mean, fp32 stream + operand planes
      mbar_wait(bar(kWAcc2Full), i & 1);
      tc_fence_after();
      if (threadIdx.x == 96) P128_TR(i, 11);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t a[32];
        __syncwarp();
        tmem_ld32_nowait(lane_addr + 256u + (uint32_t)(cc * 32), a);
        if (cc < 2) load_res(tc, cc + 2, ok2, t, rr[(cc + 2) % 3]);
        const float4 (&rc)[8] = rr[cc % 3];
        tmem_ld_wait();
        if (cc == 3) {
          tc_fence_before();
          arrive_leader(kWAcc2Empty, false);                          // all chunks are in registers
          if (threadIdx.x == 96) P128_TR(i, 12);
        }
        if (ok2) {
          float* v = reinterpret_cast<float*>(a);
          const float4* bv = reinterpret_cast<const float4*>(bias_s + kWC + cc * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bq = bv[q];
            fma2(v[4 * q], v[4 * q + 1], acc_scale, bq.x, bq.y);
            fma2(v[4 * q + 2], v[4 * q + 3], acc_scale, bq.z, bq.w);
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
      }
      if (threadIdx.x == 96 && !(p.csize & 16)) P128_TR(i, 13);
      tc = tn;
      dummy = dummy_n;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // no peer may still arrive on / read this CTA's shared memory
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// packed pair-mode blobs ([chunk][tap][half]{fp16 planes} and [chunk][tap][half] e5m2) -> [half][chunk][tap]{fp16 planes, e5m2}
__global__ void rb128_stream_kernel(const uint4* __restrict__ w, const uint4* __restrict__ w8, uint4* __restrict__ out, int k,
                                    int w16, int w8_16) {          // sizes of one tap's fp16 / e5m2 part in 16-byte units
  const int tap16 = w16 + w8_16;
  const int total = 2 * kWChunks * k * tap16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i % tap16, t = i / tap16;                          // t = (half * chunks + chunk) * k + tap
    const int j = t % k, c = (t / k) % kWChunks, hf = t / (k * kWChunks);
    const size_t blob = (size_t)(c * k + j) * 2 + hf;
    out[i] = e < w16 ? w[blob * w16 + e] : w8[blob * w8_16 + (e - w16)];
  }
}

struct P128Plan { int tg = 0, a_stages = 0, w_stages = 0; size_t smem = 0; };
// Shared-memory plan.  The intermediate tile takes 66-102 KB; what is left is split between the input chunk stages and the
// weight ring.  Measured with the trace build (tools/p128_trace.py): a weight stage lands 0.25 us after it is requested (the
// blobs stay in L2), so three stages of two taps are enough for the weights, while an input chunk (HBM, then the e5m2
// conversion, then the relay to the leader) takes ~2 us: the input gets as many stages as fit, up to all four chunks of a
// tile.  One-tap weight stages cost the MMA thread a barrier round trip per 6 MMAs and are slower.
P128Plan plan128(int k, int dil, int w_planes, int lo8) {
  static const int force_tg = [] { const char* e = getenv("DTTS_TC_P128_TG"); return e ? atoi(e) : 0; }();
  static const int force_as = [] { const char* e = getenv("DTTS_TC_P128_ASTAGES"); return e ? atoi(e) : 0; }();
  static const int force_ws = [] { const char* e = getenv("DTTS_TC_P128_WSTAGES"); return e ? atoi(e) : 0; }();
  const int h2 = (k - 1) / 2, hd = h2 * dil;
  const size_t RA = kPairRows + 2 * hd, RT = kPairRows + 2 * h2;
  const size_t w_tap = (size_t)kWPlaneBytes * w_planes + (lo8 ? kW8Bytes : 0);
  const size_t a_stage = (4 + (lo8 ? 2 : 0)) * RA * 16, t_bytes = (16 + (lo8 ? 8 : 0)) * RT * 16;
  const size_t limit = (size_t)227 * 1024;
  P128Plan pl;
  int tg = force_tg > 0 && force_tg <= k ? force_tg : 2;
  if (tg > k) tg = k;
  tg = cdiv(k, cdiv(k, tg));                                        // balanced groups
  const size_t fixed = (size_t)kWHeader + t_bytes, w_stage = (size_t)tg * w_tap;
  const int min_ws = 3;
  int as = force_as >= 2 && force_as <= kWMaxAStages ? force_as : kWMaxAStages;
  while (as > 2 && !force_as && fixed + as * a_stage + min_ws * w_stage > limit) --as;
  if (fixed + as * a_stage + 2 * w_stage > limit) return pl;
  int ws = (int)((limit - fixed - as * a_stage) / w_stage);
  if (ws > kWMaxWStages) ws = kWMaxWStages;
  if (force_ws >= 2 && force_ws < ws) ws = force_ws;
  pl.tg = tg; pl.a_stages = as; pl.w_stages = ws;
  pl.smem = fixed + as * a_stage + ws * w_stage;
  return pl;
}

}  // namespace

cudaError_t rb_pair128_pack_stream(const TcConvW& w, uint8_t* out, cudaStream_t s) {
  if (w.C_in != kWC || w.C_out != kWC || !w.pair || w.KC != 32 || w.N != kWC || (w.lo8 && !w.w8)) return cudaErrorInvalidValue;
  const int w16 = kWPlaneBytes * w.planes / 16, w8_16 = w.lo8 ? kW8Bytes / 16 : 0;
  const int total = 2 * kWChunks * w.ktaps * (w16 + w8_16);
  rb128_stream_kernel<<<(total + 255) / 256, 256, 0, s>>>(reinterpret_cast<const uint4*>(w.w), reinterpret_cast<const uint4*>(w.w8),
                                                          reinterpret_cast<uint4*>(out), w.ktaps, w16, w8_16);
  return cudaGetLastError();
}

#ifdef DTTS_P128_TRACE
extern "C" int dtts_debug_p128_trace_select(int k, int dil) {
  const int sel = k * 16 + dil;
  return (int)cudaMemcpyToSymbol(g_p128_trace_sel, &sel, sizeof(int));
}
extern "C" int dtts_debug_p128_trace(long long* out, int n) {
  if (n > kTrTiles * kTrEvents) n = kTrTiles * kTrEvents;
  return (int)cudaMemcpyFromSymbol(out, g_p128_trace, (size_t)n * sizeof(long long));
}
#endif

int rb_pair128_supported(const TcConvW& c1, const TcConvW& c2, int dil, int a_planes) {
  if (c1.ktaps > tc_fuse128_maxk()) return 0;
  if (c1.C_in != kWC || c1.C_out != kWC || c2.C_in != kWC || c2.C_out != kWC) return 0;
  if (c1.ktaps != c2.ktaps || !(c1.ktaps & 1) || c1.ktaps > 11 || dil < 1) return 0;
  if (!c1.pair || !c2.pair || c1.stack || c2.stack || c1.N != kWC || c2.N != kWC || c1.KC != 32 || c2.KC != 32 ||
      c1.il_u || c2.il_u || c1.phases != 1 || c2.phases != 1 || a_planes != 1 || c1.fmt != c2.fmt ||
      c1.planes != c2.planes || c1.lo8 != c2.lo8)
    return 0;
  if (c1.lo8 && (c1.planes != 1 || c1.fmt != 0)) return 0;
  if (!c1.wstream || !c2.wstream) return 0;
  const int h2 = (c1.ktaps - 1) / 2;
  if (h2 + h2 * dil > TC_PADF) return 0;                              // the conv1 halo of the first tile starts inside the front padding
  return plan128(c1.ktaps, dil, c1.planes, c1.lo8).tg > 0;
}

cudaError_t launch_rb_pair128(RbPairParams p, cudaStream_t stream) {
  if (p.B <= 0 || p.T <= 0) return cudaSuccess;
  if (p.lens && p.B > TC_MAX_RAGGED_ITEMS) return cudaErrorInvalidValue;
  if (p.C != kWC || p.post_part || p.w_planes < 1 || p.w_planes > 2 || !p.w1s || !p.w2s || (p.lo8 && p.w_planes != 1))
    return cudaErrorInvalidValue;
  const int h2 = (p.k - 1) / 2, hd = h2 * p.dil;
  p.S = kPairRows - 2 * h2;
  p.ntiles = cdiv(p.T, p.S);
  if (p.a_pad - h2 - hd < 0 || p.a_pad + (p.ntiles - 1) * p.S - h2 + kPairRows + hd > p.a_rows) return cudaErrorInvalidValue;
  const P128Plan pl = plan128(p.k, p.dil, p.w_planes, p.lo8);
  if (!pl.tg) return cudaErrorInvalidConfiguration;
  p.TG = pl.tg; p.a_stages = pl.a_stages; p.w_stages = pl.w_stages; p.csize = 2;
  static const int pf = [] { const char* e = getenv("DTTS_TC_P128_PREFETCH"); return e ? atoi(e) : 0; }();   // measured: no gain
  p.pf = pf;
#ifdef DTTS_P128_TRACE
  if (getenv("DTTS_P128_NOWAITW")) p.csize |= 4;        // experiments on the MMA thread's weight waits (garbage results)
  if (getenv("DTTS_P128_NORELAYW")) p.csize |= 8;
  if (getenv("DTTS_P128_COPYLAT")) p.csize |= 16;       // event 13 becomes the issue -> landed time of one weight copy per tile
  if (getenv("DTTS_P128_NOEPI")) { p.res = nullptr; p.o32 = nullptr; p.o_hi = nullptr; p.accumulate = 0; }   // experiment: no epilogue HBM traffic
#endif
  static unsigned long long attr_done = 0;
  const cudaError_t attr_err = ensure_max_dyn_smem(rb_pair128_kernel, 227 * 1024, &attr_done);
  if (attr_err != cudaSuccess) return attr_err;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  // the CTA owns all 512 TMEM columns: it must be alone on its SM (shared memory above half of the SM's guarantees it)
  const size_t smem_launch = pl.smem < 116 * 1024 ? 116 * 1024 : pl.smem;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  cfg.blockDim = dim3(kPairThreads);
  cfg.dynamicSmemBytes = smem_launch;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim = {2, 1, 1};
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = tc_pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  static std::mutex mu;
  static int max_pairs[64] = {0};                // co-resident CTA pairs, queried once per device
  {
    std::lock_guard<std::mutex> lock(mu);
    if (max_pairs[dev] == 0) {
      int n = 0;
      cfg.gridDim = dim3((unsigned)(sms / 2 * 2));
      cfg.dynamicSmemBytes = 227 * 1024;
      const cudaError_t q = cudaOccupancyMaxActiveClusters(&n, rb_pair128_kernel, &cfg);
      cfg.dynamicSmemBytes = smem_launch;
      if (q != cudaSuccess) return q;
      max_pairs[dev] = n > 0 ? n : -1;
    }
    if (max_pairs[dev] <= 0) return cudaErrorInvalidConfiguration;
  }
  const long units = ((long)p.ntiles * p.B + 1) / 2;
  const int npairs = (int)(units < max_pairs[dev] ? units : max_pairs[dev]);
  cfg.gridDim = dim3((unsigned)(npairs * 2));
  return cudaLaunchKernelEx(&cfg, rb_pair128_kernel, p);
}

}  // namespace dtts
