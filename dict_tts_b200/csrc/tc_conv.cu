// tcgen05 implicit-GEMM convolution (see tc_conv.cuh for the layout and the math).
//
// CTA = 224 threads, warp-specialised:
//   warp 0 (1 lane)  activation producer : cp.async.bulk of (tile + halo) rows, one per 8-channel slab and plane
//   warp 1 (1 lane)  weight producer     : cp.async.bulk of one pre-packed blob per (K chunk, tap)
//   warp 2           TMEM alloc/free; 1 lane issues tcgen05.mma (M=128, N=C_out, K=16) into NACC accumulators
//   warps 3..6       epilogue: tcgen05.ld -> bias / residual / mean scaling -> fp32 stream + leaky-ReLU'd bf16 hi/lo planes
// Pipelines: a_full/a_empty (activation stages), w_full/w_empty (weight stages), acc_full (MMA -> epilogue); the
// shared-memory slots are released by tcgen05.commit.
//
// Thread-block clusters: the CTAs of a cluster work on consecutive time tiles of the SAME weight group in lockstep,
// so every weight blob is fetched from L2 once per cluster: CTA r loads slice r of the blob and multicasts it into the
// shared memory of all its peers (cp.async.bulk ... .multicast::cluster); a weight slot is re-filled only after the
// MMA warps of ALL peers have released it (tcgen05.commit ... .multicast::cluster onto every peer's w_empty barrier).
// Without this the weight stream (C_out*C_in*K*2 B per 128-row tile) makes the big layers L2-bandwidth bound.
#include "tc_conv.cuh"
#include "tc16.cuh"
#include "tc_ptx.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

namespace dtts {

namespace {

constexpr int kEpiWarps = 8;
constexpr int kThreads = 96 + kEpiWarps * 32;
constexpr int kMaxAStages = 4, kMaxWStages = 6;
// mbarrier slots (8 B each) at the start of shared memory
constexpr int kBarAFull = 0, kBarAEmpty = kBarAFull + kMaxAStages, kBarWFull = kBarAEmpty + kMaxAStages,
              kBarWEmpty = kBarWFull + kMaxWStages, kBarAccFull = kBarWEmpty + kMaxWStages, kBarAccEmpty = kBarAccFull + 2,
              kBarPAFull = kBarAccEmpty + 2, kBarPWFull = kBarPAFull + kMaxAStages, kBarA8Ready = kBarPWFull + kMaxWStages,
              kNumBars = kBarA8Ready + kMaxAStages;
constexpr int kTmemPtrOff = kNumBars * 8;                 // uint32: TMEM base address
constexpr int kBiasOff = (kTmemPtrOff + 4 + 63) / 64 * 64;
constexpr int kMaxBias = 2048;               // output channels of one launch (bias staged in shared memory)
constexpr int kPrefOff = kBiasOff + kMaxBias * 4;      // ragged launches: int32 tile prefix [kMaxItems + 1], then limits [kMaxItems]
constexpr int kMaxItems = TC_MAX_RAGGED_ITEMS;
constexpr int kSmemHeader = kPrefOff + (2 * kMaxItems + 8) * 4;   // barriers + tmem ptr, bias, ragged tables
constexpr int kSmemLimit = 227 * 1024;


// Work decode.  A unit = csize consecutive row tiles (row tile rt = b * ntiles + time tile) of one output-channel
// block; the CTA of cluster rank r takes row tile (unit % nu) * csize + r.  A cluster walks its units
// (cluster_id, cluster_id + nclusters, ...) and, inside each unit, ALL polyphase components of a transposed convolution
// back to back (weight group g = nblock*phases + phase): the interleaved output samples t = q*stride + phase of one row
// tile are then written by the same SM within a few microseconds and merge into full sectors in L2 instead of
// reaching HBM as partial-sector read-modify-writes.  Row tiles past the end are dummies: they follow the weight
// pipeline (the peers depend on it) but issue no MMAs and store nothing.
struct TileCoord { int q0, g, b, lim; bool dummy; };   // lim: rows [0, lim) of item b are computed / stored by this launch
// Row-tile schedule of a launch.  Uniform: every batch item has p.ntiles tiles.  Ragged (p.lens != null): item b has
// ceil(lim_b / MT) tiles with lim_b = min(nq, lens[b] * len_mul + len_add); pref[] (shared memory) is the exclusive
// prefix of the tile counts, so the persistent CTAs walk ONLY tiles that hold rows somebody needs.
struct Sched {
  const int* pref;     // [B + 1] in shared memory, null when uniform
  const int* limv;     // [B] in shared memory
  int nu, nrt;         // units per weight group, row tiles in total
};
// Each role walks its tiles in increasing order, so the batch item of a row tile is found by advancing a cursor
// (no division / search on the per-tile path: for the short k = 3 tiles the tile decode of the MMA-issuing thread is on the
// critical path).
struct Cursor { int b = 0; uint32_t base = 0; };
__device__ __forceinline__ TileCoord decode_unit(const TcConvParams& p, const Sched& sc, int it, int cluster_id,
                                                 int nclusters, int rank, Cursor& cur) {
  TileCoord c;
  uint32_t step = (uint32_t)it, ph = 0;
  if (p.phases != 1) { step = (uint32_t)it / (uint32_t)p.phases; ph = (uint32_t)it - step * (uint32_t)p.phases; }
  const uint32_t unit = (uint32_t)cluster_id + step * (uint32_t)nclusters;
  uint32_t blk = 0, j = unit;
  if (p.nblocks != 1) { blk = unit / (uint32_t)sc.nu; j = unit - blk * (uint32_t)sc.nu; }
  c.g = (int)(blk * (uint32_t)p.phases + ph);
  uint32_t rt = j * (uint32_t)p.csize + (uint32_t)rank;
  const uint32_t nrt = (uint32_t)sc.nrt;
  c.dummy = rt >= nrt;
  if (c.dummy) rt = nrt - 1;
  if (sc.pref) {
    int b = cur.b;
    if (rt < (uint32_t)sc.pref[b]) b = 0;                      // next weight group: the walk starts over
    while (b + 1 < p.B && (uint32_t)sc.pref[b + 1] <= rt) ++b;   // largest b with pref[b] <= rt (empty items are skipped)
    cur.b = b;
    c.b = b;
    c.q0 = (int)(rt - (uint32_t)sc.pref[b]) * p.MT;
    c.lim = sc.limv[b];
  } else {
    const uint32_t nt = (uint32_t)p.ntiles;
    if (rt < cur.base) { cur.b = 0; cur.base = 0; }
    while (rt >= cur.base + nt) { cur.base += nt; ++cur.b; }
    c.b = cur.b;
    c.q0 = (int)(rt - cur.base) * p.MT;
    c.lim = p.nq;
  }
  return c;
}
// 128-row sub-tiles of a tile that hold rows below the item's limit (the MMAs of the others are not issued)
__device__ __forceinline__ int tile_nacc(const TcConvParams& p, const TileCoord& c) {
  if (c.dummy) return 0;
  const int n = (c.lim - c.q0 + 127) >> 7;
  return n < p.NACC ? n : p.NACC;
}

// Everything the MMA-issuing warp needs, precomputed once (descriptor low words are in units of 16 B).
struct MmaCtx {
  uint32_t bar0, tmem_base;
  uint32_t a_low0, b_low0;          // start >> 4 | LBO >> 4 << 16 of stage 0
  uint32_t a_stage16, w_blob16;     // stage strides
  uint32_t w_tap16;                 // one tap inside a weight stage
  uint32_t a_kstep, b_kstep;        // +K step (two 8-channel slabs)
  uint32_t a_plane, b_plane;        // + lo plane
  uint32_t idesc, idesc_n;          // instruction descriptors with N = NM (main) and N = p.N (a_lo x w_hi when stacked)
  uint32_t a8_delta, b8_low0, w8_tap16, idesc8;   // lo8: e5m2 copy of the A tile (offset inside a stage), lo weights
  int cluster_id, nclusters, n_it, rank, csize;
  Sched sc;
  uint16_t cmask;
  int pair;                         // 1: cta_group::2 -- this thread is the leader of a CTA pair
};

// The MMA-issuing warp.  One elected lane issues; the loop is instantiated per operand mode so that the single issuing
// thread -- the serial resource of the CTA -- spends a handful of uniform-datapath instructions per tcgen05.mma.
//   WMODE 0: one weight plane; 1: hi and lo planes, one MMA each; 2: hi | lo stacked along N (one MMA, the epilogue adds
//   the two column halves) -- an M=128,K=16 MMA costs max(64, N/2) cycles (A is read from shared memory at 64 B/clk), so
//   for C_out <= 64 the second weight plane is free this way; 3: fp16 hi plane + ONE FP8 MMA (K = 32) per tap for the lo
//   plane (TcMode::lo8, KC = 32).
template <int KSTEPS, int APL, int WMODE, bool PAIR>
__device__ __forceinline__ void mma_warp_loop(const TcConvParams& p, const MmaCtx& x) {
  // PAIR: tcgen05.mma.cta_group::2 issued by the leader CTA only; the peer's warp 2 relays its "stage full" events to the
  // leader's pa_full / pw_full barriers ([20..21], [22..27]) and every commit is multicast to both CTAs.
  auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (PAIR) umma_bf16_2cta(d, a, b, idesc, acc);
    else umma_bf16(d, a, b, idesc, acc);
  };
  auto mma8 = [](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (PAIR) umma_f8_2cta(d, a, b, idesc, acc);
    else umma_f8(d, a, b, idesc, acc);
  };
  if (elect_one()) {                                 // ONE thread runs the whole loop: waits, MMAs and commits
    const uint32_t hiw = (128u >> 4) | (1u << 14);   // SBO = 128 B, descriptor version 1
    const int ktaps = p.ktaps, nchunks = p.nchunks, tap_step = p.tap_step, w_stages = p.w_stages, a_stages = p.a_stages;
    const int NM = p.NM, TG = p.TG;
    const uint32_t tap0 = (uint32_t)(p.tap_off0 - p.min_off);
    auto bar = [&](int i) { return x.bar0 + 8u * (uint32_t)i; };  // index = kBar* + stage
    int sa = 0, sw = 0;
    uint32_t pa = 0, pw = 0;                         // stage cursors + phase parities (no div/mod in this loop)
    int t_it = 0;
    Cursor cur;
    for (int it = 0; it < x.n_it; ++it, ++t_it) {
      int nacc = tile_nacc(p, decode_unit(p, x.sc, it, x.cluster_id, x.nclusters, x.rank, cur));
      if (PAIR) nacc = max(nacc, tile_nacc(p, decode_unit(p, x.sc, it, x.cluster_id, x.nclusters, 1, cur)));   // M = 256 spans both tiles
      const int as = t_it & 1;
      mbar_wait(bar(kBarAccEmpty + as), ((t_it >> 1) & 1) ^ 1);  // acc_empty: the epilogue(s) drained this accumulator set
      tc_fence_after();
      const uint32_t d_base = x.tmem_base + (uint32_t)(as * 256);
      for (int c = 0; c < nchunks; ++c) {
        // a_full -- with lo8 "the e5m2 copy of the tile is in place", which the producer warp signals after a_full
        mbar_wait(bar((WMODE == 3 ? kBarA8Ready : kBarAFull) + sa), pa);
        if (PAIR) mbar_wait(bar(kBarPAFull + sa), pa);       // ... and the peer's A tile
        tc_fence_after();
        uint32_t a_tap = x.a_low0 + (uint32_t)sa * x.a_stage16 + tap0;
        for (int j0 = 0; j0 < ktaps; j0 += TG) {     // one weight stage = TG consecutive taps of this K chunk
          mbar_wait(bar(kBarWFull + sw), pw);                // w_full
          if (PAIR) mbar_wait(bar(kBarPWFull + sw), pw);     // ... and the peer's half of the weights
          tc_fence_after();
          uint32_t b_lo = x.b_low0 + (uint32_t)sw * x.w_blob16;
          uint32_t b8 = x.b8_low0 + (uint32_t)sw * x.w_blob16;
          const int j1 = min(j0 + TG, ktaps);
          for (int j = j0; j < j1; ++j, a_tap += (uint32_t)tap_step, b_lo += x.w_tap16, b8 += x.w8_tap16) {
            const uint32_t first = (c | j) != 0 ? 1u : 0u;
            uint32_t d = d_base, am = a_tap;
            for (int m = 0; m < nacc; ++m, d += (uint32_t)NM, am += 128u) {
#pragma unroll
              for (int ks = 0; ks < KSTEPS; ++ks) {
                const uint64_t a = desc64(am + ks * x.a_kstep, hiw), b = desc64(b_lo + ks * x.b_kstep, hiw);
                mma(d, a, b, x.idesc, ks == 0 ? first : 1u);
                if (WMODE == 1) mma(d, a, desc64(b_lo + ks * x.b_kstep + x.b_plane, hiw), x.idesc, 1u);
                if (APL == 2)
                  mma(d, desc64(am + ks * x.a_kstep + x.a_plane, hiw), b, WMODE == 2 ? x.idesc_n : x.idesc, 1u);
              }
              // lo plane at the FP8 rate: the chunk's 32 channels in one K = 32 instruction
              if (WMODE == 3) mma8(d, desc64(am + x.a8_delta, hiw), desc64(b8, hiw), x.idesc8, 1u);
            }
          }
          if (PAIR) umma_commit_2cta(bar(kBarWEmpty + sw));                   // w_empty of both CTAs
          else if (x.csize > 1) umma_commit_mc(bar(kBarWEmpty + sw), x.cmask);   // w_empty here and at every peer's producer
          else umma_commit(bar(kBarWEmpty + sw));
          if (++sw == w_stages) { sw = 0; pw ^= 1u; }
        }
        if (PAIR) {
          umma_commit_2cta(bar(kBarAEmpty + sa));                              // a_empty of both CTAs
          if (c == nchunks - 1) umma_commit_2cta(bar(kBarAccFull + as));       // acc_full of both CTAs
        } else {
          umma_commit(bar(kBarAEmpty + sa));                                   // a_empty
          if (c == nchunks - 1) umma_commit(bar(kBarAccFull + as));            // acc_full
        }
        if (++sa == a_stages) { sa = 0; pa ^= 1u; }
      }
    }
  }
  __syncwarp();
}

// PAIR = true is the cta_group::2 build of the same kernel (launched as clusters of 2 only: a cubin that contains
// cta_group::2 instructions cannot be launched without a cluster).
template <bool PAIR>
__global__ void __launch_bounds__(kThreads, 1) tc_conv_kernel(const TcConvParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N, NM = p.NM, KC = p.KC, APL = p.a_planes, WPL = p.w_planes;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  // mbarriers: see kBar* (a_full/a_empty per activation stage, w_full/w_empty per weight stage, acc_full/acc_empty per
  // accumulator set; pair mode, leader only: pa_full / pw_full = "the peer's stage is full")
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * (kBarAFull + s); };
  auto a_empty = [&](int s) { return bar0 + 8u * (kBarAEmpty + s); };
  auto w_full = [&](int s) { return bar0 + 8u * (kBarWFull + s); };
  auto w_empty = [&](int s) { return bar0 + 8u * (kBarWEmpty + s); };
  auto acc_full = [&](int s) { return bar0 + 8u * (kBarAccFull + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (kBarAccEmpty + s); };
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + kTmemPtrOff);
  float* bias_s = reinterpret_cast<float*>(smem + kBiasOff);          // [nblocks * N] <= kMaxBias floats

  const uint32_t a_plane_bytes = (uint32_t)(KC / 8) * p.RA * 16u;
  const uint32_t a8_bytes = p.lo8 ? (uint32_t)(KC / 16) * p.RA * 16u : 0u;      // e5m2 copy of the tile (lo8)
  const uint32_t a_stage_bytes = a_plane_bytes * APL + a8_bytes;
  constexpr int pair = PAIR ? 1 : 0;                            // CTA pair: each CTA stages half of every weight blob
  const uint32_t w_plane_bytes = (uint32_t)(pair ? NM / 2 : NM) * KC * 2u;   // NM = 2N when hi | lo are stacked (WPL = 1)
  const uint32_t w_tap_bytes = w_plane_bytes * WPL;             // one (K chunk, tap) blob
  const uint32_t w8_tap_bytes = p.lo8 ? w_plane_bytes / 2u : 0u;               // its e5m2 lo plane (lo8)
  // one weight stage = TG consecutive taps: [TG fp16 blobs][TG e5m2 blobs]
  const uint32_t w_blob_bytes = (w_tap_bytes + w8_tap_bytes) * (uint32_t)p.TG;
  const uint32_t a_base = smem_u32(smem + kSmemHeader);
  const uint32_t w_base = a_base + p.a_stages * a_stage_bytes;
  const int csize = p.csize;
  const int rank = csize > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / csize, nclusters = gridDim.x / csize;
  const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
  int* pref_s = reinterpret_cast<int*>(smem + kPrefOff);              // ragged launches only
  int* lim_s = pref_s + kMaxItems + 1;
  griddep_launch();                            // the next launch may set itself up while this one runs (common.cuh)
  if (p.lens && warp == 3) {
    griddep_wait();                            // lens is an input of the pass: written by whatever ran before
    // per-item row limits and the exclusive prefix of their tile counts (B <= kMaxItems, checked by the launcher)
    int carry = 0;
    for (int b0 = 0; b0 < p.B; b0 += 32) {
      const int b = b0 + lane;
      int lim = 0;
      if (b < p.B) {
        const long v = (long)__ldg(p.lens + b) * p.len_mul + p.len_add;
        lim = v < 0 ? 0 : (v > p.nq ? p.nq : (int)v);
        lim_s[b] = lim;
      }
      int nt = (lim + p.MT - 1) / p.MT, inc = nt;
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
    for (int s = 0; s < kMaxAStages; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < kMaxWStages; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), pair ? 1 : csize); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), pair ? 2 * kEpiWarps : kEpiWarps); }
    for (int s = 0; s < kMaxAStages + kMaxWStages; ++s) mbar_init(bar0 + 8u * (kBarPAFull + s), 1);
    for (int s = 0; s < kMaxAStages; ++s) mbar_init(bar0 + 8u * (kBarA8Ready + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem + kTmemPtrOff)),
                   "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem + kTmemPtrOff)),
                   "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();           // every peer's barriers are initialised before any remote arrive / copy
  tc_fence_after();
  // Programmatic dependent launch: everything up to here (barriers, TMEM, and -- for warp 1 -- the first weight stages,
  // which are constants) overlaps the tail of the kernel in front; activations, residuals and outputs are only touched
  // after the predecessor grid has completed.
  if (warp != 1) griddep_wait();
  const uint32_t tmem_base = *tmem_ptr_s;
  Sched sc;
  sc.pref = p.lens ? pref_s : nullptr;
  sc.limv = lim_s;
  sc.nrt = p.lens ? pref_s[p.B] : p.ntiles * p.B;
  sc.nu = p.lens ? (sc.nrt + csize - 1) / csize : p.nu;
  const int total_units = sc.nu * p.nblocks;                     // x phases each
  // (unit, phase) steps of this cluster; a ragged launch may leave a cluster without work
  const int n_it = total_units > cluster_id ? (total_units - cluster_id + nclusters - 1) / nclusters * p.phases : 0;

  if (warp == 0) {
    // ------------------------------------------------ activation producer
    const int slabs = KC / 8;
    const uint32_t row_bytes = (uint32_t)p.RA * 16u;
    int s = 0;
    uint32_t ph = 1;                                             // producers start on the "previous phase done" parity
    Cursor cur;
    // lo8: the e5m2 copy of a staged tile is made HERE, in shared memory: e5m2 is fp16 with the mantissa cut to 2 bits, i.e.
    // the high byte of every element, so no second plane ever crosses HBM.  The whole warp converts stage (cs, cpar) once
    // its bulk copies have landed -- before it blocks on the next free slot, i.e. while the previous chunk's MMAs run.
    int cs = 0, pend = 0;
    uint32_t cpar = 0;
    auto convert_stage = [&]() {
      mbar_wait(a_full(cs), cpar);
      const uint32_t src = a_base + cs * a_stage_bytes, dst = src + APL * a_plane_bytes;
      const uint32_t slab = (uint32_t)p.RA * 16u;                // bytes of one 8-channel fp16 slab = one e5m2 slab
      for (int s8 = 0; s8 < KC / 16; ++s8) {                     // e5m2 slab s8 = high bytes of fp16 slabs 2*s8, 2*s8 + 1
        const uint32_t s0 = src + (uint32_t)(2 * s8) * slab, d0 = dst + (uint32_t)s8 * slab;
#pragma unroll 4
        for (int r = lane; r < p.RA; r += 32) {
          uint4 lo, hi;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "r"(s0 + (uint32_t)r * 16u));
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "r"(s0 + slab + (uint32_t)r * 16u));
          const uint32_t o0 = __byte_perm(lo.x, lo.y, 0x7531), o1 = __byte_perm(lo.z, lo.w, 0x7531);
          const uint32_t o2 = __byte_perm(hi.x, hi.y, 0x7531), o3 = __byte_perm(hi.z, hi.w, 0x7531);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d0 + (uint32_t)r * 16u), "r"(o0), "r"(o1),
                       "r"(o2), "r"(o3)
                       : "memory");
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 8u * (kBarA8Ready + cs));
      if (++cs == p.a_stages) { cs = 0; cpar ^= 1u; }
      --pend;
    };
    for (int it = 0; it < n_it; ++it) {
      const TileCoord tc = decode_unit(p, sc, it, cluster_id, nclusters, rank, cur);
      const size_t row0 = (size_t)(p.a_pad + tc.q0 + p.min_off);
      for (int c = 0; c < p.nchunks; ++c) {
        if (p.lo8 && pend > 0) convert_stage();
        mbar_wait(a_empty(s), ph);
        if (elect_one()) {
          mbar_arrive_expect_tx(a_full(s), a_plane_bytes * APL);
          for (int pl = 0; pl < APL; ++pl) {
            const tc16* src = (pl ? p.a_lo : p.a_hi) + (size_t)tc.b * p.a_bs;
            for (int sl = 0; sl < slabs; ++sl) {
              const tc16* gp = src + ((size_t)(c * slabs + sl) * p.a_rows + row0) * 8;
              bulk_g2s(a_base + s * a_stage_bytes + pl * a_plane_bytes + sl * row_bytes, gp, row_bytes, a_full(s));
            }
          }
        }
        __syncwarp();
        ++pend;
        if (++s == p.a_stages) { s = 0; ph ^= 1u; }
      }
    }
    while (p.lo8 && pend > 0) convert_stage();
  } else if (warp == 1) {
    // ------------------------------------------------ weight producer
    const size_t tap_elems = (size_t)w_tap_bytes / 2;
    const int per_tile = p.nchunks * p.ktaps;
    const uint32_t tap_slice = w_tap_bytes / (uint32_t)csize;         // this CTA's share of every stage, per tap in it
    int s = 0;
    uint32_t ph = 1;
    Cursor cur;
    for (int it = 0; it < n_it; ++it) {
      const TileCoord tc = decode_unit(p, sc, it, cluster_id, nclusters, rank, cur);
      const tc16* wg = p.w + (size_t)tc.g * per_tile * tap_elems * (pair ? 2 : 1);
      for (int c = 0; c < p.nchunks; ++c) {
        for (int j0 = 0; j0 < p.ktaps; j0 += p.TG) {
          const uint32_t ntap = (uint32_t)min(p.TG, p.ktaps - j0);
          mbar_wait(w_empty(s), ph);                                  // released by the MMA warps of ALL peers
          if (pair) {
            // CTA pair: the packed weights hold [tap][half][...]; this CTA stages half `rank` of every tap of the group
            if (elect_one()) {
              mbar_arrive_expect_tx(w_full(s), ntap * (w_tap_bytes + w8_tap_bytes));
              for (uint32_t jj = 0; jj < ntap; ++jj)
                bulk_g2s(w_base + s * w_blob_bytes + jj * w_tap_bytes,
                         wg + ((size_t)(c * p.ktaps + j0 + (int)jj) * 2 + rank) * tap_elems, w_tap_bytes, w_full(s));
              if (p.lo8) {
                const uint8_t* wg8 = p.w8 + (size_t)tc.g * per_tile * w8_tap_bytes * 2;
                for (uint32_t jj = 0; jj < ntap; ++jj)
                  bulk_g2s(w_base + s * w_blob_bytes + (uint32_t)p.TG * w_tap_bytes + jj * w8_tap_bytes,
                           wg8 + ((size_t)(c * p.ktaps + j0 + (int)jj) * 2 + rank) * w8_tap_bytes, w8_tap_bytes,
                           w_full(s));
              }
            }
          } else if (elect_one()) {
            mbar_arrive_expect_tx(w_full(s), ntap * (w_tap_bytes + w8_tap_bytes));   // the whole stage lands here, one slice per peer
            if (p.lo8) {                                                // (lo8 launches are never multicast: csize == 1)
              const uint8_t* wg8 = p.w8 + (size_t)tc.g * per_tile * w8_tap_bytes;
              bulk_g2s(w_base + s * w_blob_bytes + (uint32_t)p.TG * w_tap_bytes,
                       wg8 + (size_t)(c * p.ktaps + j0) * w8_tap_bytes, ntap * w8_tap_bytes, w_full(s));
            }
            // the ntap tap blobs are contiguous in global and in shared memory: peer r copies bytes [r, r+1) * slice
            const uint32_t slice = ntap * tap_slice;
            const uint32_t dst = w_base + s * w_blob_bytes + rank * slice;
            const tc16* src = wg + (size_t)(c * p.ktaps + j0) * tap_elems + (size_t)rank * (slice / 2);
            if (csize > 1) bulk_g2s_mc(dst, src, slice, w_full(s), cmask);
            else bulk_g2s(dst, src, slice, w_full(s));
          }
          __syncwarp();
          if (++s == p.w_stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------ MMA issuer (converged warp, one elected lane issues)
    MmaCtx x;
    // instruction descriptor: D=f32 @4, A/B format @7/@10 (0 = f16, 1 = bf16), both K-major, N>>3 @17, M>>4 @24
    const uint32_t f16b = p.fmt ? 1u : 0u;
    const uint32_t ibase = (1u << 4) | (f16b << 7) | (f16b << 10) | ((128u >> 4) << 24);
    x.idesc = ibase | ((uint32_t)(NM >> 3) << 17);
    x.idesc_n = ibase | ((uint32_t)(N >> 3) << 17);
    if (pair) {                                                         // M = 256 across the CTA pair
      x.idesc = (x.idesc & ~(0x1Fu << 24)) | ((256u >> 4) << 24);
      x.idesc_n = (x.idesc_n & ~(0x1Fu << 24)) | ((256u >> 4) << 24);
    }
    x.bar0 = bar0; x.tmem_base = tmem_base;
    x.a_low0 = ((a_base >> 4) & 0x3FFFu) | ((uint32_t)p.RA << 16);      // LBO = RA * 16 B
    x.b_low0 = ((w_base >> 4) & 0x3FFFu) | ((uint32_t)(pair ? NM / 2 : NM) << 16);   // LBO = rows of B in this CTA * 16 B
    x.a_stage16 = a_stage_bytes >> 4; x.w_blob16 = w_blob_bytes >> 4; x.w_tap16 = w_tap_bytes >> 4;
    x.a_kstep = 2u * (uint32_t)p.RA; x.b_kstep = 2u * (uint32_t)(pair ? NM / 2 : NM);
    x.a_plane = a_plane_bytes >> 4; x.b_plane = w_plane_bytes >> 4;
    // lo8: e5m2 operands, D = f32 @4, A/B format E5M2 = 1 @7/@10, same M / N fields
    x.a8_delta = (a_plane_bytes * (uint32_t)APL) >> 4;
    x.b8_low0 = (((w_base + (uint32_t)p.TG * w_tap_bytes) >> 4) & 0x3FFFu) | ((uint32_t)(pair ? NM / 2 : NM) << 16);
    x.w8_tap16 = w8_tap_bytes >> 4;
    x.idesc8 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NM >> 3) << 17) | ((pair ? 256u >> 4 : 128u >> 4) << 24);
    x.cluster_id = cluster_id; x.nclusters = nclusters; x.n_it = n_it; x.rank = rank; x.csize = csize;
    x.sc = sc;
    x.cmask = cmask;
    x.pair = pair;
    const int wmode = p.stack ? 2 : (WPL == 2 ? 1 : 0);
    if (PAIR && rank != 0) {
      // ---- peer of a CTA pair: no MMAs are issued here; forward "my stage is full" to the leader, in pipeline order
      if (elect_one()) {
        int sa = 0, sw = 0;
        uint32_t pa = 0, pw = 0;
        for (int it = 0; it < n_it; ++it) {
          for (int c = 0; c < p.nchunks; ++c) {
            if (p.lo8) {                            // the peer's tile counts once ITS e5m2 copy is written (generic proxy:
              mbar_wait(bar0 + 8u * (kBarA8Ready + sa), pa);      // release it at cluster scope)
              mbar_arrive_remote_release(bar0 + 8u * (kBarPAFull + sa), 0);
            } else {
              mbar_wait(a_full(sa), pa);
              mbar_arrive_remote(bar0 + 8u * (kBarPAFull + sa), 0);
            }
            for (int j0 = 0; j0 < p.ktaps; j0 += p.TG) {
              mbar_wait(w_full(sw), pw);
              mbar_arrive_remote(bar0 + 8u * (kBarPWFull + sw), 0);
              if (++sw == p.w_stages) { sw = 0; pw ^= 1u; }
            }
            if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
          }
        }
      }
      __syncwarp();
    } else if (p.lo8) {                              // fp16 hi plane + FP8 lo plane (KC = 32, single-plane activations)
      mma_warp_loop<2, 1, 3, PAIR>(p, x);
    } else if constexpr (PAIR) {
      switch ((KC == 32 ? 2 : 0) + (WPL == 2 ? 1 : 0)) {      // pair mode: single-plane activations, unstacked weights
        case 0: mma_warp_loop<1, 1, 0, true>(p, x); break;
        case 1: mma_warp_loop<1, 1, 1, true>(p, x); break;
        case 2: mma_warp_loop<2, 1, 0, true>(p, x); break;
        default: mma_warp_loop<2, 1, 1, true>(p, x); break;
      }
    } else {
      switch ((KC == 32 ? 6 : 0) + (APL == 2 ? 3 : 0) + wmode) {
        case 0: mma_warp_loop<1, 1, 0, false>(p, x); break;
        case 1: mma_warp_loop<1, 1, 1, false>(p, x); break;
        case 2: mma_warp_loop<1, 1, 2, false>(p, x); break;
        case 3: mma_warp_loop<1, 2, 0, false>(p, x); break;
        case 4: mma_warp_loop<1, 2, 1, false>(p, x); break;
        case 5: mma_warp_loop<1, 2, 2, false>(p, x); break;
        case 6: mma_warp_loop<2, 1, 0, false>(p, x); break;
        case 7: mma_warp_loop<2, 1, 1, false>(p, x); break;
        case 8: mma_warp_loop<2, 1, 2, false>(p, x); break;
        case 9: mma_warp_loop<2, 2, 0, false>(p, x); break;
        case 10: mma_warp_loop<2, 2, 1, false>(p, x); break;
        default: mma_warp_loop<2, 2, 2, false>(p, x); break;
      }
    }
  } else {
    // ------------------------------------------------ epilogue (warps 3..10; two warps per TMEM lane quadrant)
    const int et = threadIdx.x - 96;                 // 0..255
    const int quad = warp & 3;                       // TMEM lanes [32*quad, 32*quad+32) are visible to this warp
    const int half = (warp - 3) >> 2;
    const int nbias = p.il_u ? p.il_cb * p.nblocks : N * p.nblocks;
    for (int i = et; i < nbias; i += kEpiWarps * 32) bias_s[i] = p.bias ? __ldg(p.bias + i) : 0.f;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const bool split = p.o_lo != nullptr;
    const int fmt = p.fmt;
    const float acc_scale = p.acc_scale;
    const int ncc = N / 32;
    const int nitems = p.NACC * ncc;
    int t_it = 0;
    Cursor cur;
    for (int it = 0; it < n_it; ++it, ++t_it) {
      const TileCoord tc = decode_unit(p, sc, it, cluster_id, nclusters, rank, cur);
      const int phase = tc.g % p.phases, co_off = (tc.g / p.phases) * N;
      const int as = t_it & 1;
      const float* resb = p.res ? p.res + (size_t)tc.b * (p.o_nct ? p.r_bs : p.o32_bs) : nullptr;
      float* o32b = p.o32 ? p.o32 + (size_t)tc.b * p.o32_bs : nullptr;

      // residual prefetch of the first item (overlaps the tail of this tile's MMAs)
      float4 rc[8];
      // item = (accumulator m, 32-column chunk cc); this warp takes every second item of its TMEM lane quadrant, walked
      // with (m, cc) counters (no divisions on this path: the epilogue is instruction-issue bound on the C <= 64 layers)
      auto item_row = [&](int m, int& t, bool& ok) {
        const int q = tc.q0 + m * 128 + quad * 32 + lane;
        t = q * p.ot_mul + p.ot_add + phase;
        ok = !tc.dummy && q < tc.lim && t >= 0 && t < p.T_out;
      };
      auto advance = [&](int& m, int& cc) {
        cc += 2;
        while (cc >= ncc) { cc -= ncc; ++m; }
      };
      auto load_res = [&](int m, int cc, float4 (&dst)[8]) {
        int t; bool ok;
        item_row(m, t, ok);
        if (resb && ok) {
          const int n0 = co_off + cc * 32;
          if (p.o_nct) {                                 // generic strides: element (c, t) at res[b*r_bs + c*r_cs + t*r_ts]
            const float* rp = resb + (size_t)n0 * p.r_cs + (size_t)t * p.r_ts;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              dst[k] = make_float4(rp[(size_t)(4 * k) * p.r_cs], rp[(size_t)(4 * k + 1) * p.r_cs],
                                   rp[(size_t)(4 * k + 2) * p.r_cs], rp[(size_t)(4 * k + 3) * p.r_cs]);
          } else {
            const float4* rp = reinterpret_cast<const float4*>(resb) + ((size_t)(n0 / 4) * p.T_out + t);
#pragma unroll
            for (int k = 0; k < 8; ++k) dst[k] = rp[(size_t)k * p.T_out];   // may alias o32 (in-place y += conv)
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) dst[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      int m = 0, cc = half;
      while (cc >= ncc) { cc -= ncc; ++m; }
      if (half < nitems && !p.gate && !p.ln_gamma) load_res(m, cc, rc);

      mbar_wait(acc_full(as), (t_it >> 1) & 1);
      tc_fence_after();
      if (p.il_u) {
        // ---- transposed convolution with all `u` polyphase components stacked along N (column = phase * cb + c):
        // the thread of row q holds out[t = q*u - pad + phase] for every phase, so two adjacent output samples of
        // one channel group leave as ONE 32-byte store = a whole L2 sector (16-byte strided pieces made the L2 fetch
        // every sector from HBM before merging: 2x the DRAM traffic and 5x the time of these layers).
        const int u = p.il_u, cb = p.il_cb, nc8 = cb / 8;
        const int blk = tc.g;                          // output-channel block (phases == 1 in this mode)
        for (int idx = half; idx < p.NACC * nc8; idx += 2) {
          const int m = idx / nc8, c8 = idx - m * nc8;
          const int q = tc.q0 + m * 128 + quad * 32 + lane;
          const bool okq = !tc.dummy && q < tc.lim;
          const int co = blk * cb + c8 * 8;            // first of the 8 output channels of this sub-item
          const float4 b0 = *reinterpret_cast<const float4*>(bias_s + co);
          const float4 b1 = *reinterpret_cast<const float4*>(bias_s + co + 4);
          const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256 + m * NM + c8 * 8);
          for (int ph = 0; ph < u; ph += 2) {
            uint32_t ra[8], rb[8];
            __syncwarp();
            tmem_ld8_nowait(tbase + (uint32_t)(ph * cb), ra);
            tmem_ld8_nowait(tbase + (uint32_t)((ph + 1) * cb), rb);
            if (p.stack) {
              uint32_t la[8], lb[8];
              tmem_ld8_nowait(tbase + (uint32_t)(N + ph * cb), la);
              tmem_ld8_nowait(tbase + (uint32_t)(N + (ph + 1) * cb), lb);
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                ra[k] = __float_as_uint(__uint_as_float(ra[k]) + __uint_as_float(la[k]));
                rb[k] = __float_as_uint(__uint_as_float(rb[k]) + __uint_as_float(lb[k]));
              }
            }
            tmem_ld_wait();
            const int t0 = q * u + p.ot_add + ph, t1 = t0 + 1;
            const bool ok0 = okq && t0 >= 0 && t0 < p.T_out, ok1 = okq && t1 >= 0 && t1 < p.T_out;
            if (!ok0 && !ok1) continue;
            float va[8], vb[8];
            va[0] = __uint_as_float(ra[0]) + b0.x; va[1] = __uint_as_float(ra[1]) + b0.y;
            va[2] = __uint_as_float(ra[2]) + b0.z; va[3] = __uint_as_float(ra[3]) + b0.w;
            va[4] = __uint_as_float(ra[4]) + b1.x; va[5] = __uint_as_float(ra[5]) + b1.y;
            va[6] = __uint_as_float(ra[6]) + b1.z; va[7] = __uint_as_float(ra[7]) + b1.w;
            vb[0] = __uint_as_float(rb[0]) + b0.x; vb[1] = __uint_as_float(rb[1]) + b0.y;
            vb[2] = __uint_as_float(rb[2]) + b0.z; vb[3] = __uint_as_float(rb[3]) + b0.w;
            vb[4] = __uint_as_float(rb[4]) + b1.x; vb[5] = __uint_as_float(rb[5]) + b1.y;
            vb[6] = __uint_as_float(rb[6]) + b1.z; vb[7] = __uint_as_float(rb[7]) + b1.w;
            if (o32b) {                                // fp32 stream [C/4][T][4]: samples t0, t1 of a quad are adjacent
#pragma unroll
              for (int hq = 0; hq < 2; ++hq) {
                float4* op = reinterpret_cast<float4*>(o32b) + ((size_t)(co / 4 + hq) * p.T_out + t0);
                st_pair(op,
                        make_uint4(__float_as_uint(va[4 * hq]), __float_as_uint(va[4 * hq + 1]),
                                   __float_as_uint(va[4 * hq + 2]), __float_as_uint(va[4 * hq + 3])),
                        make_uint4(__float_as_uint(vb[4 * hq]), __float_as_uint(vb[4 * hq + 1]),
                                   __float_as_uint(vb[4 * hq + 2]), __float_as_uint(vb[4 * hq + 3])),
                        ok0, ok1);
              }
            }
            if (p.o_hi) {                              // operand planes [C/8][rows][8]: rows t0, t1 of a slab are adjacent
              const size_t prow = (size_t)tc.b * p.op_bs + ((size_t)(co / 8) * p.op_rows + p.op_pad + t0) * 8;
              uint32_t ha[4], hb[4], la[4], lb[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a0 = fmaxf(va[2 * e], va[2 * e] * p.slope), a1 = fmaxf(va[2 * e + 1], va[2 * e + 1] * p.slope);
                const float c0 = fmaxf(vb[2 * e], vb[2 * e] * p.slope), c1 = fmaxf(vb[2 * e + 1], vb[2 * e + 1] * p.slope);
                if (split) {
                  split2(a0, a1, fmt, ha[e], la[e]);
                  split2(c0, c1, fmt, hb[e], lb[e]);
                } else {
                  ha[e] = pack2(a0, a1, fmt);
                  hb[e] = pack2(c0, c1, fmt);
                }
              }
              st_pair(p.o_hi + prow, make_uint4(ha[0], ha[1], ha[2], ha[3]), make_uint4(hb[0], hb[1], hb[2], hb[3]), ok0, ok1);
              if (split)
                st_pair(p.o_lo + prow, make_uint4(la[0], la[1], la[2], la[3]), make_uint4(lb[0], lb[1], lb[2], lb[3]), ok0,
                        ok1);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (pair && rank != 0) mbar_arrive_remote(acc_empty(as), 0);
          else mbar_arrive(acc_empty(as));
        }
        continue;
      }
      if constexpr (!PAIR) {                        // (not in the CTA-pair build: the vocoder's epilogue keeps its registers)
      if (p.ln_gamma) {
        // ---- residual add + channel LayerNorm in the epilogue (rel_transformer_encoder.py:58-79): the row of a tile is
        // split between the two warps of a TMEM lane quadrant (32-column chunks half, half + 2, ...); they exchange their
        // partial sums through shared memory and a 64-thread named barrier
        float* ln_s = bias_s + 1024;                    // [128 rows][2 halves][2]
        int t; bool ok;
        item_row(0, t, ok);
        const float cm = (p.mask && ok) ? __ldg(p.mask + (size_t)tc.b * p.m_bs + t) : 1.f;
        const float im = (p.ln_in_mask && ok) ? __ldg(p.ln_in_mask + (size_t)tc.b * p.m_bs + t) : 1.f;
        float s1 = 0.f, s2 = 0.f;
        for (int c2 = half; c2 < ncc; c2 += 2) {
          uint32_t r[32];
          __syncwarp();
          tmem_ld32_nowait(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256 + c2 * 32), r);
          float rr[32];
          const int n0 = co_off + c2 * 32;
          if (resb && ok) {
            const float* rp = resb + (size_t)n0 * p.r_cs + (size_t)t * p.r_ts;
#pragma unroll
            for (int k = 0; k < 32; ++k) rr[k] = rp[(size_t)k * p.r_cs];
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) rr[k] = 0.f;
          }
          tmem_ld_wait();
          if (ok) {
            float* op = o32b + (size_t)n0 * p.o_cs + (size_t)t * p.o_ts;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const float v = ((__uint_as_float(r[k]) + bias_s[n0 + k]) * cm + rr[k]) * im;
              op[(size_t)k * p.o_cs] = v;
              s1 += v;
              s2 = fmaf(v, v, s2);
            }
          }
        }
        tc_fence_before();                              // the accumulators are not read again
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(as));
        const int row = quad * 32 + lane;
        ln_s[row * 4 + half * 2] = s1;
        ln_s[row * 4 + half * 2 + 1] = s2;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");
        const float S1 = ln_s[row * 4] + ln_s[row * 4 + 2], S2 = ln_s[row * 4 + 1] + ln_s[row * 4 + 3];
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");     // both warps have read before the next tile writes
        const float mean = S1 / (float)N;
        const float rstd = rsqrtf(fmaxf(S2 / (float)N - mean * mean, 0.f) + p.ln_eps);
        if (ok) {
          const float om = p.ln_out_mask ? __ldg(p.ln_out_mask + (size_t)tc.b * p.m_bs + t) : 1.f;
          for (int c2 = half; c2 < ncc; c2 += 2) {
            const int n0 = co_off + c2 * 32;
            const float* op = o32b + (size_t)n0 * p.o_cs + (size_t)t * p.o_ts;      // this thread's own stores
            float v[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = op[(size_t)k * p.o_cs];
#pragma unroll
            for (int k = 0; k < 32; ++k)
              v[k] = ((v[k] - mean) * rstd * __ldg(p.ln_gamma + n0 + k) + __ldg(p.ln_beta + n0 + k)) * om;
            if (p.ln_y) {
              float* yp = p.ln_y + (size_t)tc.b * p.o32_bs + (size_t)n0 * p.o_cs + (size_t)t * p.o_ts;
#pragma unroll
              for (int k = 0; k < 32; ++k) yp[(size_t)k * p.o_cs] = v[k];
            }
            const size_t prow = (size_t)tc.b * p.op_bs + ((size_t)(n0 / 8) * p.op_rows + p.op_pad + t) * 8;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (split) split2(v[8 * h + 2 * e], v[8 * h + 2 * e + 1], fmt, hw[e], lw[e]);
                else hw[e] = pack2(v[8 * h + 2 * e], v[8 * h + 2 * e + 1], fmt);
              }
              const size_t off = prow + (size_t)h * p.op_rows * 8;
              *reinterpret_cast<uint4*>(p.o_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              if (split) *reinterpret_cast<uint4*>(p.o_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
        }
        continue;
      }
      if (p.gate) {
        // ---- WaveNet gate (wavenet.py:64-70) on the accumulators: acts never exist as a [B, 2H, T] fp32 tensor
        const int nh = N / 64;                         // 32-channel gate items per sub-tile
        const int cbase = (tc.g / p.phases) * (N / 2);
        for (int idx = half; idx < p.NACC * nh; idx += 2) {
          const int gm = idx / nh, gc = idx - gm * nh;
          int t; bool ok;
          item_row(gm, t, ok);
          const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256 + gm * NM + gc * 32);
          const int ct = co_off + gc * 32, cs = ct + N / 2;
          const size_t prow = (size_t)tc.b * p.op_bs + ((size_t)((cbase + gc * 32) / 8) * p.op_rows + p.op_pad + (ok ? t : 0)) * 8;
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {                // 16 channels at a time (register pressure)
            uint32_t ra[16], rb[16];
            __syncwarp();                                 // tcgen05.ld is .sync.aligned
            tmem_ld16_nowait(tcol + (uint32_t)(16 * hh), ra);
            tmem_ld16_nowait(tcol + (uint32_t)(N / 2 + 16 * hh), rb);
            float rt[16], rs[16];
            if (resb && ok) {
              const float* pt = resb + (size_t)(ct + 16 * hh) * p.r_cs + (size_t)t * p.r_ts;
              const float* ps = resb + (size_t)(cs + 16 * hh) * p.r_cs + (size_t)t * p.r_ts;
#pragma unroll
              for (int k = 0; k < 16; ++k) { rt[k] = pt[(size_t)k * p.r_cs]; rs[k] = ps[(size_t)k * p.r_cs]; }
            } else {
#pragma unroll
              for (int k = 0; k < 16; ++k) rt[k] = rs[k] = 0.f;
            }
            tmem_ld_wait();
            if (!ok) continue;
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k)
              v[k] = gate_fast(__uint_as_float(ra[k]) + bias_s[ct + 16 * hh + k] + rt[k],
                               __uint_as_float(rb[k]) + bias_s[cs + 16 * hh + k] + rs[k]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (split) split2(v[8 * h + 2 * e], v[8 * h + 2 * e + 1], fmt, hw[e], lw[e]);
                else hw[e] = pack2(v[8 * h + 2 * e], v[8 * h + 2 * e + 1], fmt);
              }
              const size_t off = prow + (size_t)(2 * hh + h) * p.op_rows * 8;
              *reinterpret_cast<uint4*>(p.o_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              if (split) *reinterpret_cast<uint4*>(p.o_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (pair && rank != 0) mbar_arrive_remote(acc_empty(as), 0);
          else mbar_arrive(acc_empty(as));
        }
        continue;
      }
      }
      // The item loop exists twice: with a residual (conv2 of a ResBlock: the next item's residual is prefetched while
      // this one is processed -- that ordering is what keeps those HBM-bound layers at 6 TB/s) and without one (conv1:
      // no residual registers at all; its epilogue is latency bound, 64 fewer moves per item).
      auto items = [&](auto has_res_tag) {
      constexpr bool HAS_RES = decltype(has_res_tag)::value;
      for (int idx = half; idx < nitems; idx += 2) {
        int t; bool ok;
        item_row(m, t, ok);
        uint32_t r[32];
        __syncwarp();                                  // tcgen05.ld is .sync.aligned
        const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256 + m * NM + cc * 32);
        tmem_ld32_nowait(tcol, r);
        float4 rn[8];
        int mn = m, ccn = cc;
        advance(mn, ccn);
        if constexpr (HAS_RES) {
          if (idx + 2 < nitems) load_res(mn, ccn, rn);   // next item's residual is in flight while this one is processed
        }
        if (p.stack) {                                 // a*w_lo landed N columns further: fold it in
          uint32_t r2[32];
          tmem_ld32_nowait(tcol + (uint32_t)N, r2);
          tmem_ld_wait();
          float* rf = reinterpret_cast<float*>(r);
          const float* r2f = reinterpret_cast<const float*>(r2);
#pragma unroll
          for (int k = 0; k < 32; k += 2) add2(rf[k], rf[k + 1], r2f[k], r2f[k + 1]);
        }
        tmem_ld_wait();
        if (ok) {
          const int n0 = co_off + cc * 32;
          float v[32];
          const float4* bs4 = reinterpret_cast<const float4*>(bias_s + n0);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 bq = bs4[k];
            v[4 * k] = __uint_as_float(r[4 * k]); v[4 * k + 1] = __uint_as_float(r[4 * k + 1]);
            v[4 * k + 2] = __uint_as_float(r[4 * k + 2]); v[4 * k + 3] = __uint_as_float(r[4 * k + 3]);
            fma2(v[4 * k], v[4 * k + 1], acc_scale, bq.x, bq.y);      // lo8 layers accumulate 2^10 * conv (weights pre-scaled)
            fma2(v[4 * k + 2], v[4 * k + 3], acc_scale, bq.z, bq.w);
          }
          if (p.act == 1) {
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = fmaxf(v[k], 0.f);
          } else if (p.act == 3) {                      // exact GELU (FFT-block FFN, common_layers.py:640-647)
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = 0.5f * v[k] * (1.f + erff(v[k] * 0.70710678118654752f));
          }
          if (p.alpha != 1.f || p.mask) {               // out = post * (act(conv + bias) * alpha * mask + res)
            const float am = p.alpha * (p.mask ? __ldg(p.mask + (size_t)tc.b * p.m_bs + t) : 1.f);
#pragma unroll
            for (int k = 0; k < 32; k += 2) mul2(v[k], v[k + 1], am, am);
          }
          if constexpr (HAS_RES) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              add2(v[4 * k], v[4 * k + 1], rc[k].x, rc[k].y);
              add2(v[4 * k + 2], v[4 * k + 3], rc[k].z, rc[k].w);
            }
          }
          if (p.post != 1.f) {
#pragma unroll
            for (int k = 0; k < 32; k += 2) mul2(v[k], v[k + 1], p.post, p.post);
          }
          if (o32b && p.o_nct) {                        // element (c, t) at out[b*o32_bs + c*o_cs + t*o_ts]
            float* op = o32b + (size_t)n0 * p.o_cs + (size_t)t * p.o_ts;
            const int kmax = p.c_valid - n0;            // channels >= c_valid are zero padding of the weights (C_out % 32)
            if (p.accumulate) {
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (k < kmax) v[k] += op[(size_t)k * p.o_cs];
            }
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (k < kmax) op[(size_t)k * p.o_cs] = v[k];
          } else if (o32b) {
            float4* op = reinterpret_cast<float4*>(o32b) + ((size_t)(n0 / 4) * p.T_out + t);
            if (p.accumulate) {
              float4 old[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) old[k] = op[(size_t)k * p.T_out];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                v[4 * k] += old[k].x; v[4 * k + 1] += old[k].y; v[4 * k + 2] += old[k].z; v[4 * k + 3] += old[k].w;
              }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
              op[(size_t)k * p.T_out] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          }
          if (p.o_hi) {
            const size_t prow = (size_t)tc.b * p.op_bs + ((size_t)(n0 / 8) * p.op_rows + p.op_pad + t) * 8;
            // leaky(v) = max(v, slope * v) for 0 <= slope <= 1 (all HiFi-GAN slopes)
            if (split) {
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float l0, l1;
                  leaky2(v[8 * h + 2 * e], v[8 * h + 2 * e + 1], p.slope, l0, l1);
                  split2(l0, l1, fmt, hw[e], lw[e]);
                }
                const size_t off = prow + (size_t)h * p.op_rows * 8;
                *reinterpret_cast<uint4*>(p.o_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(p.o_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              }
            } else {
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                uint32_t hw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float l0, l1;
                  leaky2(v[8 * h + 2 * e], v[8 * h + 2 * e + 1], p.slope, l0, l1);
                  hw[e] = pack2(l0, l1, fmt);
                }
                *reinterpret_cast<uint4*>(p.o_hi + prow + (size_t)h * p.op_rows * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              }
            }
          }
        }
        if constexpr (HAS_RES) {
#pragma unroll
          for (int k = 0; k < 8; ++k) rc[k] = rn[k];
        }
        m = mn; cc = ccn;
      }
      };
      if (resb) items(std::true_type{});
      else items(std::false_type{});
      // this warp no longer reads accumulator set `as`
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (pair && rank != 0) mbar_arrive_remote(acc_empty(as), 0);   // the leader issues the MMAs of both CTAs
        else mbar_arrive(acc_empty(as));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();           // no peer may still multicast into / arrive on this CTA's shared memory
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// helper kernels
// ---------------------------------------------------------------------------------------------------------------
__global__ void tc_pack_weights_kernel(const float* __restrict__ w, tc16* __restrict__ out, int C_out,
                                       int C_in, int K, int transposed, int stride, int N, int KC, int planes,
                                       int ktaps, int phases, int fmt, int stack, int il_cb, int pair, float wscale) {
  // il_cb > 0 (transposed only): all `stride` polyphase components of a block of il_cb output channels are stacked along
  // N (row n = phase * il_cb + c); the caller passes phases = 1 and N = stride * il_cb.
  const size_t total = (size_t)C_out * C_in * ktaps * (il_cb ? stride : phases) * planes * (stack ? 2 : 1);
  const int NMs = stack ? 2 * N : N;                 // rows per K slab of one blob plane
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // destination index [g][chunk][tap][plane][slab][n][e]
  size_t r = i;
  const int e = r % 8; r /= 8;
  int n, sl, pl;
  if (pair) {                                        // [..][tap][half][plane][slab][N/2][8]: each CTA of a pair copies its half
    n = r % (NMs / 2); r /= (NMs / 2);
    sl = r % (KC / 8); r /= (KC / 8);
    pl = r % planes; r /= planes;
    n += (int)(r % 2) * (NMs / 2); r /= 2;
  } else {
    n = r % NMs; r /= NMs;
    sl = r % (KC / 8); r /= (KC / 8);
    pl = r % planes; r /= planes;
  }
  if (stack) { pl = n >= N; n -= pl * N; }           // stacked: rows [0,N) = hi, [N,2N) = lo of the same channel
  const int j = r % ktaps; r /= ktaps;
  const int nchunks = C_in / KC;
  const int c = r % nchunks; r /= nchunks;
  const int g = (int)r;
  int phase = g % phases, nb = g / phases;
  int co = nb * N + n;
  const int ci = c * KC + sl * 8 + e;
  if (il_cb) { phase = n / il_cb; co = nb * il_cb + n % il_cb; }
  float v;
  if (transposed) v = w[((size_t)ci * C_out + co) * K + phase + j * stride];
  else v = w[((size_t)co * C_in + ci) * K + j];
  v *= wscale;                                       // lo8 layers: 2^10 (exact), undone by the epilogue's acc_scale
  const uint32_t hi = cvt16(v, fmt);
  out[i] = (tc16)(pl ? cvt16(v - back16(hi, fmt), fmt) : hi);
}

// e5m2 lo plane of the weights (TcConvW::lo8): out[g][chunk][tap]([half])[KC/16][Nh][16] = e5m2((w - fp16(w)) * 2^10)
__global__ void tc_pack_weights_lo8_kernel(const float* __restrict__ w, uint8_t* __restrict__ out, int C_out, int C_in,
                                           int K, int N, int KC, int fmt, int pair) {
  const size_t total = (size_t)C_out * C_in * K;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  size_t r = i;
  const int e = r % 16; r /= 16;
  const int Nh = pair ? N / 2 : N;
  int n = r % Nh; r /= Nh;
  const int sl = r % (KC / 16); r /= (KC / 16);
  if (pair) { n += (int)(r % 2) * Nh; r /= 2; }
  const int j = r % K; r /= K;
  const int nchunks = C_in / KC;
  const int c = r % nchunks; r /= nchunks;
  const int nb = (int)r;
  const int co = nb * N + n, ci = c * KC + sl * 16 + e;
  const float v = w[((size_t)co * C_in + ci) * K + j];
  const float lo = (v - back16(cvt16(v, fmt), fmt)) * kLo8WScale;
  uint16_t q;
  asm("cvt.rn.satfinite.e5m2x2.f32 %0, %1, %2;" : "=h"(q) : "f"(0.f), "f"(lo));
  out[i] = (uint8_t)(q & 0xFF);
}

__global__ void tc_to_planes_kernel(const float* __restrict__ x, long bs, long cs, long ts, int C, int T, float slope,
                                    tc16* __restrict__ hi, tc16* __restrict__ lo, int rows, int pad, int fmt, int full,
                                    int slabs_total, int slab0, int s2d_H) {
  // s2d_H > 0: space-to-depth source -- destination channel ch reads x[(ch % s2d_H) * cs + t * ts + ch / s2d_H]
  // (the stride-4 g_pre_net as a stride-1 convolution over 4 * H channels, acoustic.cu)
  // full: the grid walks ALL rows of the slab and zero-fills the halo rows (one launch instead of convert + zero_halo)
  // slabs_total / slab0: the C source channels fill slabs [slab0, slab0 + C/8) of planes holding slabs_total slabs
  griddep_launch_if_resident();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = full ? r - pad : r;
  const int sl = blockIdx.y, b = blockIdx.z;
  const size_t dslab = (size_t)b * slabs_total + slab0 + sl;
  if (full) {
    if (r >= rows) return;
    if (t < 0 || t >= T) {
      const size_t zoff = (dslab * rows + r) * 8;
      *reinterpret_cast<uint4*>(hi + zoff) = make_uint4(0, 0, 0, 0);
      if (lo) *reinterpret_cast<uint4*>(lo + zoff) = make_uint4(0, 0, 0, 0);
      return;
    }
  } else if (t >= T) {
    return;
  }
  uint32_t hw[4], lw[4];
  if (cs == 1 && ((bs | ts) & 3) == 0 && (((uintptr_t)x) & 15) == 0) {
    // channels-last source (e.g. the [.., dict_dim] gloss rows of S2PA): the 8 channels of a slab are 32 contiguous bytes
    const float4* src = reinterpret_cast<const float4*>(x + (size_t)b * bs + (size_t)t * ts + (size_t)sl * 8);
    const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
    split2(leaky(v0.x, slope), leaky(v0.y, slope), fmt, hw[0], lw[0]);
    split2(leaky(v0.z, slope), leaky(v0.w, slope), fmt, hw[1], lw[1]);
    split2(leaky(v1.x, slope), leaky(v1.y, slope), fmt, hw[2], lw[2]);
    split2(leaky(v1.z, slope), leaky(v1.w, slope), fmt, hw[3], lw[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c0 = sl * 8 + 2 * e, c1 = c0 + 1;
      const size_t o0 = s2d_H ? (size_t)(c0 % s2d_H) * cs + c0 / s2d_H : (size_t)c0 * cs;
      const size_t o1 = s2d_H ? (size_t)(c1 % s2d_H) * cs + c1 / s2d_H : (size_t)c1 * cs;
      const float a0 = leaky(x[(size_t)b * bs + o0 + (size_t)t * ts], slope);
      const float a1 = leaky(x[(size_t)b * bs + o1 + (size_t)t * ts], slope);
      split2(a0, a1, fmt, hw[e], lw[e]);
    }
  }
  const size_t off = (dslab * rows + pad + t) * 8;
  *reinterpret_cast<uint4*>(hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  if (lo) *reinterpret_cast<uint4*>(lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

__global__ void tc_zero_halo_kernel(tc16* __restrict__ hi, tc16* __restrict__ lo, int rows, int pad,
                                    int T) {
  const size_t slab = blockIdx.x;
  const int nz = rows - T;                       // zero rows: [0,pad) then [pad+T, rows)
  uint4* ph = reinterpret_cast<uint4*>(hi) + slab * rows;
  uint4* pl = lo ? reinterpret_cast<uint4*>(lo) + slab * rows : nullptr;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < nz; i += blockDim.x) {
    const int row = i < pad ? i : T + i;
    ph[row] = z;
    if (pl) pl[row] = z;
  }
}

__global__ void tc_stream_to_nct_kernel(const float* __restrict__ st, float* __restrict__ out, int C, int T) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int s4 = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const float4 v = reinterpret_cast<const float4*>(st)[((size_t)b * (C / 4) + s4) * T + t];
  float* o = out + ((size_t)b * C + s4 * 4) * T + t;
  o[0] = v.x; o[(size_t)T] = v.y; o[(size_t)2 * T] = v.z; o[(size_t)3 * T] = v.w;
}
__global__ void tc_nct_to_stream_kernel(const float* __restrict__ in, float* __restrict__ st, int C, int T) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int s4 = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const float* o = in + ((size_t)b * C + s4 * 4) * T + t;
  reinterpret_cast<float4*>(st)[((size_t)b * (C / 4) + s4) * T + t] =
      make_float4(o[0], o[(size_t)T], o[(size_t)2 * T], o[(size_t)3 * T]);
}

__global__ void tc_planes_to_nct_kernel(const tc16* __restrict__ hi, const tc16* __restrict__ lo,
                                        float* __restrict__ out, int C, int T, int rows, int pad, int fmt) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int sl = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const size_t off = (((size_t)b * (C / 8) + sl) * rows + pad + t) * 8;
  for (int e = 0; e < 8; ++e) {
    float v = back16(hi[off + e], fmt);
    if (lo) v += back16(lo[off + e], fmt);
    out[((size_t)b * C + sl * 8 + e) * T + t] = v;
  }
}

// conv_post + tanh on the CUDA cores (C_out = 1: 2*C*K flop per sample against C*4 bytes read).  Block = 256 samples of
// one item: the (256 + K - 1) x C input tile is staged once in shared memory with the leaky-ReLU applied (the first
// version loaded and activated every element K times from L1 and was instruction bound: 0.40 ms at cfg 2), then every
// thread runs its C*K FMAs from shared memory in the same order as before (bit-identical results).
template <int KMAX>
__global__ void __launch_bounds__(256) tc_conv_post_kernel(const float* __restrict__ st, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ wav,
                                                           int C, int T, int K, float slope,
                                                           const int* __restrict__ lens, int len_mul) {
  __shared__ float ws[64 * KMAX];
  extern __shared__ float4 tile[];                   // [C/4][256 + K - 1]
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * 256;
  // ragged: samples past the item's valid length are zero (their inputs were never computed)
  const long valid = lens ? (long)__ldg(lens + b) * len_mul : (long)T;
  const int t = t0 + threadIdx.x;
  if (lens && (long)t0 >= valid) {                   // whole block past the end
    if (t < T) wav[(size_t)b * T + t] = 0.f;
    return;
  }
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) ws[i] = w[i];
  const int pad = (K - 1) / 2, W = 256 + K - 1;
  const float4* sp = reinterpret_cast<const float4*>(st) + (size_t)b * (C / 4) * T;
  for (int i = threadIdx.x; i < (C / 4) * W; i += blockDim.x) {
    const int s4 = i / W, x = i - s4 * W;
    const int tt = t0 + x - pad;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tt >= 0 && tt < T) v = __ldg(sp + (size_t)s4 * T + tt);
    tile[i] = make_float4(leaky(v.x, slope), leaky(v.y, slope), leaky(v.z, slope), leaky(v.w, slope));
  }
  __syncthreads();
  if (t >= T) return;
  if ((long)t >= valid) { wav[(size_t)b * T + t] = 0.f; return; }
  float acc = bias ? bias[0] : 0.f;
  for (int s4 = 0; s4 < C / 4; ++s4) {
    const float4* row = tile + s4 * W + threadIdx.x;
    const float* wc = ws + s4 * 4 * K;
    for (int j = 0; j < K; ++j) {
      const int tt = t + j - pad;
      if (tt < 0 || tt >= T) continue;               // (zero rows: skipped like the original loop, same summation)
      const float4 x = row[j];
      acc = fmaf(wc[j], x.x, acc);
      acc = fmaf(wc[K + j], x.y, acc);
      acc = fmaf(wc[2 * K + j], x.z, acc);
      acc = fmaf(wc[3 * K + j], x.w, acc);
    }
  }
  wav[(size_t)b * T + t] = tanhf(acc);
}

// wav[b,t] = tanh(bias + sum_j part[b][j][t + j - 3]): the shifted sum of the per-tap partials the last fused pair wrote
__global__ void __launch_bounds__(256) tc_conv_post_finish_kernel(const float* __restrict__ part,
                                                                  const float* __restrict__ bias,
                                                                  float* __restrict__ wav, int T,
                                                                  const int* __restrict__ lens, int len_mul) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const long valid = lens ? (long)__ldg(lens + b) * len_mul : (long)T;
  if ((long)t >= valid) { wav[(size_t)b * T + t] = 0.f; return; }
  const float* pb = part + (size_t)b * 7 * T;
  float acc = bias ? __ldg(bias) : 0.f;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const int tt = t + j - 3;
    if (tt >= 0 && tt < T) acc += __ldg(pb + (size_t)j * T + tt);
  }
  wav[(size_t)b * T + t] = tanhf(acc);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
cudaError_t tc_conv_init() {
  cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(tc_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
// The DTTS_TC_* tuning knobs, read ONCE per process (thread-safe function-local static): tc_conv_plan runs for every
// convolution launch (~100 per vocode) and getenv is neither free nor safe against a concurrent setenv.
struct TcEnv { int nacc, astages, tg, wstages, cluster; };
static const TcEnv& tc_env() {
  static const TcEnv e = {env_int("DTTS_TC_NACC", 0), env_int("DTTS_TC_ASTAGES", 0), env_int("DTTS_TC_TG", 0),
                          env_int("DTTS_TC_WSTAGES", 0), env_int("DTTS_TC_CLUSTER", 0)};
  return e;
}

void tc_conv_plan(TcConvParams* p, const TcConvW& w, int nq, int a_planes) {
  p->w = w.w; p->bias = w.bias;
  p->C_in = w.C_in; p->N = w.N; p->KC = w.KC; p->nchunks = w.C_in / w.KC; p->ktaps = w.ktaps;
  p->a_planes = a_planes; p->w_planes = w.planes; p->fmt = w.fmt; p->stack = w.stack; p->NM = w.stack ? 2 * w.N : w.N;
  p->nblocks = (w.il_u ? w.C_out * w.il_u : w.C_out) / w.N; p->phases = w.phases;
  p->il_u = w.il_u; p->il_cb = w.il_cb;
  p->pair = w.pair;
  p->lo8 = w.lo8; p->w8 = w.w8;
  p->acc_scale = w.lo8 ? 1.f / kLo8WScale : 1.f;
  p->nq = nq;
  int nacc = 256 / p->NM;                                    // one accumulator set = 256 TMEM columns (two sets)
  if (nacc > 4) nacc = 4;
  const int force = tc_env().nacc;
  if (force > 0 && force <= nacc) nacc = force;
  while (nacc > 1 && (nacc - 1) * 128 >= nq) --nacc;         // short sequences: do not compute empty sub-tiles
  p->NACC = nacc;
  p->MT = 128 * nacc;
  p->ntiles = cdiv(nq, p->MT);
  const int o0 = p->tap_off0, o1 = p->tap_off0 + (p->ktaps - 1) * p->tap_step;
  p->min_off = o0 < o1 ? o0 : o1;
  const int max_off = o0 < o1 ? o1 : o0;
  p->RA = p->MT + (max_off - p->min_off);
  const size_t a_stage = (size_t)(p->KC / 8) * p->RA * 16 * p->a_planes + (p->lo8 ? (size_t)(p->KC / 16) * p->RA * 16 : 0);
  // activation stages: the (tile + halo) loads come from HBM with ~1.5 us latency; short tiles (few taps, small C) need
  // more of them in flight than the 2 a long K loop gets away with.  Keep at least ~64 KB for the weight stages.
  int as = tc_env().astages;
  if (as <= 0) as = p->lo8 ? 3 : 2;                      // measured: 3 or 4 stages do not help (profiles/r01_summary.md);
                                                         // lo8 tiles are converted in shared memory after they land: one more
  if (as > kMaxAStages) as = kMaxAStages;
  while (as > 2 && (size_t)as * a_stage > (size_t)kSmemLimit - kSmemHeader - 64 * 1024) --as;
  p->a_stages = as;
  const size_t w_tap = (size_t)(p->pair ? p->NM / 2 : p->NM) * p->KC * (2 * p->w_planes + (p->lo8 ? 1 : 0));   // per CTA
  // taps per weight stage: ~32 KB stages, so that the per-stage barrier round trip is amortised over >= 8 MMAs
  int tg = tc_env().tg;
  if (tg <= 0) tg = (int)((32 * 1024) / w_tap);
  if (tg < 1) tg = 1;
  if (tg > p->ktaps) tg = p->ktaps;
  tg = cdiv(p->ktaps, cdiv(p->ktaps, tg));                   // balance the groups (11 taps, 4 per stage -> 4 + 4 + 3)
  p->TG = tg;
  const size_t w_blob = w_tap * tg;
  const size_t budget = kSmemLimit - kSmemHeader - p->a_stages * a_stage;
  int ws = (int)(budget / w_blob);
  if (ws > kMaxWStages) ws = kMaxWStages;
  const int force_ws = tc_env().wstages;
  if (force_ws > 0 && force_ws < ws) ws = force_ws;
  p->w_stages = ws;
  p->csize = 1;
  p->nu = 0;
  p->o_nct = 0; p->o_cs = p->o_ts = 0; p->r_bs = p->r_cs = p->r_ts = 0;
  p->mask = nullptr; p->m_bs = 0; p->act = 0; p->alpha = 1.f;
  p->c_valid = w.C_out;
  p->lens = nullptr; p->len_mul = 0; p->len_add = 0;
}

static size_t tc_smem_bytes(const TcConvParams& p) {
  const size_t a_stage = (size_t)(p.KC / 8) * p.RA * 16 * p.a_planes + (p.lo8 ? (size_t)(p.KC / 16) * p.RA * 16 : 0);
  const size_t w_blob = (size_t)(p.pair ? p.NM / 2 : p.NM) * p.KC * (2 * p.w_planes + (p.lo8 ? 1 : 0)) * p.TG;
  size_t bytes = kSmemHeader + p.a_stages * a_stage + p.w_stages * w_blob;
  // the CTA owns all 512 TMEM columns: it must be alone on its SM, or a co-resident CTA would block in tcgen05.alloc
  if (bytes < 116 * 1024) bytes = 116 * 1024;
  return bytes;
}

int tc_pair_enabled() {
  static const int v = env_int("DTTS_TC_PAIR", 1) != 0;
  return v;
}

int tc_pdl_enabled() {
  static const int v = env_int("DTTS_TC_PDL", 1) != 0;
  return v;
}

int tc_nmax() {
  static const int v = [] {
    const int n = env_int("DTTS_TC_NMAX", 256);
    return (n == 64 || n == 128 || n == 256) ? n : 256;
  }();
  return v;
}

int tc_pair64_cluster_enabled() {
  static const int v = env_int("DTTS_TC_PAIR64_CLUSTER", 0) != 0;
  return v;
}

int tc_fuse64_enabled() {
  static const int v = env_int("DTTS_TC_FUSE64", 1) != 0;
  return v;
}

static int g_fuse_override = -1;            // dtts_debug_set_tc_fuse (unit tests): -1 = follow DTTS_TC_FUSE
void tc_fuse_override(int v) { g_fuse_override = v; }
int tc_fuse_enabled() {
  static const int v = env_int("DTTS_TC_FUSE", 1) != 0;
  return g_fuse_override >= 0 ? (g_fuse_override != 0) : v;
}
int tc_fuse128_maxk() {
  static const int v = env_int("DTTS_TC_FUSE128_MAXK", 3);
  return g_fuse_override == 3 ? 11 : v;
}
int tc_fuse_block_enabled() {
  static const int v = env_int("DTTS_TC_FUSE_BLOCK", 1) != 0;
  if (g_fuse_override >= 0) return g_fuse_override == 4;
  return v && tc_fuse_enabled();
}
int tc_fold_post_enabled() {
  static const int v = env_int("DTTS_TC_FOLD_POST", 1) != 0;
  if (g_fuse_override >= 0) return g_fuse_override == 1;
  return v && tc_fuse_enabled();
}

int tc_lo8_min_taps() {
  static const int v = env_int("DTTS_TC_LO8_MINTAPS", 7);
  return v;
}

// Occupancy facts of the device a launch goes to, cached PER DEVICE ORDINAL (one process may drive several GPUs, e.g.
// the reference-plugin path) and filled under a mutex (several host threads, one handle each).
struct DevInfo {
  int num_sms = 0;
  int max_clusters[9] = {0};     // [csize] -> co-resident clusters of tc_conv_kernel<false> (0 = not queried yet)
  int max_pairs = 0;             // co-resident CTA pairs of tc_conv_kernel<true>
};
constexpr int kMaxDevices = 64;
static DevInfo g_dev[kMaxDevices];
static std::mutex g_dev_mu;

// Cluster size (CTAs sharing every weight stage through multicast).  Measured on B200 at the cfg-2 vocoder shapes
// (profiles/r01_summary.md): 1 -> 35.2 ms, 2 -> 36.0 ms, 4 -> 36.0 ms per pass; the weight stream (<= 3.8 TB/s out of
// L2) is not what bounds these layers, so the default is 1 and DTTS_TC_CLUSTER=2|4 turns the multicast path on.
static int pick_cluster(const TcConvParams& p, long row_tiles) {
  if (p.pair) return 2;
  if (p.lo8) return 1;
  int c = tc_env().cluster;
  if (c <= 0) c = 1;
  while (c > 1 && (row_tiles < c || (p.NM * p.KC * 2 * p.w_planes) % (16 * c))) c >>= 1;
  return c;
}

cudaError_t launch_tc_conv(TcConvParams p, int B, cudaStream_t stream) {
  if (B <= 0 || p.nq <= 0) return cudaSuccess;
  if (p.a_planes < 1 || p.a_planes > 2 || p.w_planes < 1 || p.w_planes > 2 || (p.a_planes == 2 && !p.a_lo))
    return cudaErrorInvalidValue;
  if (p.pair && (p.stack || p.a_planes != 1 || p.NM % 32 || p.NM < 64)) return cudaErrorInvalidConfiguration;
  if (p.lo8 && (!p.w8 || p.KC != 32 || p.a_planes != 1 || p.w_planes != 1 || p.stack || p.il_u || p.fmt != 0))
    return cudaErrorInvalidConfiguration;
  if (p.TG < 1 || p.w_stages < 1 || p.N % 32 != 0 || p.N > 256 || p.KC % 16 != 0 || p.C_in % p.KC != 0 ||
      p.N * p.nblocks > kMaxBias || p.NACC * p.NM > 256 || (p.stack && (p.w_planes != 1 || p.NM != 2 * p.N)) ||
      (!p.stack && p.NM != p.N))
    return cudaErrorInvalidConfiguration;
  if (p.ln_gamma && (!p.ln_beta || !p.o32 || !p.o_hi || !p.o_nct || p.nblocks != 1 || p.NACC != 1 || p.N > 256 ||
                     p.c_valid != p.N || p.il_u || p.stack || p.pair || p.phases != 1 || p.ot_mul != 1 || p.ot_add != 0 ||
                     p.act || p.accumulate || p.alpha != 1.f || p.post != 1.f || p.gate || p.lo8))
    return cudaErrorInvalidConfiguration;
  if (p.gate && (p.N % 64 || !p.o_hi || !p.o_nct || p.il_u || p.stack || p.pair || p.phases != 1 || p.ot_mul != 1))
    return cudaErrorInvalidConfiguration;
  const int max_off = p.min_off + (p.RA - p.MT);
  if (p.a_pad + p.min_off < 0 || p.a_pad + p.ntiles * p.MT + max_off > p.a_rows) return cudaErrorInvalidValue;
  const size_t smem = tc_smem_bytes(p);
  if (smem > (size_t)kSmemLimit) return cudaErrorInvalidConfiguration;
  int dev = 0;
  {
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  }
  std::lock_guard<std::mutex> lock(g_dev_mu);      // the queries below run once per (device, cluster size)
  DevInfo& di = g_dev[dev];
  int& g_num_sms = di.num_sms;
  int& g_max_pairs = di.max_pairs;
  int (&g_max_clusters)[9] = di.max_clusters;
  if (g_num_sms == 0) {
    cudaError_t e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
  }
  p.B = B;
  if (p.lens && B > TC_MAX_RAGGED_ITEMS) return cudaErrorInvalidValue;
  // (a ragged launch has at most as many row tiles as the uniform one: the grid below is an upper bound, CTAs that
  // find no unit in the device-side schedule leave after the prologue)
  const long row_tiles = (long)p.ntiles * B;
  int csize = pick_cluster(p, row_tiles);
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // see griddep_wait() in the kernel
  attr[1].val.programmaticStreamSerializationAllowed = tc_pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  int max_clusters = 0;
  if (p.pair) {
    if (g_max_pairs == 0) {
      attr[0].val.clusterDim = {2, 1, 1};
      cfg.gridDim = dim3((unsigned)(g_num_sms / 2 * 2));
      int n = 0;
      cfg.dynamicSmemBytes = kSmemLimit;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, tc_conv_kernel<true>, &cfg);
      cfg.dynamicSmemBytes = smem;
      if (e != cudaSuccess) return e;
      g_max_pairs = n > 0 ? n : -1;
    }
    if (g_max_pairs <= 0) return cudaErrorInvalidConfiguration;
    p.csize = 2;
    p.nu = (int)((row_tiles + 1) / 2);
    const long units = (long)p.nu * p.nblocks;
    const int npairs = (int)(units < g_max_pairs ? units : g_max_pairs);
    attr[0].val.clusterDim = {2, 1, 1};
    cfg.gridDim = dim3((unsigned)(npairs * 2));
    return cudaLaunchKernelEx(&cfg, tc_conv_kernel<true>, p);
  }
  for (; csize >= 1; csize >>= 1) {
    if (csize == 1) { max_clusters = g_num_sms; break; }
    if (g_max_clusters[csize] == 0) {
      attr[0].val.clusterDim = {(unsigned)csize, 1, 1};
      cfg.gridDim = dim3((unsigned)(g_num_sms / csize * csize));
      int n = 0;
      cfg.dynamicSmemBytes = kSmemLimit;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, tc_conv_kernel<false>, &cfg);
      cfg.dynamicSmemBytes = smem;
      g_max_clusters[csize] = (e == cudaSuccess && n > 0) ? n : -1;
      if (e != cudaSuccess) (void)cudaGetLastError();
    }
    if (g_max_clusters[csize] > 0) { max_clusters = g_max_clusters[csize]; break; }
  }
  p.csize = csize;
  p.nu = (int)((row_tiles + csize - 1) / csize);
  const long total_units = (long)p.nu * p.nblocks;             // each walks its `phases` polyphase components
  const int nclusters = (int)(total_units < max_clusters ? total_units : max_clusters);   // persistent
  attr[0].val.clusterDim = {(unsigned)csize, 1, 1};
  cfg.gridDim = dim3((unsigned)(nclusters * csize));
  return cudaLaunchKernelEx(&cfg, tc_conv_kernel<false>, p);
}

cudaError_t tc_pack_weights(const float* w_ref, tc16* out, int C_out, int C_in, int K, int transposed,
                            int stride, int N, int KC, int planes, int fmt, int stack, cudaStream_t s, int il_cb,
                            int pair, float wscale) {
  const int phases = (transposed && !il_cb) ? stride : 1;
  const int ktaps = transposed ? K / stride : K;
  if (il_cb) {
    if (!transposed || N != stride * il_cb || C_out % il_cb || (stride & 1) || il_cb % 8) return cudaErrorInvalidValue;
  } else if (C_out % N) {
    return cudaErrorInvalidValue;
  }
  if (C_in % KC || (transposed && K % stride)) return cudaErrorInvalidValue;
  if (stack && (planes != 1 || 2 * N > 256)) return cudaErrorInvalidValue;
  const size_t total = (size_t)C_out * C_in * ktaps * (il_cb ? stride : phases) * planes * (stack ? 2 : 1);
  tc_pack_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w_ref, out, C_out, C_in, K, transposed, stride,
                                                                        N, KC, planes, ktaps, phases, fmt, stack, il_cb, pair, wscale);
  return cudaGetLastError();
}

cudaError_t tc_to_planes(const float* x, long bs, long cs, long ts, int B, int C, int T, float slope,
                         tc16* hi, tc16* lo, int rows, int pad, int fmt, cudaStream_t s) {
  if (C % 8) return cudaErrorInvalidValue;
  dim3 grid(cdiv(T, 128), C / 8, B);
  tc_to_planes_kernel<<<grid, 128, 0, s>>>(x, bs, cs, ts, C, T, slope, hi, lo, rows, pad, fmt, 0, C / 8, 0, 0);
  return cudaGetLastError();
}

// max |w| over a weight tensor (one-off, at create time)
__global__ void tc_absmax_kernel(const float* __restrict__ w, size_t n, unsigned int* __restrict__ out) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(w[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));      // non-negative floats order like their bit patterns
}

cudaError_t tc_lo8_weights_fit(const float* w_ref, size_t n, cudaStream_t s, int* fits) {
  // The hi plane of an lo8 layer is fp16(w * 2^10): it must stay finite, with headroom (|w| < 32 -> < 32768)
  unsigned int* d = nullptr;
  cudaError_t e = cudaMalloc((void**)&d, sizeof(unsigned int));
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(d, 0, sizeof(unsigned int), s);
  if (e == cudaSuccess) {
    tc_absmax_kernel<<<64, 256, 0, s>>>(w_ref, n, d);
    e = cudaGetLastError();
  }
  unsigned int bits = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&bits, d, sizeof(bits), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(d);
  if (e != cudaSuccess) return e;
  float m;
  memcpy(&m, &bits, sizeof(m));
  *fits = (m < 32.f) ? 1 : 0;                                            // NaN compares false: no lo8 either
  return cudaSuccess;
}

cudaError_t tc_pack_weights_lo8(const float* w_ref, uint8_t* out, int C_out, int C_in, int K, int N, int KC, int fmt,
                                int pair, cudaStream_t s) {
  if (C_out % N || C_in % KC || KC % 16 || (pair && N % 2)) return cudaErrorInvalidValue;
  const size_t total = (size_t)C_out * C_in * K;
  tc_pack_weights_lo8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w_ref, out, C_out, C_in, K, N, KC, fmt, pair);
  return cudaGetLastError();
}

cudaError_t tc_to_planes_full(const float* x, long bs, long cs, long ts, int B, int C, int T, float slope, tc16* hi,
                              tc16* lo, int rows, int pad, int fmt, cudaStream_t s, int C_total, int c_off, int s2d_H) {
  if (C_total <= 0) C_total = C;
  if (C % 8 || C_total % 8 || c_off % 8 || c_off + C > C_total || (s2d_H && (cs == 1 || C % s2d_H))) return cudaErrorInvalidValue;
  dim3 grid(cdiv(rows, 128), C / 8, B);
  tc_to_planes_kernel<<<grid, 128, 0, s>>>(x, bs, cs, ts, C, T, slope, hi, lo, rows, pad, fmt, 1, C_total / 8, c_off / 8,
                                           s2d_H);
  return cudaGetLastError();
}

cudaError_t tc_zero_halo(tc16* hi, tc16* lo, int n_slabs_total, int rows, int pad, int T,
                         cudaStream_t s) {
  tc_zero_halo_kernel<<<n_slabs_total, 128, 0, s>>>(hi, lo, rows, pad, T);
  return cudaGetLastError();
}

cudaError_t tc_stream_to_nct(const float* st, float* out, int B, int C, int T, cudaStream_t s) {
  dim3 grid(cdiv(T, 128), C / 4, B);
  tc_stream_to_nct_kernel<<<grid, 128, 0, s>>>(st, out, C, T);
  return cudaGetLastError();
}
cudaError_t tc_nct_to_stream(const float* in, float* st, int B, int C, int T, cudaStream_t s) {
  dim3 grid(cdiv(T, 128), C / 4, B);
  tc_nct_to_stream_kernel<<<grid, 128, 0, s>>>(in, st, C, T);
  return cudaGetLastError();
}

cudaError_t tc_planes_to_nct(const tc16* hi, const tc16* lo, float* out, int B, int C, int T,
                             int rows, int pad, int fmt, cudaStream_t s) {
  dim3 grid(cdiv(T, 128), C / 8, B);
  tc_planes_to_nct_kernel<<<grid, 128, 0, s>>>(hi, lo, out, C, T, rows, pad, fmt);
  return cudaGetLastError();
}

cudaError_t tc_conv_post_finish(const float* part, const float* bias, float* wav, int B, int T, cudaStream_t s,
                                const int* lens, int len_mul) {
  dim3 grid(cdiv(T, 256), B);
  tc_conv_post_finish_kernel<<<grid, 256, 0, s>>>(part, bias, wav, T, lens, len_mul);
  return cudaGetLastError();
}

cudaError_t tc_conv_post(const float* st, const float* w, const float* bias, float* wav, int B, int C, int T, int K,
                         float slope, cudaStream_t s, const int* lens, int len_mul) {
  if (C % 4 || C > 64 || K > 16) return cudaErrorInvalidValue;
  dim3 grid(cdiv(T, 256), B);
  const size_t smem = (size_t)(C / 4) * (256 + K - 1) * sizeof(float4);
  tc_conv_post_kernel<16><<<grid, 256, smem, s>>>(st, w, bias, wav, C, T, K, slope, lens, len_mul);
  return cudaGetLastError();
}

}  // namespace dtts
