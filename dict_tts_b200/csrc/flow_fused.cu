// flow_fused_kernel: the reverse pass of the FVAE prior flow in one launch (flow_fused.cuh).
//
// Per coupling layer e (execution order), with z = [z0 | z1] (which physical half is which follows the folded Flip):
//   x      = pre(z0)                                                       8 -> 64 channels, CUDA cores
//   for i in layers:  a = in_i(x) [k = 3] + cond_i(g) [1x1, 192 -> 128]    ONE accumulation of K = 3*64 + 192 in TMEM
//                     acts = tanh(a[:64]) * sigmoid(a[64:])                epilogue -> bf16 hi/lo operand tile in smem
//                     [x | skip] += rs_i(acts)                             accumulates in TMEM across the layers:
//                     x_{i+1} = x_0 + sum_{j<=i} res_j                       the epilogue adds x_0 and the bias prefix
//   z1    -= post(skip)                                                    64 -> 8 channels, CUDA cores
// Every product runs as three tcgen05.mma.kind::f16 (bf16 hi*hi + lo*hi + hi*lo, fp32 accumulate): the fp32-class
// arithmetic of the acoustic model (tc_conv.cuh), so z_p stays inside the same tolerance as the per-layer path.
//
// Tile: 128 rows (one MMA M) of one utterance; every convolution is k = 3 / dilation 1, so a row is exact as long as
// it is >= (rows of receptive field) away from a tile edge that is not a sequence edge: tiles overlap by 2 * halo rows
// (halo = couplings * layers) and store only their core rows; rows outside [0, T) are kept at zero in the x tile (the
// zero padding every layer's convolution sees).  cfg 2 (T/4 = 100) is one tile per utterance.
//
// 320 threads: warp 0 streams the weights (32 KB stages in consumption order, 3-deep ring) and loads the g tile;
// warp 1 owns TMEM (256 columns: a | x, skip) and one of its threads issues every MMA; warps 2-9 are the epilogue
// (TMEM lane quadrant = warp & 3, channel half = (warp - 2) / 4).  MMA and epilogue alternate strictly (acc_full /
// xa_ready barriers) except that the conditioning MMAs of layer i+1 -- they need only g -- are issued right behind the
// res_skip MMAs of layer i and run under its epilogue.
#include "flow_fused.cuh"

#include <cstdlib>
#include <mutex>

#include "tc16.cuh"
#include "tc_ptx.cuh"

namespace dtts {

namespace {

constexpr int kFH = 64;                          // WaveNet channels
constexpr int kFHalf = 8;                        // latent / 2
constexpr int kFN = 128;                         // MMA N: tanh | sigmoid halves, x | skip halves
constexpr int kFRows = 128;                      // tile rows
constexpr int kFXRows = kFRows + 2;              // x / acts tile: one zero row on both sides (k = 3)
constexpr int kFPlane = 8 * kFN * 16;            // one 64-channel plane of a weight stage: 16 KB
constexpr int kFStage = 2 * kFPlane;             // hi | lo
constexpr int kFWStages = 3;
constexpr int kFThreads = 64 + 8 * 32;
constexpr int kFXPlane = (kFH / 8) * kFXRows * 16;   // one plane of the x / acts tile
// mbarriers
constexpr int kFWFull = 0, kFWEmpty = kFWFull + kFWStages, kFGFull = kFWEmpty + kFWStages, kFXReady = kFGFull + 1,
              kFAccFull = kFXReady + 1, kFNumBars = kFAccFull + 1;
constexpr int kFTmemOff = kFNumBars * 8;
constexpr int kFHeader = 128;
static_assert(kFTmemOff + 4 <= kFHeader, "header");

struct FlowKernelArgs {
  FlowFusedParams p;
  const uint8_t* wstream;
  const float* par;
  int par_stride, n_flows, n_layers, n_chunks, H;
  uint32_t odd_mask;
  int ntiles, halo, core;
};

// parameter block of one coupling layer (floats)
__host__ __device__ constexpr int par_bias_a(int) { return 0; }                         // [L][128]  in_b + cond_b
__host__ __device__ constexpr int par_bias_x(int L) { return L * 128; }                 // [L][64]   prefix sums of res biases
__host__ __device__ constexpr int par_bias_s(int L) { return L * 192; }                 // [64]      sum of skip biases
__host__ __device__ constexpr int par_pre_w(int L) { return L * 192 + 64; }             // [8][64]
__host__ __device__ constexpr int par_pre_b(int L) { return L * 192 + 64 + 512; }       // [64]
__host__ __device__ constexpr int par_post_w(int L) { return L * 192 + 128 + 512; }     // [64][8]
__host__ __device__ constexpr int par_post_b(int L) { return L * 192 + 128 + 1024; }    // [8]
__host__ __device__ constexpr int par_floats(int L) { return (L * 192 + 128 + 1024 + 8 + 63) / 64 * 64; }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(kFThreads, 1) flow_fused_kernel(const FlowKernelArgs k) {
  extern __shared__ __align__(128) uint8_t smem[];
  const FlowFusedParams& p = k.p;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = k.n_layers;
  const int b = blockIdx.x / k.ntiles, tile = blockIdx.x - b * k.ntiles;
  const int s0 = tile * k.core - k.halo;                       // sequence position of tile row 0
  const uint32_t bar0 = smem_u32(smem);
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + kFTmemOff);
  const uint32_t g_plane = (uint32_t)(k.H / 8) * kFRows * 16u;
  const uint32_t x_base = smem_u32(smem + kFHeader);           // hi plane, lo plane
  const uint32_t g_base = x_base + 2u * kFXPlane;
  const uint32_t w_base = g_base + 2u * g_plane;
  // rows of the tile that exist in the sequence
  const int r_lo = s0 < 0 ? -s0 : 0;
  const int r_hi = p.T - s0 < kFRows ? p.T - s0 : kFRows;

  griddep_launch();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kFWStages; ++s) { mbar_init(bar(kFWFull + s), 1); mbar_init(bar(kFWEmpty + s), 1); }
    mbar_init(bar(kFGFull), 1);
    mbar_init(bar(kFXReady), 8);
    mbar_init(bar(kFAccFull), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem + kFTmemOff)),
                 "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    // zero the x / acts tile (its two margin rows stay zero for the life of the CTA) and the g rows no copy will fill
    for (uint32_t i = threadIdx.x; i < 2u * kFXPlane / 16u; i += kFThreads) st_shared_v4(x_base + i * 16u, 0, 0, 0, 0);
    const int nz = kFRows - (r_hi - r_lo);
    const int slabs2 = 2 * (k.H / 8);
    for (int i = threadIdx.x; i < nz * slabs2; i += kFThreads) {
      const int sl = i / nz, q = i - sl * nz;
      const int row = q < r_lo ? q : r_hi + (q - r_lo);
      st_shared_v4(g_base + (uint32_t)sl * kFRows * 16u + (uint32_t)row * 16u, 0, 0, 0, 0);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const int n_stage_layer = k.n_chunks + 4;

  if (warp == 0) {
    // ------------------------------------------------ producer: weight stages in consumption order, the g tile once
    if (elect_one()) {
      const int total = k.n_flows * L * n_stage_layer;
      int s = 0, issued = 0;
      uint32_t ph = 1;
      auto put = [&]() {
        mbar_wait(bar(kFWEmpty + s), ph);
        mbar_arrive_expect_tx(bar(kFWFull + s), kFStage);
        bulk_g2s(w_base + (uint32_t)s * kFStage, k.wstream + (size_t)issued * kFStage, kFStage, bar(kFWFull + s));
        ++issued;
        if (++s == kFWStages) { s = 0; ph ^= 1u; }
      };
      while (issued < total && issued < kFWStages) put();          // constants: may run ahead of the predecessor grid
      griddep_wait();
      {
        const uint32_t bytes = (uint32_t)(r_hi - r_lo) * 16u;
        const int slabs = k.H / 8;
        mbar_arrive_expect_tx(bar(kFGFull), 2u * (uint32_t)slabs * bytes);
        const size_t row0 = (size_t)(p.g_pad + s0 + r_lo);
        const tc16* gh = p.g_hi + (size_t)b * p.g_bs;
        const tc16* gl = p.g_lo + (size_t)b * p.g_bs;
        for (int sl = 0; sl < slabs; ++sl) {
          const uint32_t dst = g_base + (uint32_t)sl * kFRows * 16u + (uint32_t)r_lo * 16u;
          bulk_g2s(dst, gh + ((size_t)sl * p.g_rows + row0) * 8, bytes, bar(kFGFull));
          bulk_g2s(dst + g_plane, gl + ((size_t)sl * p.g_rows + row0) * 8, bytes, bar(kFGFull));
        }
      }
      while (issued < total) put();
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t hiw = (128u >> 4) | (1u << 14);                 // SBO = 128 B, descriptor version 1
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kFN >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t x_hi = ((x_base >> 4) & 0x3FFFu) | ((uint32_t)kFXRows << 16), x_lo = x_hi + (kFXPlane >> 4);
      const uint32_t g_hi = ((g_base >> 4) & 0x3FFFu) | ((uint32_t)kFRows << 16), g_lo = g_hi + (g_plane >> 4);
      const uint32_t w_lo0 = ((w_base >> 4) & 0x3FFFu) | ((uint32_t)kFN << 16);
      const uint32_t acc_a = tmem_base, acc_rs = tmem_base + 128u;
      int st = 0;
      uint32_t wph = 0, xph = 0;
      // one 64-channel K chunk: 4 K-steps x (hi*hi, lo*hi, hi*lo)
      auto chunk = [&](uint32_t a_hi, uint32_t a_lo, uint32_t a_kstep, uint32_t d, bool fresh) {
        mbar_wait(bar(kFWFull + st), wph);
        tc_fence_after();
        const uint32_t b_hi = w_lo0 + (uint32_t)st * (kFStage >> 4), b_lo = b_hi + (kFPlane >> 4);
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) {
          const uint64_t ah = desc64(a_hi + ks * a_kstep, hiw), al = desc64(a_lo + ks * a_kstep, hiw);
          const uint64_t bh = desc64(b_hi + ks * 2u * kFN, hiw), bl = desc64(b_lo + ks * 2u * kFN, hiw);
          umma_bf16(d, ah, bh, idesc, (fresh && ks == 0) ? 0u : 1u);
          umma_bf16(d, al, bh, idesc, 1u);
          umma_bf16(d, ah, bl, idesc, 1u);
        }
        umma_commit(bar(kFWEmpty + st));
        if (++st == kFWStages) { st = 0; wph ^= 1u; }
      };
      mbar_wait(bar(kFGFull), 0);
      for (int e = 0; e < k.n_flows; ++e) {
        for (int i = 0; i < L; ++i) {
          for (int c = 0; c < k.n_chunks; ++c)                        // conditioning: needs only g
            chunk(g_hi + (uint32_t)c * 8u * kFRows, g_lo + (uint32_t)c * 8u * kFRows, 2u * kFRows, acc_a, c == 0);
          mbar_wait(bar(kFXReady), xph); xph ^= 1u;                   // x tile written
          tc_fence_after();
          for (uint32_t j = 0; j < 3; ++j) chunk(x_hi + j, x_lo + j, 2u * kFXRows, acc_a, false);
          umma_commit(bar(kFAccFull));
          mbar_wait(bar(kFXReady), xph); xph ^= 1u;                   // acts written over the x tile
          tc_fence_after();
          chunk(x_hi + 1u, x_lo + 1u, 2u * kFXRows, acc_rs, i == 0);
          umma_commit(bar(kFAccFull));
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ epilogue: row r of the tile, channels [32 * half, 32 * half + 32)
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const int t = s0 + r;
    const bool inside = t >= 0 && t < p.T;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t x_row = x_base + (uint32_t)(half * 4) * kFXRows * 16u + (uint32_t)(1 + r) * 16u;
    // 32 channels of this row -> bf16 hi / lo planes of the x / acts tile
    auto write_tile = [&](const float (&v)[32], bool keep) {
#pragma unroll
      for (int sl = 0; sl < 4; ++sl) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          split2(v[8 * sl + 2 * e2], v[8 * sl + 2 * e2 + 1], 1, hw[e2], lw[e2]);
          if (!keep) hw[e2] = lw[e2] = 0u;
        }
        const uint32_t a = x_row + (uint32_t)sl * kFXRows * 16u;
        st_shared_v4(a, hw[0], hw[1], hw[2], hw[3]);
        st_shared_v4(a + kFXPlane, lw[0], lw[1], lw[2], lw[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kFXReady));
    };
    griddep_wait();
    float za[kFHalf], zb[kFHalf];                                      // physical latent channels [0, 8) and [8, 16)
    {
      const float* zp = p.z_in + (size_t)b * 2 * kFHalf * p.T + (inside ? t : 0);
#pragma unroll
      for (int j = 0; j < kFHalf; ++j) {
        za[j] = inside ? zp[(size_t)j * p.T] : 0.f;
        zb[j] = inside ? zp[(size_t)(kFHalf + j) * p.T] : 0.f;
      }
    }
    float x0[32];
    uint32_t aph = 0;
    for (int e = 0; e < k.n_flows; ++e) {
      const float* par = k.par + (size_t)e * k.par_stride;
      const bool odd = (k.odd_mask >> e) & 1u;
      {
        // x = pre(z0): the conditioning half is physical [8, 16) on flipped layers
        const float4* pb = reinterpret_cast<const float4*>(par + par_pre_b(L) + half * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = __ldg(pb + q);
          x0[4 * q] = v.x; x0[4 * q + 1] = v.y; x0[4 * q + 2] = v.z; x0[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < kFHalf; ++j) {
          const float c = odd ? zb[j] : za[j];
          const float4* pw = reinterpret_cast<const float4*>(par + par_pre_w(L) + j * kFH + half * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 w = __ldg(pw + q);
            x0[4 * q] = fmaf(w.x, c, x0[4 * q]); x0[4 * q + 1] = fmaf(w.y, c, x0[4 * q + 1]);
            x0[4 * q + 2] = fmaf(w.z, c, x0[4 * q + 2]); x0[4 * q + 3] = fmaf(w.w, c, x0[4 * q + 3]);
          }
        }
        write_tile(x0, inside);
      }
      for (int i = 0; i < L; ++i) {
        {
          // gate: acts = tanh(a[:64]) * sigmoid(a[64:])
          mbar_wait(bar(kFAccFull), aph); aph ^= 1u;
          tc_fence_after();
          uint32_t ta[32], sg[32];
          __syncwarp();
          tmem_ld32_nowait(lane_addr + (uint32_t)(half * 32), ta);
          tmem_ld32_nowait(lane_addr + (uint32_t)(kFH + half * 32), sg);
          tmem_ld_wait();
          const float4* bt = reinterpret_cast<const float4*>(par + par_bias_a(L) + i * kFN + half * 32);
          const float4* bs = reinterpret_cast<const float4*>(par + par_bias_a(L) + i * kFN + kFH + half * 32);
          float v[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b0 = __ldg(bt + q), b1 = __ldg(bs + q);
            const float tb[4] = {b0.x, b0.y, b0.z, b0.w}, sb[4] = {b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float tv = __uint_as_float(ta[4 * q + u]) + tb[u];
              const float sv = __uint_as_float(sg[4 * q + u]) + sb[u];
              v[4 * q + u] = gate_fast(tv, sv);
            }
          }
          write_tile(v, true);
        }
        mbar_wait(bar(kFAccFull), aph); aph ^= 1u;
        tc_fence_after();
        if (i < L - 1) {
          // x_{i+1} = x_0 + sum_{j <= i} res_j (accumulated in TMEM) + the prefix sum of their biases
          uint32_t xr[32];
          __syncwarp();
          tmem_ld32_nowait(lane_addr + 128u + (uint32_t)(half * 32), xr);
          tmem_ld_wait();
          const float4* bx = reinterpret_cast<const float4*>(par + par_bias_x(L) + i * kFH + half * 32);
          float v[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b0 = __ldg(bx + q);
            v[4 * q] = x0[4 * q] + (__uint_as_float(xr[4 * q]) + b0.x);
            v[4 * q + 1] = x0[4 * q + 1] + (__uint_as_float(xr[4 * q + 1]) + b0.y);
            v[4 * q + 2] = x0[4 * q + 2] + (__uint_as_float(xr[4 * q + 2]) + b0.z);
            v[4 * q + 3] = x0[4 * q + 3] + (__uint_as_float(xr[4 * q + 3]) + b0.w);
          }
          write_tile(v, inside);
        } else {
          // m = post(skip); z1 -= m (mean-only coupling, logs = 0).  Both channel halves keep their own copy of z.
          uint32_t s1[32], s2[32];
          __syncwarp();
          tmem_ld32_nowait(lane_addr + 128u + (uint32_t)kFH, s1);
          tmem_ld32_nowait(lane_addr + 128u + (uint32_t)kFH + 32u, s2);
          tmem_ld_wait();
          tc_fence_before();
          float m[kFHalf];
          {
            const float4* pb = reinterpret_cast<const float4*>(par + par_post_b(L));
            const float4 b0 = __ldg(pb), b1 = __ldg(pb + 1);
            m[0] = b0.x; m[1] = b0.y; m[2] = b0.z; m[3] = b0.w; m[4] = b1.x; m[5] = b1.y; m[6] = b1.z; m[7] = b1.w;
          }
          const float* bsk = par + par_bias_s(L);
          const float4* pw = reinterpret_cast<const float4*>(par + par_post_w(L));
#pragma unroll
          for (int c = 0; c < kFH; ++c) {
            const float sk = __uint_as_float(c < 32 ? s1[c & 31] : s2[c & 31]) + __ldg(bsk + c);
            const float4 w0 = __ldg(pw + 2 * c), w1 = __ldg(pw + 2 * c + 1);
            m[0] = fmaf(w0.x, sk, m[0]); m[1] = fmaf(w0.y, sk, m[1]); m[2] = fmaf(w0.z, sk, m[2]); m[3] = fmaf(w0.w, sk, m[3]);
            m[4] = fmaf(w1.x, sk, m[4]); m[5] = fmaf(w1.y, sk, m[5]); m[6] = fmaf(w1.z, sk, m[6]); m[7] = fmaf(w1.w, sk, m[7]);
          }
#pragma unroll
          for (int j = 0; j < kFHalf; ++j) {
            if (odd) za[j] -= m[j];
            else zb[j] -= m[j];
          }
        }
      }
    }
    // core rows only (a tile edge that is not a sequence edge has seen its neighbours' rows as zeros)
    const bool first = tile == 0, last = tile == k.ntiles - 1;
    const bool mine = inside && (first || r >= k.halo) && (last || r < k.halo + k.core);
    if (mine) {
      float* zo = p.z_out + (size_t)b * 2 * kFHalf * p.T + t;
#pragma unroll
      for (int j = 0; j < kFHalf; ++j) {
        if (half == 0) zo[(size_t)j * p.T] = za[j];
        else zo[(size_t)(kFHalf + j) * p.T] = zb[j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// create-time packing
__global__ void flow_pack_layer_kernel(const float* __restrict__ in_w, const float* __restrict__ cond_w,
                                       const float* __restrict__ rs_w, int rs_rows, int H, int n_chunks,
                                       uint8_t* __restrict__ out) {
  const int per_stage = 8 * kFN * 8;                                   // elements of one plane
  const int n = (n_chunks + 4) * per_stage;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int stage = i / per_stage, q = i - stage * per_stage;
    const int sl = q / (kFN * 8), nn = (q / 8) % kFN, e = q & 7;
    const int kk = sl * 8 + e;                                         // channel inside the 64-channel chunk
    float w;
    if (stage < n_chunks) {
      w = cond_w[(size_t)nn * H + stage * 64 + kk];
    } else if (stage < n_chunks + 3) {
      w = in_w[((size_t)nn * kFH + kk) * 3 + (stage - n_chunks)];
    } else if (rs_rows == kFN) {
      w = rs_w[(size_t)nn * kFH + kk];                                 // [0, 64): res -> x, [64, 128): skip
    } else {
      w = nn < kFH ? 0.f : rs_w[(size_t)(nn - kFH) * kFH + kk];        // last layer: everything goes to the skip sum
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    tc16* dst = reinterpret_cast<tc16*>(out + (size_t)stage * kFStage);
    dst[q] = __bfloat16_as_ushort(hi);
    dst[per_stage + q] = __bfloat16_as_ushort(lo);
  }
}

__global__ void flow_pack_par_kernel(const FlowParSrc s, float* __restrict__ out) {
  const int L = s.n_layers;
  for (int i = threadIdx.x; i < L * kFN; i += blockDim.x) {
    const int l = i / kFN, c = i - l * kFN;
    out[par_bias_a(L) + i] = s.in_b[l][c] + s.cond_b[l * kFN + c];
  }
  for (int c = threadIdx.x; c < kFH; c += blockDim.x) {
    float ax = 0.f, as = 0.f;
    for (int l = 0; l < L; ++l) {
      if (l < L - 1) {
        ax += s.rs_b[l][c];
        as += s.rs_b[l][kFH + c];
      } else {
        as += s.rs_b[l][c];
      }
      out[par_bias_x(L) + l * kFH + c] = ax;
    }
    out[par_bias_s(L) + c] = as;
    out[par_pre_b(L) + c] = s.pre_b[c];
  }
  for (int i = threadIdx.x; i < kFHalf * kFH; i += blockDim.x) {
    out[par_pre_w(L) + i] = s.pre_w[i];                                 // [8][64]
    out[par_post_w(L) + i] = s.post_w[i];                               // [64][8]
  }
  for (int i = threadIdx.x; i < kFHalf; i += blockDim.x) out[par_post_b(L) + i] = s.post_b[i];
}

size_t flow_smem_bytes(int H) { return (size_t)kFHeader + 2 * kFXPlane + 2 * (size_t)(H / 8) * kFRows * 16 + kFWStages * kFStage; }

}  // namespace

int flow_fused_supported(int H, int flow_hidden, int latent, int flow_kernel, int n_layers, int n_flows) {
  if (flow_hidden != kFH || latent != 2 * kFHalf || flow_kernel != 3) return 0;
  if (H <= 0 || H % 64 || n_layers < 1 || n_layers > 8 || n_flows < 1 || n_flows > 16) return 0;
  if (n_layers * n_flows > 32) return 0;                                // halo: the core of a 128-row tile stays >= 64 rows
  return flow_smem_bytes(H) <= (size_t)227 * 1024;
}
int flow_fused_par_floats(int n_layers) { return par_floats(n_layers); }

cudaError_t flow_fused_pack_layer(const float* in_w, const float* cond_w_rows, const float* rs_w, int rs_rows, int H,
                                  uint8_t* out, cudaStream_t s) {
  if ((rs_rows != kFN && rs_rows != kFH) || H % 64) return cudaErrorInvalidValue;
  flow_pack_layer_kernel<<<64, 256, 0, s>>>(in_w, cond_w_rows, rs_w, rs_rows, H, H / 64, out);
  return cudaGetLastError();
}
cudaError_t flow_fused_pack_par(const FlowParSrc& src, float* out, cudaStream_t s) {
  if (src.n_layers < 1 || src.n_layers > 8) return cudaErrorInvalidValue;
  flow_pack_par_kernel<<<1, 256, 0, s>>>(src, out);
  return cudaGetLastError();
}

cudaError_t launch_flow_fused(const FlowFusedW& w, const FlowFusedParams& p, cudaStream_t stream) {
  if (p.B <= 0 || p.T <= 0) return cudaSuccess;
  if (!w.ready() || !p.g_hi || !p.g_lo || !p.z_in || !p.z_out) return cudaErrorInvalidValue;
  FlowKernelArgs k;
  k.p = p;
  k.wstream = w.stream; k.par = w.par; k.par_stride = w.par_stride;
  k.n_flows = w.n_flows; k.n_layers = w.n_layers; k.n_chunks = w.n_chunks; k.H = w.H; k.odd_mask = w.odd_mask;
  k.halo = w.n_flows * w.n_layers;                                     // k = 3, dilation 1: one row per layer
  k.core = kFRows - 2 * k.halo;
  // the first and the last tile keep their sequence-edge rows: T <= 128 - halo fits one tile
  k.ntiles = p.T <= kFRows - k.halo ? 1 : 1 + cdiv(p.T - (kFRows - k.halo), k.core);
  const size_t smem = flow_smem_bytes(w.H);
  static unsigned long long attr_done = 0;
  const cudaError_t attr_err = ensure_max_dyn_smem(flow_fused_kernel, 227 * 1024, &attr_done);
  if (attr_err != cudaSuccess) return attr_err;
  if (smem > (size_t)227 * 1024) return cudaErrorInvalidConfiguration;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3((unsigned)(p.B * k.ntiles));
  cfg.blockDim = dim3(kFThreads);
  cfg.dynamicSmemBytes = smem < 116 * 1024 ? 116 * 1024 : smem;        // alone on its SM
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = tc_pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, flow_fused_kernel, k);
}

static int g_ac_fuse_override = -1;
void ac_fuse_override(int v) { g_ac_fuse_override = v; }
int ac_fuse_enabled() {
  static const int v = [] {
    const char* e = getenv("DTTS_AC_FUSE");
    return (e && *e) ? (atoi(e) != 0) : 1;
  }();
  return g_ac_fuse_override >= 0 ? (g_ac_fuse_override != 0) : v;
}

}  // namespace dtts
