// Host-side plumbing shared by the acoustic and vocoder handles: error reporting, weight lookup / re-layout pool,
// workspace bump allocator and convolution descriptors.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/dtts.h"
#include "conv1d_f32.cuh"
#include "kernels.cuh"

namespace dtts {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define DTTS_CUDA(expr)                                                                       \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return ::dtts::fail(DTTS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

#define DTTS_TRY(expr)          \
  do {                          \
    int _r = (expr);            \
    if (_r != DTTS_OK) return _r; \
  } while (0)

// Packed convolution weights living in the handle's pool.
struct ConvW {
  const float* w = nullptr;     // [phases][C_in][ktaps][C_out]
  const float* bias = nullptr;  // [C_out] or null
  int C_out = 0, C_in = 0, ktaps = 0, phases = 1;
};

struct WeightTable {
  const float* arena = nullptr;
  uint64_t arena_floats = 0;
  std::map<std::string, std::pair<uint64_t, uint64_t>> entries;   // name -> (offset, numel)
  int init(const float* arena_dev, uint64_t floats, const dtts_weight_entry* table, int n);
  // returns null and sets the error if missing or of the wrong size
  const float* get(const std::string& name, uint64_t numel);
};

// Device pool owned by a handle (the only device allocation the library makes).
struct Pool {
  float* base = nullptr;
  size_t cap = 0, used = 0;
  int reserve(size_t floats);
  float* take(size_t floats);
  void release();
};

struct Bump {
  char* base;
  size_t cap, off = 0;
  bool ok = true;
  Bump(void* p, size_t n) : base((char*)p), cap(n) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    const size_t bytes = n * sizeof(T);
    if (off + bytes > cap) { ok = false; return nullptr; }
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
};
static inline size_t ws_round(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

struct Launcher {
  cudaStream_t stream = nullptr;
  uint64_t* counter = nullptr;
  cudaError_t err = cudaSuccess;
  void operator()(cudaError_t e) {
    if (counter) ++*counter;
    if (err == cudaSuccess && e != cudaSuccess) err = e;
  }
};

// Standard Conv1d: x [B,C_in,T_in] -> out [B,C_out',T_out] using output channels [co_off, co_off+co_n) of w.
ConvParams conv_params(const float* x, int T_in, const ConvW& w, int co_off, int co_n, float* out, int T_out, int dil,
                       int stride, int pad);
// ConvTranspose1d with kernel = ktaps*stride: x [B,C_in,T_in] -> out [B,C_out,T_out]
ConvParams convT_params(const float* x, int T_in, const ConvW& w, float* out, int T_out, int stride, int pad);

int arch_check();

}  // namespace dtts
