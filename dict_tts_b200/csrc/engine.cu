#include "engine.cuh"

namespace dtts {

static thread_local std::string g_err = "";

void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
const char* last_error_cstr() { return g_err.c_str(); }

int WeightTable::init(const float* arena_dev, uint64_t floats, const dtts_weight_entry* table, int n) {
  if (!arena_dev || !table || n <= 0) return fail(DTTS_ERR_BAD_ARG, "weight arena/table is null or empty");
  if (((uintptr_t)arena_dev & 15) != 0) return fail(DTTS_ERR_ALIGNMENT, "weight arena must be 16-byte aligned");
  arena = arena_dev;
  arena_floats = floats;
  for (int i = 0; i < n; ++i) {
    if (!table[i].name) return fail(DTTS_ERR_BAD_ARG, "weight table entry without a name");
    if (table[i].offset + table[i].numel > floats)
      return fail(DTTS_ERR_BAD_ARG, std::string("weight entry out of arena bounds: ") + table[i].name);
    entries[table[i].name] = std::make_pair(table[i].offset, table[i].numel);
  }
  return DTTS_OK;
}

const float* WeightTable::get(const std::string& name, uint64_t numel) {
  auto it = entries.find(name);
  if (it == entries.end()) {
    fail(DTTS_ERR_MISSING_WEIGHT, "missing weight: " + name);
    return nullptr;
  }
  if (it->second.second != numel) {
    fail(DTTS_ERR_BAD_SHAPE, "weight " + name + " has " + std::to_string(it->second.second) + " elements, expected " +
                                 std::to_string(numel));
    return nullptr;
  }
  return arena + it->second.first;
}

int Pool::reserve(size_t floats) {
  release();
  cudaError_t e = cudaMalloc((void**)&base, floats * sizeof(float));
  if (e != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("cudaMalloc(weight pool): ") + cudaGetErrorString(e));
  cap = floats;
  used = 0;
  return DTTS_OK;
}
float* Pool::take(size_t floats) {
  used = (used + 63) & ~(size_t)63;
  if (used + floats > cap) return nullptr;
  float* r = base + used;
  used += floats;
  return r;
}
void Pool::release() {
  if (base) cudaFree(base);
  base = nullptr;
  cap = used = 0;
}

ConvParams conv_params(const float* x, int T_in, const ConvW& w, int co_off, int co_n, float* out, int T_out, int dil,
                       int stride, int pad) {
  ConvParams p{};
  p.x = x; p.x_bs = (long)w.C_in * T_in; p.x_cs = T_in; p.x_ts = 1; p.C_in = w.C_in; p.T_in = T_in;
  p.w = w.w + co_off; p.w_ld = w.C_out; p.w_phase_stride = 0;
  p.bias = w.bias ? w.bias + co_off : nullptr;
  p.out = out; p.o_bs = (long)co_n * T_out; p.o_cs = T_out; p.o_ts = 1; p.C_out = co_n; p.T_out = T_out;
  p.res = nullptr; p.mask = nullptr;
  p.ktaps = w.ktaps; p.xs = stride; p.xd = dil; p.x0 = -pad;
  p.ot_mul = 1; p.ot_add = 0; p.nq = T_out; p.phases = 1;
  p.pre_slope = 1.f; p.act = ACT_NONE; p.alpha = 1.f; p.post = 1.f; p.accumulate = 0;
  return p;
}

ConvParams convT_params(const float* x, int T_in, const ConvW& w, float* out, int T_out, int stride, int pad) {
  ConvParams p{};
  p.x = x; p.x_bs = (long)w.C_in * T_in; p.x_cs = T_in; p.x_ts = 1; p.C_in = w.C_in; p.T_in = T_in;
  p.w = w.w; p.w_ld = w.C_out; p.w_phase_stride = (long)w.C_in * w.ktaps * w.C_out;
  p.bias = w.bias;
  p.out = out; p.o_bs = (long)w.C_out * T_out; p.o_cs = T_out; p.o_ts = 1; p.C_out = w.C_out; p.T_out = T_out;
  p.res = nullptr; p.mask = nullptr;
  // out[t = q*stride + phase - pad] = sum_m Wp[phase][ci][m][co] * x[ci, q - m]
  p.ktaps = w.ktaps; p.xs = 1; p.xd = -1; p.x0 = 0;
  p.ot_mul = stride; p.ot_add = -pad; p.nq = T_in + w.ktaps - 1; p.phases = stride;
  p.pre_slope = 1.f; p.act = ACT_NONE; p.alpha = 1.f; p.post = 1.f; p.accumulate = 0;
  return p;
}

int arch_check() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("cudaGetDevice: ") + cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(DTTS_ERR_UNSUPPORTED_ARCH, std::string("libdtts is built for sm_100a only; device is sm_") +
                                               std::to_string(prop.major) + std::to_string(prop.minor));
  e = conv1d_f32_init();
  if (e != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("conv1d_f32_init: ") + cudaGetErrorString(e));
  return DTTS_OK;
}

}  // namespace dtts

extern "C" int dtts_abi_version(void) { return DTTS_ABI_VERSION; }
namespace dtts { const char* last_error_cstr(); }
extern "C" const char* dtts_last_error(void) { return dtts::last_error_cstr(); }
