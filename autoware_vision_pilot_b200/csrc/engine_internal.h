// engine_internal.h — host-side pieces shared by the engines (engine.cu: the four segmentation / depth / lane
// networks; autospeed.cu: the AutoSpeed detector): the .vpw weight-file reader, shape-checked lookups, K-major
// repacking, the device guard.
#pragma once
#include <cuda_runtime.h>
#include <initializer_list>
#include <map>
#include <string>
#include <vector>

namespace vpb {

struct HostTensor {
  std::vector<int> dims;
  std::vector<float> f;
  size_t numel() const { size_t n = 1; for (int d : dims) n *= d; return n; }
};
using WeightMap = std::map<std::string, HostTensor>;

int load_vpw(const char* path, WeightMap& out);
const HostTensor* find_w(const WeightMap& w, const std::string& key);
// dims: expected shape, -1 = any; sets vpb_last_error and returns NULL on a mismatch
const HostTensor* find_w_shaped(const WeightMap& w, const std::string& key, std::initializer_list<int> dims);
// Conv2d weight [Cout][Cin][k][k] -> [k*k][Cout][Cin] (optionally scaled per Cout)
std::vector<float> pack_conv(const HostTensor& t, const std::vector<float>* scale);

// RAII: make the engine's device current for the duration of a C-ABI call and restore the caller's device
// afterwards (several engines / threads / GPUs may share one process).
struct DeviceGuard {
  int prev = -1; bool changed = false;
  explicit DeviceGuard(int d) { if (cudaGetDevice(&prev) == cudaSuccess && prev != d) changed = cudaSetDevice(d) == cudaSuccess; }
  ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
};

}  // namespace vpb
