#include "tensormap.h"
#include "host_common.h"
#include <cudaTypedefs.h>
#include <mutex>
#include <unordered_map>
#include <memory>
#include <string>
#include <cstring>

namespace lsvs {
namespace {
struct Key {
  const void* base; uint64_t inner, outer, pitch; uint32_t bi, bo; uint64_t kind;
  bool operator==(const Key& o) const { return std::memcmp(this, &o, sizeof(Key)) == 0; }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    const uint64_t* p = reinterpret_cast<const uint64_t*>(&k);
    size_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(Key) / 8; ++i) h = (h ^ p[i]) * 1099511628211ull;
    return h;
  }
};
std::mutex g_mu;
// Two generations bound the cache: caller-owned tensors (tapped layers, per-forward outputs) bring a new address — a new key —
// on every call, so entries accumulate over long runs.  When the young generation reaches kGenLimit entries the old one is
// dropped and the young one takes its place; a hit in the old generation is promoted.  A pointer handed out stays valid for at
// least kGenLimit further insertions — a call encodes at most a handful of maps and passes them to its kernels by value.
using Map = std::unordered_map<Key, std::unique_ptr<CUtensorMap>, KeyHash>;
constexpr size_t kGenLimit = 2048;
Map g_young, g_old;
int g_device = -1;   // the library keeps per-process state (this cache, kernel attributes, SM count): one GPU per process
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
}  // namespace

namespace {
const CUtensorMap* tmap_2d(const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer,
                           CUtensorMapDataType dtype, int elem_bytes);
}
const CUtensorMap* tmap_2d_bf16(const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                                uint32_t box_inner, uint32_t box_outer) {
  return tmap_2d(base, inner, outer, pitch_bytes, box_inner, box_outer, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
}
const CUtensorMap* tmap_2d_f32_box32(const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_outer) {
  return tmap_2d(base, inner, outer, pitch_bytes, 32, box_outer, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4);
}
namespace {
const CUtensorMap* tmap_2d(const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer,
                           CUtensorMapDataType dtype, int elem_bytes) {
  Key key;
  std::memset(&key, 0, sizeof(key));
  key.base = base; key.inner = inner; key.outer = outer; key.pitch = pitch_bytes; key.bi = box_inner; key.bo = box_outer; key.kind = (uint64_t)dtype;
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  if (g_device < 0) g_device = dev;
  if (dev != g_device) {
    fail(LSVS_EUNSUPPORTED, "liblsvs_b200 was first used on CUDA device %d and is now called on device %d: one process per GPU "
         "(per-process tensor-map cache, kernel attributes and SM count)", g_device, dev);
    return nullptr;
  }
  auto it = g_young.find(key);
  if (it != g_young.end()) return it->second.get();
  it = g_old.find(key);
  if (it != g_old.end()) {   // promote
    const CUtensorMap* hit = it->second.get();
    g_young.emplace(key, std::move(it->second));
    g_old.erase(it);
    return hit;
  }
  if (g_young.size() >= kGenLimit) {
    g_old.swap(g_young);
    g_young.clear();
  }
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      fail(LSVS_ECUDA, "cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
      return nullptr;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  if (((uintptr_t)base & 15) || (pitch_bytes & 15) || box_inner * elem_bytes != 128 || box_outer > 256 || box_outer == 0) {
    fail(LSVS_EINVAL, "tensor map: base %p pitch %llu box %ux%u not TMA compatible", base, (unsigned long long)pitch_bytes,
         box_inner, box_outer);
    return nullptr;
  }
  auto tm = std::make_unique<CUtensorMap>();
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm.get(), dtype, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fail(LSVS_ECUDA, "cuTensorMapEncodeTiled failed (%d) dims %llux%llu pitch %llu box %ux%u", (int)r,
         (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_bytes, box_inner, box_outer);
    return nullptr;
  }
  const CUtensorMap* out = tm.get();
  g_young.emplace(key, std::move(tm));
  return out;
}
}  // namespace
}  // namespace lsvs
