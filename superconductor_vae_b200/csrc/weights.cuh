// Engine-owned weight storage keyed by the reference's state_dict names.
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace scv {

struct Lin {               // nn.Linear: weight [N, K] bf16 (row stride ldw, zero padded), bias [N] fp32
  __nv_bfloat16* w = nullptr;
  __nv_bfloat16* wt = nullptr;   // tiled + swizzled copy for the tcgen05 GEMM (null when the shape never uses it)
  float* b = nullptr;
  int N = 0, K = 0, ldw = 0;
};
struct LNp {               // nn.LayerNorm
  float* g = nullptr;
  float* b = nullptr;
  int N = 0;
};

class WeightStore {
 public:
  ~WeightStore();
  // registration (allocates device memory); returns nullptr and sets the error on failure
  __nv_bfloat16* add_matrix(const std::string& name, int rows, int cols, int* ld_out);
  // extra tcgen05-tiled copy of rows [row0, row0 + rows) of an already registered matrix
  __nv_bfloat16* add_tiled_view(const std::string& name, int row0, int rows);
  float* add_vector(const std::string& name, int64_t numel);
  int add_linear(const std::string& prefix, int N, int K, Lin* out, bool bias = true, bool tiled = false);
  int add_layernorm(const std::string& prefix, int N, LNp* out);
  // optional entries are accepted by load() but not required by missing()
  void mark_optional(const std::string& name);
  int load(const char* name, const float* src, int64_t numel, cudaStream_t s);
  int missing(std::string* first) const;
  bool loaded(const std::string& name) const;

 private:
  struct TiledView { __nv_bfloat16* dst; int row0, rows; };
  struct Slot {
    void* dst = nullptr;
    bool is_matrix = false;
    int rows = 0, cols = 0, ld = 0;
    int64_t numel = 0;
    bool loaded = false, optional = false;
    std::vector<TiledView> tiled;
  };
  std::unordered_map<std::string, Slot> slots_;
  std::vector<std::string> order_;
  std::vector<void*> allocs_;
};

}  // namespace scv
