// Error reporting, launch accounting, weight store and the kernel-level C-ABI taps.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/scvae_b200.h"
#include "common.cuh"
#include "weights.cuh"

namespace scv {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_total() { return g_launches.load(); }
// PDL: every step kernel carries the programmatic-stream-serialization attribute and triggers its dependents late
// (GEMM: when its main loop is over; attention / LayerNorm / sampler: when their loads are in), so the next kernel's
// launch latency and prologue overlap the tail of the running one.  Measured on B200: -9 % per step at 32-256 rows,
// -2.5 % at 4096 rows.  (Triggering at kernel entry was 3 % SLOWER at 4096 rows: early dependents took shared memory
// and TMEM from CTAs of the running grid that had not been scheduled yet.)  SCV_PDL=0/1 forces it.
static bool g_pdl_call = true;
void set_pdl_for_call(bool on) { g_pdl_call = on; }
bool pdl_enabled() {
  static const int forced = [] { const char* e = getenv("SCV_PDL"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
  return forced >= 0 ? forced == 1 : g_pdl_call;
}

// ------------------------------------------------------------------ tunables
static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
static unsigned g_tune_epoch = 0;
Tunables& tun() {
  static Tunables t = {env_int("SCV_ATTN_CTAS", 0), env_int("SCV_GEMM_STAGES", 0), env_int("SCV_SUBBATCHES", 0),
                       env_int("SCV_GRAPH", 1), env_int("SCV_SUB_MIN_ROWS", 2048), env_int("SCV_ATTN_BULK", 0),
                       env_int("SCV_ATTN_BULK_MIN_ROWS", 256), env_int("SCV_ATTN_BULK_PIECE_KB", 0),
                       env_int("SCV_COND_TC_MIN_ROWS", 16384), env_int("SCV_CLUSTER", 0), env_int("SCV_CLUSTER_MAX_ROWS", 64), env_int("SCV_CLUSTER_ROWS", 0),
                       env_int("SCV_GEMM_BN64", 8), env_int("SCV_GEMM_BN64_MAX_CTAS", 148),
                       env_int("SCV_GEMM_MC", 0), env_int("SCV_GEMM_MC_MIN_ROW_TILES", 9), env_int("SCV_GEMM_MC_MIN_KBLOCKS", 1), env_int("SCV_ATTN_PAGES_REGS", 1), env_int("SCV_ATTN_FORWARD", 1), env_int("SCV_ATTN_FORWARD_MIN_CTAS", 1024),
                       env_int("SCV_ATTN_SHARED", 1)};
  return t;
}
unsigned tune_epoch() { return g_tune_epoch; }
int sm_count() {
  static int cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ------------------------------------------------------------------ CTA residency trace
static void* g_trace_dev = nullptr;
static unsigned g_trace_cap = 0;
static bool g_trace_on = false;
void* trace_ptr() { return g_trace_on ? g_trace_dev : nullptr; }

// ------------------------------------------------------------------ profiling
struct ProfRec { int cat; cudaEvent_t a, b; double flops, bytes; };
static bool g_prof = false;
static bool g_recs_pending = false;                 // records whose events have been (re)recorded since the last harvest
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_event_pool;
static int g_acc_n[PC_COUNT];
static double g_acc_ms[PC_COUNT], g_acc_flops[PC_COUNT], g_acc_bytes[PC_COUNT];
bool prof_enabled() { return g_prof; }
static cudaEvent_t get_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
ProfScope::ProfScope(int cat, cudaStream_t s, double flops, double bytes) : idx_(-1), s_(s) {
  if (!g_prof) return;
  ProfRec r{cat, get_event(), get_event(), flops, bytes};
  cudaEventRecord(r.a, s);
  idx_ = (int)g_recs.size();
  g_recs.push_back(r);
  g_recs_pending = true;
}
ProfScope::~ProfScope() {
  if (idx_ >= 0) cudaEventRecord(g_recs[idx_].b, s_);
}
void prof_step_begin() {
  for (auto& r : g_recs) { g_event_pool.push_back(r.a); g_event_pool.push_back(r.b); }
  g_recs.clear();
  g_recs_pending = false;
}
void prof_mark_pending() { g_recs_pending = !g_recs.empty(); }
int prof_harvest() {           // the caller has synchronised the stream(s) the records were taken on
  if (!g_recs_pending) return 0;
  for (auto& r : g_recs) {
    float t = 0.f;
    SCV_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    g_acc_n[r.cat] += 1; g_acc_ms[r.cat] += t; g_acc_flops[r.cat] += r.flops; g_acc_bytes[r.cat] += r.bytes;
  }
  g_recs_pending = false;
  return 0;
}

// ------------------------------------------------------------------ WeightStore
WeightStore::~WeightStore() {
  for (void* p : allocs_) cudaFree(p);
}

__nv_bfloat16* WeightStore::add_matrix(const std::string& name, int rows, int cols, int* ld_out) {
  const int ld = round_up(cols, 8);      // 16-byte rows for vector loads
  void* p = nullptr;
  const size_t bytes = (size_t)rows * ld * sizeof(__nv_bfloat16);
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    set_error("cudaMalloc(%zu) failed for %s", bytes, name.c_str());
    return nullptr;
  }
  cudaMemset(p, 0, bytes);
  allocs_.push_back(p);
  Slot s;
  s.dst = p; s.is_matrix = true; s.rows = rows; s.cols = cols; s.ld = ld; s.numel = (int64_t)rows * cols;
  slots_[name] = s;
  order_.push_back(name);
  if (ld_out) *ld_out = ld;
  return static_cast<__nv_bfloat16*>(p);
}

float* WeightStore::add_vector(const std::string& name, int64_t numel) {
  void* p = nullptr;
  if (cudaMalloc(&p, (size_t)numel * sizeof(float)) != cudaSuccess) {
    set_error("cudaMalloc failed for %s", name.c_str());
    return nullptr;
  }
  allocs_.push_back(p);
  Slot s;
  s.dst = p; s.numel = numel;
  slots_[name] = s;
  order_.push_back(name);
  return static_cast<float*>(p);
}

__nv_bfloat16* WeightStore::add_tiled_view(const std::string& name, int row0, int rows) {
  auto it = slots_.find(name);
  if (it == slots_.end() || !it->second.is_matrix) {
    set_error("add_tiled_view: %s is not a registered matrix", name.c_str());
    return nullptr;
  }
  void* p = nullptr;
  const size_t bytes = tc_packed_elems(rows, it->second.cols) * sizeof(__nv_bfloat16);
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    set_error("cudaMalloc(%zu) failed for the tiled copy of %s", bytes, name.c_str());
    return nullptr;
  }
  allocs_.push_back(p);
  it->second.tiled.push_back({static_cast<__nv_bfloat16*>(p), row0, rows});
  return static_cast<__nv_bfloat16*>(p);
}

int WeightStore::add_linear(const std::string& prefix, int N, int K, Lin* out, bool bias, bool tiled) {
  out->N = N; out->K = K;
  out->w = add_matrix(prefix + ".weight", N, K, &out->ldw);
  if (!out->w) return 2;
  if (tiled) {
    out->wt = add_tiled_view(prefix + ".weight", 0, N);
    if (!out->wt) return 2;
  }
  if (bias) {
    out->b = add_vector(prefix + ".bias", N);
    if (!out->b) return 2;
  }
  return 0;
}

int WeightStore::add_layernorm(const std::string& prefix, int N, LNp* out) {
  out->N = N;
  out->g = add_vector(prefix + ".weight", N);
  out->b = add_vector(prefix + ".bias", N);
  return (out->g && out->b) ? 0 : 2;
}

void WeightStore::mark_optional(const std::string& name) {
  auto it = slots_.find(name);
  if (it != slots_.end()) it->second.optional = true;
}

int WeightStore::load(const char* name, const float* src, int64_t numel, cudaStream_t s) {
  auto it = slots_.find(name);
  SCV_REQUIRE(it != slots_.end(), "unknown state_dict key '%s'", name);
  Slot& sl = it->second;
  SCV_REQUIRE(sl.numel == numel, "state_dict key '%s': expected %lld elements, got %lld", name,
              (long long)sl.numel, (long long)numel);
  if (sl.is_matrix) {
    SCV_TRY(launch_pack_bf16(src, static_cast<__nv_bfloat16*>(sl.dst), sl.rows, sl.cols, sl.ld, s));
    for (const TiledView& tv : sl.tiled)
      SCV_TRY(launch_pack_tiled(src + (size_t)tv.row0 * sl.cols, tv.dst, tv.rows, sl.cols, s));
  } else {
    SCV_TRY(launch_copy_f32(src, static_cast<float*>(sl.dst), numel, s));
  }
  sl.loaded = true;
  return 0;
}

bool WeightStore::loaded(const std::string& name) const {
  auto it = slots_.find(name);
  return it != slots_.end() && it->second.loaded;
}

int WeightStore::missing(std::string* first) const {
  int n = 0;
  for (const auto& name : order_) {
    const Slot& s = slots_.at(name);
    if (!s.loaded && !s.optional) {
      if (n == 0 && first) *first = name;
      ++n;
    }
  }
  return n;
}

// impl 0: tensor cores whenever the shape allows (M >= 64, K >= 64, 16-byte aligned rows), else CUDA cores.
// SCV_LINEAR_IMPL=1 in the environment forces the CUDA-core path everywhere (A/B testing of the two paths).
int launch_linear(const LinearArgs& a, int impl, cudaStream_t s) {
  static const int forced = [] { const char* e = getenv("SCV_LINEAR_IMPL"); return e ? atoi(e) : 0; }();
  if (a.a_split != nullptr || a.y_split != nullptr) return launch_linear_tcgen05(a, s);   // no fp32 copy exists
  if (impl == 0) impl = forced;
  if (impl == 2) return launch_linear_tcgen05(a, s);
  if (impl == 0 && tc_shape_ok(a)) return launch_linear_tcgen05(a, s);
  return launch_linear_simt(a, s);
}

}  // namespace scv

using namespace scv;

extern "C" {

int scv_abi_version(void) { return SCV_ABI_VERSION; }
const char* scv_last_error(void) { return g_err; }
int64_t scv_launch_count(void) { return g_launches.load(); }

int scv_tune(const char* key, int32_t value) {
  SCV_REQUIRE(key != nullptr, "tune: null key");
  Tunables& t = tun();
  const std::string k(key);
  int* slot = k == "attn_ctas_per_sm" ? &t.attn_ctas_per_sm : k == "gemm_stages" ? &t.gemm_stages :
              k == "subbatches" ? &t.subbatches : k == "graph" ? &t.graph : k == "sub_min_rows" ? &t.sub_min_rows : k == "attn_bulk" ? &t.attn_bulk :
              k == "attn_bulk_min_rows" ? &t.attn_bulk_min_rows : k == "attn_bulk_piece_kb" ? &t.attn_bulk_piece_kb :
              k == "cond_tc_min_rows" ? &t.cond_tc_min_rows : k == "cluster" ? &t.cluster : k == "cluster_max_rows" ? &t.cluster_max_rows : k == "cluster_rows" ? &t.cluster_rows :
              k == "attn_shared" ? &t.attn_shared : k == "gemm_bn64" ? &t.gemm_bn64 :
              k == "gemm_bn64_max_ctas" ? &t.gemm_bn64_max_ctas : k == "gemm_mc" ? &t.gemm_mc :
              k == "gemm_mc_min_row_tiles" ? &t.gemm_mc_min_row_tiles : k == "gemm_mc_min_kblocks" ? &t.gemm_mc_min_kblocks :
              k == "attn_pages_regs" ? &t.attn_pages_regs : k == "attn_forward" ? &t.attn_forward :
              k == "attn_forward_min_ctas" ? &t.attn_forward_min_ctas : nullptr;
  SCV_REQUIRE(slot != nullptr, "tune: unknown key '%s'", key);
  if (*slot != value) { *slot = value; ++g_tune_epoch; }
  return 0;
}

int scv_trace_begin(int32_t max_records) {
  SCV_REQUIRE(max_records > 0, "trace_begin: max_records must be positive");
  if (g_trace_dev == nullptr || g_trace_cap < (unsigned)max_records) {
    if (g_trace_dev) cudaFree(g_trace_dev);
    g_trace_dev = nullptr;
    SCV_CUDA(cudaMalloc(&g_trace_dev, sizeof(TraceBuf) + (size_t)max_records * sizeof(TraceRec)));
    g_trace_cap = (unsigned)max_records;
  }
  TraceBuf h{0u, g_trace_cap, 0ull};
  SCV_CUDA(cudaMemcpy(g_trace_dev, &h, sizeof(h), cudaMemcpyHostToDevice));
  g_trace_on = true;
  ++g_tune_epoch;             // captured graphs hold the old (null) trace pointer
  return 0;
}

int scv_trace_read(void* records_host, int32_t max_records, int32_t* n_out) {
  SCV_REQUIRE(g_trace_dev != nullptr && records_host != nullptr && n_out != nullptr, "trace_read: nothing traced");
  SCV_CUDA(cudaDeviceSynchronize());
  g_trace_on = false;
  ++g_tune_epoch;
  TraceBuf h;
  SCV_CUDA(cudaMemcpy(&h, g_trace_dev, sizeof(h), cudaMemcpyDeviceToHost));
  const unsigned n = std::min(std::min(h.n, h.cap), (unsigned)max_records);
  SCV_CUDA(cudaMemcpy(records_host, static_cast<char*>(g_trace_dev) + sizeof(TraceBuf), (size_t)n * sizeof(TraceRec),
                      cudaMemcpyDeviceToHost));
  *n_out = (int32_t)n;
  return 0;
}

int scv_profile_begin(void) {
  prof_step_begin();
  for (int i = 0; i < PC_COUNT; ++i) { g_acc_n[i] = 0; g_acc_ms[i] = g_acc_flops[i] = g_acc_bytes[i] = 0.0; }
  g_prof = true;
  return 0;
}

int scv_profile_end(int32_t n_cats, int32_t* counts, double* ms, double* flops, double* bytes) {
  g_prof = false;
  SCV_REQUIRE(n_cats >= PC_COUNT && counts && ms && flops && bytes, "profile_end: need room for %d categories", (int)PC_COUNT);
  SCV_CUDA(cudaDeviceSynchronize());
  SCV_TRY(prof_harvest());
  prof_step_begin();
  for (int i = 0; i < n_cats; ++i) { counts[i] = 0; ms[i] = flops[i] = bytes[i] = 0.0; }
  for (int i = 0; i < PC_COUNT; ++i) { counts[i] = g_acc_n[i]; ms[i] = g_acc_ms[i]; flops[i] = g_acc_flops[i]; bytes[i] = g_acc_bytes[i]; }
  return 0;
}

const char* scv_profile_category_name(int32_t cat) {
  static const char* names[PC_COUNT] = {"linear_simt", "layernorm", "attention_self", "attention_cross", "sampler",
                                        "embed", "misc", "gemm_tcgen05", "event_pair_overhead"};
  return (cat >= 0 && cat < PC_COUNT) ? names[cat] : "";
}

int scv_op_linear(const float* x, int32_t ldx, const uint16_t* w_bf16, int32_t ldw, const float* bias,
                  const float* residual, int32_t ldr, float* y, int32_t ldy, int32_t M, int32_t N, int32_t K,
                  int32_t act, int32_t impl, void* stream) {
  LinearArgs a;
  a.x = x; a.ldx = ldx; a.bias = bias;
  if (impl == 2) a.wt = reinterpret_cast<const __nv_bfloat16*>(w_bf16);   // tiled layout (scv_op_pack_tiled)
  else { a.w = reinterpret_cast<const __nv_bfloat16*>(w_bf16); a.ldw = ldw; }
  a.residual = residual; a.ldr = ldr; a.y = y; a.ldy = ldy; a.M = M; a.N = N; a.K = K; a.act = act;
  return launch_linear(a, impl, static_cast<cudaStream_t>(stream));
}

int64_t scv_op_split_tile_bytes(int32_t M, int32_t K) { return (int64_t)split_tile_bytes(M, K); }

int scv_op_split_rows(const float* x, int32_t ldx, const float* gamma, const float* beta, void* out_split, int32_t M,
                      int32_t N, int32_t normalize, void* stream) {
  return launch_layernorm_split(x, ldx, gamma, beta, out_split, M, N, normalize, nullptr, static_cast<cudaStream_t>(stream));
}

int scv_op_linear_split(const void* a_split, const uint16_t* w_tiled, const float* bias, const float* residual,
                        int32_t ldr, float* y, int32_t ldy, void* y_split, int32_t M, int32_t N, int32_t K, int32_t act,
                        void* stream) {
  LinearArgs a;
  a.a_split = a_split; a.wt = reinterpret_cast<const __nv_bfloat16*>(w_tiled); a.bias = bias; a.residual = residual;
  a.ldr = ldr; a.y = y; a.ldy = ldy; a.y_split = y_split; a.M = M; a.N = N; a.K = K; a.act = act;
  return launch_linear_tcgen05(a, static_cast<cudaStream_t>(stream));
}

int64_t scv_op_tiled_elems(int32_t N, int32_t K) { return (int64_t)tc_packed_elems(N, K); }

int scv_op_pack_tiled(const float* src, uint16_t* dst, int32_t N, int32_t K, void* stream) {
  return launch_pack_tiled(src, reinterpret_cast<__nv_bfloat16*>(dst), N, K, static_cast<cudaStream_t>(stream));
}

int scv_op_pack_bf16(const float* src, uint16_t* dst, int32_t rows, int32_t cols, int32_t ld_dst, void* stream) {
  SCV_REQUIRE(ld_dst >= cols, "pack: ld_dst < cols");
  return launch_pack_bf16(src, reinterpret_cast<__nv_bfloat16*>(dst), rows, cols, ld_dst,
                          static_cast<cudaStream_t>(stream));
}

int scv_op_layernorm(const float* x, int32_t ldx, const float* gamma, const float* beta, float* y, int32_t ldy,
                     int32_t M, int32_t N, int32_t act, void* stream) {
  return launch_layernorm(x, ldx, gamma, beta, y, ldy, M, N, act, nullptr, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
