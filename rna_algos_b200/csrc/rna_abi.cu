// rna_abi.cu — C ABI (include/rna_algos_b200.h) over the CUDA kernels: handle, table packing,
// length bucketing, launches, host<->device copies.  No CPU fallback: every entry point that computes
// needs a CUDA device and reports RNA_ERR_NO_DEVICE / RNA_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <queue>
#include <string>
#include <vector>

#include "../../include/rna_algos_b200.h"
#include "dev_tables.h"
#include "durbin_kernel.cuh"
#include "fast_kernel.cuh"
#include "fold_kernel2.cuh"
#include "table_pack.h"
#include "twoloop_export.cuh"

using namespace rna;

// ---------------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct rna_handle {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;
  static const int kLanes = 4;                  // length buckets of one call run on up to 4 concurrent streams
  cudaStream_t aux[kLanes] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kLanes] = {nullptr, nullptr, nullptr, nullptr};
  std::string err;
  bool has_turner = false, has_contra = false, has_align = false;
  DevTurner* d_turner = nullptr;
  DevContra* d_contra = nullptr;
  DevAlign* d_align = nullptr;
  float *d_hp_ext = nullptr, *d_int11 = nullptr, *d_int12 = nullptr, *d_int22 = nullptr;
  DevBuf ws, counters, order, stream_ws;                           // kernel scratch
  DevBuf fast_ws, fast_bpp;                                        // FAST numeric mode: matrix slots, internal BPP
  int numeric_mode = RNA_NUMERIC_REF_EXACT;
  DevBuf b_bases, b_offsets, b_bppoff, b_gammas, b_logz, b_bpp, b_structs, b_ea, b_pairs, b_probs, b_proboff;
  RnaCallStats stats{};
  // cached answers of the CUDA runtime about kernel configurations (kern_prepare)
  struct OccEntry { const void* fn; int nt; size_t smem; int occ; };
  std::vector<OccEntry> occ_cache;
  std::vector<const void*> attr_set;
  int lsmem[2] = {-1, -1};                       // largest shared-memory-mode length per model
  cudaEvent_t ev_done = nullptr;                 // end of the last enqueued call: calls of one handle share its scratch
  bool ev_done_valid = false;
};

// Development switches (A/B timing, role timers) are read from the environment only in builds made with
// -DRNA_DEV_SWITCHES (make EXTRA=-DRNA_DEV_SWITCHES); the product build has none.
#ifdef RNA_DEV_SWITCHES
static inline const char* dev_env(const char* name) { return getenv(name); }
#else
static inline const char* dev_env(const char*) { return nullptr; }
#endif

#define CU(h, call)                                                                                  \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                                \
      return RNA_ERR_CUDA;                                                                           \
    }                                                                                                \
  } while (0)

static int ensure(rna_handle* h, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap && b.p) return RNA_OK;
  if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
  size_t want = std::max<size_t>(bytes, 256);
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    h->err = std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e);
    cudaGetLastError();
    return RNA_ERR_NOMEM;
  }
  b.cap = want;
  return RNA_OK;
}
#define TRY(x) do { int rc__ = (x); if (rc__ != RNA_OK) return rc__; } while (0)

extern "C" const char* rna_version(void) { return "rna_algos_b200 0.1.0 (sm_100a)"; }
extern "C" size_t rna_sizeof_turner_tables(void) { return sizeof(RnaTurnerTables); }
extern "C" size_t rna_sizeof_contra_tables(void) { return sizeof(RnaContraTables); }
extern "C" size_t rna_sizeof_align_tables(void) { return sizeof(RnaAlignTables); }

extern "C" int rna_create(int device, rna_handle** out) {
  if (!out) return RNA_ERR_BAD_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return RNA_ERR_NO_DEVICE; }
  if (device < 0 || device >= n) return RNA_ERR_BAD_ARG;
  rna_handle* h = new rna_handle();
  h->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete h; return RNA_ERR_NO_DEVICE; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return RNA_ERR_NO_DEVICE; }
  h->sm_count = prop.multiProcessorCount;
  h->smem_optin = prop.sharedMemPerBlockOptin;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return RNA_ERR_CUDA; }
  bool ok = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming) == cudaSuccess;
  for (int x = 0; x < rna_handle::kLanes; x++) {
    ok = ok && cudaStreamCreateWithFlags(&h->aux[x], cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_join[x], cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) { delete h; return RNA_ERR_CUDA; }
  *out = h;
  return RNA_OK;
}

static void free_buf(DevBuf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }

extern "C" int rna_destroy(rna_handle* h) {
  if (!h) return RNA_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  DevBuf* bufs[] = {&h->ws, &h->counters, &h->order, &h->stream_ws, &h->fast_ws, &h->fast_bpp, &h->b_bases, &h->b_offsets, &h->b_bppoff, &h->b_gammas,
                    &h->b_logz, &h->b_bpp, &h->b_structs, &h->b_ea, &h->b_pairs, &h->b_probs, &h->b_proboff};
  for (DevBuf* b : bufs) free_buf(*b);
  cudaFree(h->d_turner); cudaFree(h->d_contra); cudaFree(h->d_align);
  cudaFree(h->d_hp_ext); cudaFree(h->d_int11); cudaFree(h->d_int12); cudaFree(h->d_int22);
  for (int x = 0; x < rna_handle::kLanes; x++) { cudaStreamSynchronize(h->aux[x]); cudaStreamDestroy(h->aux[x]); cudaEventDestroy(h->ev_join[x]); }
  cudaEventDestroy(h->ev_fork);
  cudaEventDestroy(h->ev_done);
  cudaStreamDestroy(h->stream);
  delete h;
  return RNA_OK;
}

extern "C" const char* rna_last_error(const rna_handle* h) { return h ? h->err.c_str() : "null handle"; }
extern "C" int rna_device(const rna_handle* h) { return h ? h->device : -1; }
extern "C" int rna_set_numeric_mode(rna_handle* h, int mode) {
  if (!h || mode < RNA_NUMERIC_REF_EXACT || mode > RNA_NUMERIC_FAST_F64) return RNA_ERR_BAD_ARG;
  h->numeric_mode = mode;
  return RNA_OK;
}
extern "C" int rna_get_numeric_mode(const rna_handle* h) { return h ? h->numeric_mode : -1; }
extern "C" int rna_get_stats(const rna_handle* h, RnaCallStats* out) {
  if (!h || !out) return RNA_ERR_BAD_ARG;
  *out = h->stats;
  return RNA_OK;
}

// ---------------------------------------------------------------------------------------------------
// table packing (host, plain IEEE f32 — compiled with -ffp-contract=off)
// ---------------------------------------------------------------------------------------------------
// A table upload must not overtake a call that is still reading the old tables (the *_dev entry points return
// before their kernels have run, on streams that a blocking cudaMemcpy does not order against).
static int tables_quiesce(rna_handle* h) {
  CU(h, cudaSetDevice(h->device));
  if (h->ev_done_valid) CU(h, cudaEventSynchronize(h->ev_done));
  return RNA_OK;
}

extern "C" void rna_contra_tables_accumulate(RnaContraTables* t) {
  // FoldScoreSets::accumulate, reference src/mccaskill_algo.rs:60-86
  struct { const float* src; float* dst; int n; } jobs[] = {
      {t->hairpin_scores_len, t->hairpin_scores_len_cumulative, RNA_CONTRA_MAX_LOOP_LEN + 1},
      {t->bulge_scores_len, t->bulge_scores_len_cumulative, RNA_CONTRA_MAX_LOOP_LEN},
      {t->interior_scores_len, t->interior_scores_len_cumulative, RNA_CONTRA_MAX_LOOP_LEN - 1},
      {t->interior_scores_symmetric, t->interior_scores_symmetric_cumulative, RNA_CONTRA_MAX_INTERIOR_SYMMETRIC},
      {t->interior_scores_asymmetric, t->interior_scores_asymmetric_cumulative, RNA_CONTRA_MAX_INTERIOR_ASYMMETRIC}};
  for (auto& j : jobs) {
    volatile float sum = 0.f;
    for (int i = 0; i < j.n; i++) { sum = sum + j.src[i]; j.dst[i] = sum; }
  }
}

extern "C" void rna_align_tables_contralign_v201(RnaAlignTables* t) {
  // reference src/compiled_align_scores.rs:2-19
  static const float m[4][4] = {{0.5256508867f, -0.40906402f, -0.2502759109f, -0.3252306723f},
                                {-0.40906402f, 0.6665219366f, -0.3289391181f, -0.1326088918f},
                                {-0.2502759109f, -0.3289391181f, 0.6684676551f, -0.3565888168f},
                                {-0.3252306723f, -0.1326088918f, -0.3565888168f, 0.459052045f}};
  static const float ins[4] = {-0.002521927159f, -0.08313891561f, -0.07443970653f, -0.01290054598f};
  memcpy(t->match_scores, m, sizeof m);
  memcpy(t->insert_scores, ins, sizeof ins);
  t->init_match_score = 0.3959924457f;
  t->init_insert_score = -0.3488104904f;
  t->match2match_score = 2.50575671f;
  t->match2insert_score = 0.1970448791f;
  t->insert_extend_score = 1.014026583f;
  t->insert_switch_score = -7.346968782f;
}

extern "C" int rna_set_turner_tables(rna_handle* h, const RnaTurnerTables* t) {
  if (!h || !t) return RNA_ERR_BAD_ARG;
  DevTurner d;
  std::vector<float> hp;
  TRY(pack_turner(t, &d, &hp, &h->err));
  TRY(tables_quiesce(h));
  if (!h->d_turner) {
    CU(h, cudaMalloc(&h->d_turner, sizeof(DevTurner)));
    CU(h, cudaMalloc(&h->d_hp_ext, sizeof(float) * RNA_HAIRPIN_EXT_LEN));
    CU(h, cudaMalloc(&h->d_int11, sizeof t->interior_scores_1x1));
    CU(h, cudaMalloc(&h->d_int12, sizeof t->interior_scores_1x2));
    CU(h, cudaMalloc(&h->d_int22, sizeof t->interior_scores_2x2));
  }
  d.hairpin_init_ext = h->d_hp_ext;
  d.int11 = h->d_int11;
  d.int12 = h->d_int12;
  d.int22 = h->d_int22;
  CU(h, cudaMemcpy(h->d_hp_ext, hp.data(), sizeof(float) * RNA_HAIRPIN_EXT_LEN, cudaMemcpyHostToDevice));
  CU(h, cudaMemcpy(h->d_int11, t->interior_scores_1x1, sizeof t->interior_scores_1x1, cudaMemcpyHostToDevice));
  CU(h, cudaMemcpy(h->d_int12, t->interior_scores_1x2, sizeof t->interior_scores_1x2, cudaMemcpyHostToDevice));
  CU(h, cudaMemcpy(h->d_int22, t->interior_scores_2x2, sizeof t->interior_scores_2x2, cudaMemcpyHostToDevice));
  CU(h, cudaMemcpy(h->d_turner, &d, sizeof d, cudaMemcpyHostToDevice));
  h->has_turner = true;
  return RNA_OK;
}

extern "C" int rna_set_contra_tables(rna_handle* h, const RnaContraTables* t) {
  if (!h || !t) return RNA_ERR_BAD_ARG;
  DevContra d;
  TRY(pack_contra(t, &d, &h->err));
  TRY(tables_quiesce(h));
  if (!h->d_contra) CU(h, cudaMalloc(&h->d_contra, sizeof(DevContra)));
  CU(h, cudaMemcpy(h->d_contra, &d, sizeof d, cudaMemcpyHostToDevice));
  h->has_contra = true;
  return RNA_OK;
}

extern "C" int rna_set_align_tables(rna_handle* h, const RnaAlignTables* t) {
  if (!h || !t) return RNA_ERR_BAD_ARG;
  DevAlign d;
  TRY(pack_align(t, &d, &h->err));
  TRY(tables_quiesce(h));
  if (!h->d_align) CU(h, cudaMalloc(&h->d_align, sizeof(DevAlign)));
  CU(h, cudaMemcpy(h->d_align, &d, sizeof d, cudaMemcpyHostToDevice));
  h->has_align = true;
  return RNA_OK;
}

// ---------------------------------------------------------------------------------------------------
// validation + partition (pure host)
// ---------------------------------------------------------------------------------------------------
extern "C" int rna_validate_bases(const uint8_t* bases, const uint32_t* offsets, uint32_t n_seqs) {
  if (!offsets || (n_seqs && !bases)) return RNA_ERR_BAD_ARG;
  for (uint32_t s = 0; s < n_seqs; s++) {
    if (offsets[s + 1] < offsets[s]) return RNA_ERR_BAD_ARG;
    const uint32_t L = offsets[s + 1] - offsets[s];
    if (L == 0) return RNA_ERR_EMPTY_SEQ;
    if (L > RNA_MAX_SEQ_LEN) return RNA_ERR_TOO_LONG;
  }
  const uint32_t tot = n_seqs ? offsets[n_seqs] : 0, beg = n_seqs ? offsets[0] : 0;
  for (uint32_t x = beg; x < tot; x++)
    if (bases[x] > 3) return RNA_ERR_INVALID_BASE;
  return RNA_OK;
}

extern "C" int rna_validate_fold_lengths(const uint32_t* offsets, uint32_t n_seqs) {
  if (!offsets) return RNA_ERR_BAD_ARG;
  for (uint32_t s = 0; s < n_seqs; s++)
    if (offsets[s + 1] >= offsets[s] && offsets[s + 1] - offsets[s] > (uint32_t)RNA_MAX_FOLD_LEN) return RNA_ERR_TOO_LONG;
  return RNA_OK;
}

extern "C" int rna_partition_lpt(const uint64_t* costs, uint32_t n_units, uint32_t n_parts, uint32_t* part_of) {
  if (!costs || !part_of || n_parts == 0) return RNA_ERR_BAD_ARG;
  std::vector<uint32_t> idx(n_units);
  for (uint32_t i = 0; i < n_units; i++) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return costs[a] > costs[b]; });
  typedef std::pair<uint64_t, uint32_t> LP;   // (load, part): least-loaded part first, ties -> lowest index
  std::priority_queue<LP, std::vector<LP>, std::greater<LP>> pq;
  for (uint32_t p = 0; p < n_parts; p++) pq.push(LP(0, p));
  for (uint32_t u : idx) {
    LP t = pq.top();
    pq.pop();
    part_of[u] = t.second;
    pq.push(LP(t.first + costs[u], t.second));
  }
  return RNA_OK;
}

// ---------------------------------------------------------------------------------------------------
// fold launcher
// ---------------------------------------------------------------------------------------------------
// Per-handle cache of what the CUDA runtime is asked about a kernel configuration (a batch call used to repeat these
// queries for every bucket of every call: cudaFuncSetAttribute + occupancy query cost ~10 us each).
static int kern_prepare(rna_handle* h, const void* fn, int nt, size_t smem, int* occ_out) {
  bool attr_done = false;
  for (const void* f : h->attr_set) attr_done = attr_done || f == fn;
  if (!attr_done) {
    cudaFuncAttributes fa;
    CU(h, cudaFuncGetAttributes(&fa, fn));   // (static shared memory comes out of the same opt-in budget)
    CU(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin - (int)std::max<size_t>(1024, fa.sharedSizeBytes)));
    h->attr_set.push_back(fn);
  }
  if (!occ_out) return RNA_OK;
  for (const rna_handle::OccEntry& e : h->occ_cache)
    if (e.fn == fn && e.nt == nt && e.smem == smem) { *occ_out = e.occ; return RNA_OK; }
  int occ = 1;
  CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, nt, smem));
  h->occ_cache.push_back(rna_handle::OccEntry{fn, nt, smem, occ});
  *occ_out = occ;
  return RNA_OK;
}

struct Bucket {
  int mode;
  int Lcap;
  uint32_t begin, end;   // range in the sorted order array
};

// sequence indices by length, longest first (LPT inside each launch's work queue); stable, O(n + max length)
// Sequences (all, or the ones listed in `subset`) in order of decreasing length (counting sort, stable).
static void order_by_length(const uint32_t* ho, uint32_t n, std::vector<uint32_t>& order, const std::vector<uint32_t>* subset = nullptr) {
  const uint32_t m = subset ? (uint32_t)subset->size() : n;
  auto id = [&](uint32_t x) { return subset ? (*subset)[x] : x; };
  uint32_t maxlen = 0;
  for (uint32_t x = 0; x < m; x++) maxlen = std::max(maxlen, ho[id(x) + 1] - ho[id(x)]);
  std::vector<uint32_t> start((size_t)maxlen + 2, 0);
  for (uint32_t x = 0; x < m; x++) start[maxlen - (ho[id(x) + 1] - ho[id(x)]) + 1]++;
  for (size_t x = 1; x < start.size(); x++) start[x] += start[x - 1];
  order.resize(m);
  for (uint32_t x = 0; x < m; x++) order[start[maxlen - (ho[id(x) + 1] - ho[id(x)])]++] = id(x);
}

// Every call of a handle uses the same scratch (workspace, work counters, order array): calls are stream-ordered
// behind one another whatever stream they were enqueued on.
static int call_begin(rna_handle* h, cudaStream_t st) {
  if (h->ev_done_valid) CU(h, cudaStreamWaitEvent(st, h->ev_done, 0));
  return RNA_OK;
}
static int call_end(rna_handle* h, cudaStream_t st) {
  CU(h, cudaEventRecord(h->ev_done, st));
  h->ev_done_valid = true;
  return RNA_OK;
}

// centroid_fold over packed BPP matrices (rna_centroid_batch): buckets by length, one CTA per sequence; sequences
// beyond 1024 nt one at a time on the whole grid
static int launch_centroid(rna_handle* h, const RnaFoldBatchDev* b, cudaStream_t st, const float* d_bpp_in,
                           const std::vector<uint32_t>* subset = nullptr) {
  const uint32_t* ho = b->h_offsets;
  std::vector<uint32_t> order;
  order_by_length(ho, b->n_seqs, order, subset);
  const uint32_t n = (uint32_t)order.size();   // sequences of this launch
  auto len_of = [&](uint32_t pos) { return (int)(ho[order[pos] + 1] - ho[order[pos]]); };
  const int gran = 8;
  int Lsmem = 0;
  for (int L = 16; L <= 1024; L += gran) { if (centroid_ws_floats(L) * 4 + 1024 <= h->smem_optin) Lsmem = L; else break; }
  // sequences beyond the shared-memory mode: one CTA each when there are many, else one at a time on the whole grid
  // (a lone 1024-nt sequence takes > 100 ms on one CTA, ~5 ms on the grid)
  uint32_t n_long = 0;
  while (n_long < n && len_of(n_long) > Lsmem) n_long++;
  const int Lcoop_min = (n_long <= 32) ? std::max(Lsmem + 1, 384) : 1025;
  std::vector<Bucket> buckets;
  for (uint32_t pos = 0; pos < n;) {
    const int L = len_of(pos);
    Bucket bk;
    bk.begin = pos;
    if (L >= Lcoop_min) {
      if (L > RNA_MAX_FOLD_LEN) { h->err = "sequence longer than RNA_MAX_FOLD_LEN (46340 nt)"; return RNA_ERR_TOO_LONG; }
      bk.mode = MODE_COOP; bk.Lcap = L; bk.end = pos + 1;
    } else if (L > Lsmem) {
      bk.mode = MODE_GLOBAL; bk.Lcap = L;
      uint32_t e = pos;
      while (e < n && len_of(e) > Lsmem) e++;
      bk.end = e;
    } else {
      bk.mode = MODE_SMEM;
      bk.Lcap = std::max(16, (L + gran - 1) / gran * gran);
      const int lo = bk.Lcap > 16 ? bk.Lcap - gran : 0;
      uint32_t e = pos;
      while (e < n && len_of(e) > lo) e++;
      bk.end = e;
    }
    buckets.push_back(bk);
    pos = bk.end;
  }
  TRY(ensure(h, h->order, sizeof(uint32_t) * (size_t)std::max<uint32_t>(n, 1)));
  TRY(ensure(h, h->counters, sizeof(int) * buckets.size()));
  CU(h, cudaMemcpyAsync(h->order.p, order.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemsetAsync(h->counters.p, 0, sizeof(int) * buckets.size(), st));
  h->stats.h2d_bytes += sizeof(uint32_t) * n;
  size_t ws_floats = 0;
  std::vector<int> grid_of(buckets.size(), 0);
  for (size_t k = 0; k < buckets.size(); k++) {
    const Bucket& bk = buckets[k];
    const size_t per = centroid_ws_floats(bk.Lcap) + 32;
    if (bk.mode == MODE_COOP) {
      const size_t Tc = (size_t)bk.Lcap * (bk.Lcap + 1) / 2;
      ws_floats = std::max(ws_floats, (size_t)std::max<uint32_t>(1, b->n_gammas) * (Tc + 2 * ((size_t)bk.Lcap + 2)) + 128);
    } else if (bk.mode == MODE_GLOBAL) {
      size_t freeb = 0, totb = 0;
      cudaMemGetInfo(&freeb, &totb);
      const size_t budget = (freeb + h->ws.cap) / 2;
      int g = (int)std::min<size_t>(bk.end - bk.begin, (size_t)h->sm_count * 2);
      while (g > 1 && (size_t)g * per * 4 > budget) g--;
      if ((size_t)g * per * 4 > budget) { h->err = "sequence too long for device memory"; return RNA_ERR_NOMEM; }
      grid_of[k] = g;
      ws_floats = std::max(ws_floats, (size_t)g * per);
    }
  }
  if (ws_floats) TRY(ensure(h, h->ws, ws_floats * 4));
  FoldArgs a;
  memset(&a, 0, sizeof a);
  a.offsets = b->d_offsets;
  a.n_seqs = b->n_seqs;
  a.total_len = b->total_len;
  a.gammas = b->d_gammas;
  a.n_gammas = b->n_gammas;
  a.bpp_offsets = b->d_bpp_offsets;
  a.out_structs = b->d_out_structs;
  a.out_ea = b->d_out_expect_acc;
  a.out_pairs = b->d_out_pairs;
  a.out_npairs = b->d_out_num_pairs;
  a.workspace = (float*)h->ws.p;
  for (size_t k = 0; k < buckets.size(); k++) {
    const Bucket& bk = buckets[k];
    a.order = (const uint32_t*)h->order.p + bk.begin;
    a.n_launch = bk.end - bk.begin;
    a.work_counter = (int*)h->counters.p + k;
    a.Lcap = bk.Lcap;
    a.ws_stride = centroid_ws_floats(bk.Lcap) + 32;
    if (bk.mode == MODE_SMEM) {
      const int nt = std::min(256, std::max(32, (bk.Lcap + 31) / 32 * 32));
      const size_t smem = centroid_ws_floats(bk.Lcap) * 4;
      int occ = 1;
      TRY(kern_prepare(h, (const void*)centroid_kernel<MODE_SMEM>, nt, smem, &occ));
      const int grid = (int)std::min<size_t>(a.n_launch, (size_t)std::max(1, occ) * h->sm_count);
      centroid_kernel<MODE_SMEM><<<grid, nt, smem, st>>>(a, d_bpp_in);
    } else if (bk.mode == MODE_GLOBAL) {
      const int nt = std::min(512, std::max(32, (bk.Lcap + 31) / 32 * 32));
      centroid_kernel<MODE_GLOBAL><<<grid_of[k], nt, 16, st>>>(a, d_bpp_in);
    } else {
      // whole grid: the chunked atomic-max fill has (thresholds x cell groups x chunks) warp tasks per diagonal
      const int nt = 256;
      int occ = 1;
      TRY(kern_prepare(h, (const void*)centroid_kernel<MODE_COOP>, nt, 16, &occ));
      const int grid = std::max(1, std::min(occ, 2)) * h->sm_count;
      void* params[] = {(void*)&a, (void*)&d_bpp_in};
      CU(h, cudaLaunchCooperativeKernel((void*)centroid_kernel<MODE_COOP>, dim3(grid), dim3(nt), params, 16, st));
    }
    CU(h, cudaGetLastError());
    h->stats.kernel_launches++;
  }
  return RNA_OK;
}

#ifdef RNA_DEV_SWITCHES
// RNA_FOLD_DBG=1 (development builds): where do the cycles of one sequence go, per role and pass
static void fold_dbg_report(int mode, int Lcap, const FoldArgs& a, long long* d_dbg, cudaStream_t st, int nt, int grid, size_t smem) {
    if (mode == MODE_COOP) {
      cudaStreamSynchronize(st);
      long long hd[32];
      cudaMemcpy(hd, d_dbg, sizeof hd, cudaMemcpyDeviceToHost);
      fprintf(stderr, "[RNA_FOLD_DBG] coop L=%d roles %d/%d/%d: setup+streams=%lld inside=%lld outside=%lld bpp+centroid=%lld cycles\n", Lcap,
              a.nXw, a.nYw, a.nZw, hd[0], hd[1], hd[2], hd[3]);
      fprintf(stderr, "[RNA_FOLD_DBG]   inside, first warp of a role (work, wait A, phase B): X %lld %lld %lld | Y %lld %lld %lld | Z %lld %lld %lld\n",
              hd[8], hd[9], hd[10], hd[12], hd[13], hd[14], hd[16], hd[17], hd[18]);
      fprintf(stderr, "[RNA_FOLD_DBG]   outside, first warp of a role (work, wait): X %lld %lld | Y %lld %lld\n", hd[20], hd[21], hd[24], hd[25]);
      {
        std::vector<long long> hs((size_t)(Lcap / 2 + 2) * 8);
        cudaMemcpy(hs.data(), d_dbg + 64, hs.size() * 8, cudaMemcpyDeviceToHost);
        long long sm[5] = {0, 0, 0, 0, 0}, crit = 0, nsA = 0, nsBar = 0, nsB = 0, prev_end = 0;
        for (size_t stp = 0; stp * 8 < hs.size(); stp++) {
          long long mx = 0;
          for (int r = 0; r < 5; r++) { sm[r] += hs[stp * 8 + r]; mx = std::max(mx, hs[stp * 8 + r]); }
          crit += mx;
          const long long ga = hs[stp * 8 + 5], gr = hs[stp * 8 + 6], ge = hs[stp * 8 + 7];
          if (ga && gr && ge) { if (prev_end) nsA += ga - prev_end; nsBar += gr - ga; nsB += ge - gr; prev_end = ge; }
        }
        fprintf(stderr, "[RNA_FOLD_DBG]   inside wall (globaltimer, last CTA): phase A %.3f ms, its barrier %.3f ms, phase B + barrier %.3f ms\n",
                nsA * 1e-6, nsBar * 1e-6, nsB * 1e-6);
        {
          std::vector<long long> ho((size_t)Lcap * 4);
          cudaMemcpy(ho.data(), d_dbg + (1 << 19), ho.size() * 8, cudaMemcpyDeviceToHost);
          long long ox = 0, oy = 0, oall = 0, ostream = 0;
          for (int dd = 0; dd < Lcap; dd++) {
            ox += ho[(size_t)dd * 4]; oy += ho[(size_t)dd * 4 + 1]; oall += std::max(ho[(size_t)dd * 4], ho[(size_t)dd * 4 + 1]);
            ostream += (ho[(size_t)dd * 4 + 2] & 0xffffff) << 4;
          }
          fprintf(stderr, "[RNA_FOLD_DBG]   outside, sum over steps of the slowest warp: X %lld (of which exterior + two-loop stream part %lld) Y %lld, all %lld\n", ox, ostream, oy, oall);
        }
        fprintf(stderr, "[RNA_FOLD_DBG]   inside phase A, sum over steps of the slowest warp: X %lld Y %lld Z-E %lld Z-M1 %lld Z-M %lld, all %lld\n",
                sm[0], sm[1], sm[2], sm[3], sm[4], crit);
      }
    } else {   // debug aid: where do the cycles of one sequence go, per role and pass
      cudaStreamSynchronize(st);
      std::vector<long long> hd(2048 * 16);
      cudaMemcpy(hd.data(), d_dbg, hd.size() * 8, cudaMemcpyDeviceToHost);
      for (int pass = 0; pass < 2; pass++) {
        long long role[3] = {0, 0, 0}, crit = 0, critrole[3] = {0, 0, 0};
        for (int t = 0; t < 1000; t++) {
          long long m[3] = {0, 0, 0};
          for (int wv = 0; wv < 16; wv++) {
            const int nxo = (pass == 1 && a.nXw_out > 0) ? a.nXw_out : a.nXw;   // (the outside pass may split differently: X | Y)
            const int r = pass == 1 ? (wv < nxo ? 0 : 1) : (wv < a.nXw ? 0 : wv < a.nXw + a.nYw ? 1 : 2);
            m[r] = std::max(m[r], hd[(size_t)(pass * 1024 + t) * 16 + wv]);
          }
          const long long mx = std::max(m[0], std::max(m[1], m[2]));
          crit += mx;
          for (int r = 0; r < 3; r++) { role[r] += m[r]; if (m[r] == mx && mx) critrole[r] += mx; }
        }
        fprintf(stderr, "[RNA_FOLD_DBG] bucket Lcap=%d pass=%s roles X/Y/Z warps %d/%d/%d: sum of per-step max cycles X=%lld Y=%lld Z=%lld, "
                "critical path=%lld (X %lld, Y %lld, Z %lld)\n", Lcap, pass ? "outside" : "inside", a.nXw, a.nYw, a.nZw,
                role[0], role[1], role[2], crit, critrole[0], critrole[1], critrole[2]);
      }
      for (int t = 1024; t < 1024; t++) {
        long long my = 0, mx = 0;
        for (int wv = 0; wv < 16; wv++) { (wv < a.nXw ? mx : my) = std::max(wv < a.nXw ? mx : my, hd[(size_t)t * 16 + wv]); }
        if (my > 200000) fprintf(stderr, "[RNA_FOLD_DBG]   outside d=%d: X=%lld Y=%lld\n", t - 1024, mx, my);
      }
      fprintf(stderr, "[RNA_FOLD_DBG]   setup=%lld count=%lld scan=%lld fill=%lld cycles, terms=%lld\n", hd[2047 * 16], hd[2047 * 16 + 1],
              hd[2047 * 16 + 2], hd[2047 * 16 + 3], hd[2047 * 16 + 4]);
      fprintf(stderr, "[RNA_FOLD_DBG]   L=%lld: output+centroid=%lld cycles, whole sequence=%lld cycles; launch: %d threads, grid %d, smem %zu\n",
              hd[2047 * 16 + 7], hd[2047 * 16 + 5], hd[2047 * 16 + 6], nt, grid, smem);
      if (!hd[2047 * 16 + 3]) fprintf(stderr, "[RNA_FOLD_DBG]   (streams did not fit: scored on the fly)\n");
      cudaMemset(d_dbg, 0, 2048 * 16 * 8);
    }
}
#endif

// fastnum: the FAST_F32 build of the batch kernel (fold_fastnum.cu); its caller passes sequences <= 1024 nt only
extern "C" const void* rna_fastnum_fold_kernel(int contra, int hbm_mode);
extern "C" unsigned long rna_fastnum_sizeof_fold_args();
template <bool CONTRA>
static int launch_fold_model(rna_handle* h, const RnaFoldBatchDev* b, cudaStream_t st, const std::vector<uint32_t>* subset = nullptr,
                             bool fastnum = false) {
  const bool want_sums = b->d_out_sums != nullptr || b->inside_only;   // (the SUMS build of the batch kernel)
  const uint32_t* ho = b->h_offsets;
  std::vector<uint32_t> order;
  order_by_length(ho, b->n_seqs, order, subset);
  const uint32_t n = (uint32_t)order.size();   // sequences of this launch
  if (fastnum && rna_fastnum_sizeof_fold_args() != sizeof(FoldArgs)) { h->err = "FoldArgs image mismatch"; return RNA_ERR_CUDA; }
  auto len_of = [&](uint32_t pos) { return (int)(ho[order[pos] + 1] - ho[order[pos]]); };
  const size_t smem_cap = h->smem_optin;
  const int gran = 4;                                     // bucket width in nt
  auto smem_need = [&](int Lcap) -> size_t { return fold2_fixed_bytes<CONTRA>(Lcap) + fold2_seq_bytes(Lcap, 1, 2); };
  // largest L whose shared-memory working set fits (the u8 cell lists cap it at 252)
  int& Lsmem = h->lsmem[CONTRA ? 1 : 0];
  if (Lsmem < 0) {
    Lsmem = 0;
    for (int L = 16; L <= 252; L += gran) { if (smem_need(L) + 1024 <= smem_cap) Lsmem = L; else break; }
  }
  // Sequences beyond 1024 nt get the whole grid, one at a time.  So do the few sequences that are too long for the
  // shared-memory mode when there are not enough of them to occupy the GPU one CTA each.
  int Lcoop_min = 1025;
  if (fastnum) Lcoop_min = 0x7fffffff;   // (long sequences of the FAST mode run on fast_fold_kernel's cooperative grid)
  else {
    uint32_t n_long = 0;
    while (n_long < n && len_of(n_long) > Lsmem) n_long++;
    // Cost model from B200 measurements (CONTRAfold; Turner is alike).  A cooperative run alone takes ~140 ms x
    // (L/1024)^1.7 up to 1024 nt (29 / 78 / 139 ms at 400 / 700 / 1000 nt) and x (L/1024)^2 beyond (300 / 487 ms at
    // 1500 / 2048); up to four run side by side (coop_lanes below) at ~0.8 of the speed-up their share of the SMs
    // allows (4 x 700 nt: 77 ms; 8 x 1024 nt: 442 ms).  A wave of up to 2 x SM-count one-CTA sequences takes ~2.3 s x
    // (L/1024)^3 when full and ~0.55 of that when nearly empty.  Sequences are sorted longest first: give the longest c
    // of the long ones to the cooperative kernel, c chosen to minimise the sum.
    auto coop_ms = [](int L) { const double r = L / 1024.0; return 140.0 * (r <= 1.0 ? std::pow(r, 1.7) : r * r); };
    auto coop_conc = [&](uint32_t c) {
      if (c < 2) return 1.0;
      const int Lmax = len_of(0), full = (Lmax + 31) / 32;
      const double sms_needed = std::max(1, ((3 * Lmax / 4 + 31) / 32 + 6 * full + (CONTRA ? 4 * full : 0) + 7) / 8);
      const int lanes = std::max(1, std::min<int>({(int)rna_handle::kLanes, (int)c, (int)(h->sm_count / (0.85 * sms_needed))}));
      return std::max(1.0, 0.8 * std::min<double>(lanes, h->sm_count / sms_needed));
    };
    auto wave_ms = [&](int Lmax, uint32_t cnt) {
      if (cnt == 0) return 0.0;
      const double r = Lmax / 1024.0, per_wave = 2.0 * h->sm_count;
      const double waves = cnt / per_wave;
      return 2300.0 * r * r * r * std::max(waves, std::min(1.0, 0.55 + 0.65 * waves));
    };
    uint32_t forced = 0;
    while (forced < n_long && len_of(forced) > 1024) forced++;
    double best = 1e300, prefix = 0;
    uint32_t best_c = forced;
    for (uint32_t c = 0; c <= std::min<uint32_t>(n_long, forced + 64); c++) {
      if (c >= forced) {
        const double cost = prefix / coop_conc(c) + (c < n_long ? wave_ms(len_of(c), n_long - c) : 0.0);
        if (cost < best) { best = cost; best_c = c; }
      }
      if (c < n_long) prefix += coop_ms(len_of(c));
    }
    if (best_c > forced) Lcoop_min = std::max(std::max(Lsmem + 1, 384), len_of(best_c - 1));
  }
  std::vector<Bucket> buckets;
  for (uint32_t pos = 0; pos < n;) {
    const int L = len_of(pos);
    Bucket bk;
    bk.begin = pos;
    if (L >= Lcoop_min) {
      // the cooperative kernel indexes its triangular matrices with 32-bit offsets (d * L must not overflow)
      if (L > RNA_MAX_FOLD_LEN) { h->err = "sequence longer than RNA_MAX_FOLD_LEN (46340 nt)"; return RNA_ERR_TOO_LONG; }
      bk.mode = MODE_COOP; bk.Lcap = L; bk.end = pos + 1;
    } else if (L > Lsmem) {
      // HBM-resident one-CTA mode, in length classes of 96 nt: the role split and the slot size follow the class
      bk.mode = MODE_GLOBAL; bk.Lcap = L;
      static const int cls = dev_env("RNA_GLOBAL_CLASS") ? atoi(dev_env("RNA_GLOBAL_CLASS")) : 96;
      const int lo = std::max(Lsmem, L - cls);
      uint32_t e = pos;
      while (e < n && len_of(e) > lo) e++;
      bk.end = e;
    } else {
      bk.mode = MODE_SMEM;
      bk.Lcap = std::max(16, (L + gran - 1) / gran * gran);
      const int lo = bk.Lcap > 16 ? bk.Lcap - gran : 0;   // bucket = lengths in (Lcap-gran, Lcap]; the smallest takes 1..16
      uint32_t e = pos;
      while (e < n && len_of(e) > lo) e++;
      bk.end = e;
    }
    buckets.push_back(bk);
    pos = bk.end;
  }
  TRY(ensure(h, h->order, sizeof(uint32_t) * (size_t)std::max<uint32_t>(n, 1)));
  TRY(ensure(h, h->counters, sizeof(int) * buckets.size()));
  CU(h, cudaMemcpyAsync(h->order.p, order.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemsetAsync(h->counters.p, 0, sizeof(int) * buckets.size(), st));
  h->stats.h2d_bytes += sizeof(uint32_t) * n;

  // workspace for the GLOBAL / COOP buckets
  size_t ws_floats = 0;
  std::vector<int> grid_of(buckets.size(), 0);
  std::vector<size_t> stride_of(buckets.size(), 0);
  size_t freeb = 0, totb = 0;
  bool have_meminfo = false;
  auto meminfo = [&]() { if (!have_meminfo) { cudaMemGetInfo(&freeb, &totb); have_meminfo = true; } };
  for (size_t k = 0; k < buckets.size(); k++) {
    const Bucket& bk = buckets[k];
    if (bk.mode == MODE_SMEM) continue;   // (sized below, once the grid is known)
    const size_t per = fold2_seq_bytes(bk.Lcap, 2, 5) / 4 + 64 + 32;
    stride_of[k] = per;
    if (bk.mode == MODE_COOP) {
      // + row-major sums_1ormore and the multibranch closing-score table of the cooperative kernel's outside pass
      // + one max-plus matrix and traceback stack per centroid threshold (all thresholds are filled in one sweep)
      const size_t Tc = (size_t)bk.Lcap * (bk.Lcap + 1) / 2;
      const size_t cent = (size_t)std::max<uint32_t>(1, b->n_gammas) * (Tc + 2 * ((size_t)bk.Lcap + 2)) + 64;
      ws_floats = std::max(ws_floats, per + 2 * Tc + 64 + cent);
    } else {
      meminfo();
      const size_t budget = (freeb + h->ws.cap) / 2;
      static const int gocc = dev_env("RNA_GLOBAL_OCC") ? atoi(dev_env("RNA_GLOBAL_OCC")) : 2;   // CTAs per SM of the HBM-resident mode
      int g = (int)std::min<size_t>(bk.end - bk.begin, (size_t)h->sm_count * gocc);
      while (g > 1 && (size_t)g * per * 4 > budget) g--;
      if ((size_t)g * per * 4 > budget) { h->err = "sequence too long for device memory"; return RNA_ERR_NOMEM; }
      grid_of[k] = g;
      ws_floats = std::max(ws_floats, (size_t)g * per);
    }
  }
  // shared-memory buckets: grid size and the per-CTA slots of the two-loop term streams
  std::vector<size_t> stream_stride_of(buckets.size(), 0);
  std::vector<uint32_t> tcap_of(buckets.size(), 0);
  static const bool no_streams = dev_env("RNA_FOLD_NOSTREAMS") != nullptr;   // A/B switches (RNA_DEV_SWITCHES builds)
  static const bool no_helpers = dev_env("RNA_FOLD_NOHELPERS") != nullptr;
  std::vector<int> nt_of(buckets.size(), 0);
  std::vector<Roles> ro_of(buckets.size());
  size_t stream_bytes = 0;
  // Several long sequences: the cooperative kernel keeps one chain per warp, so a sequence of L nt occupies
  // coop_warps(L) warps = that many / 8 SMs' worth of CTAs and the rest of the GPU idles (1024 nt: 43 of 148 SMs).
  // Up to kLanes of them run side by side, each on its own cooperative grid of sm_count / c CTAs (one CTA per SM:
  // the kernel takes the whole register file of an SM), sized for the longest one.
  auto coop_warps = [](int L) { const int full = (L + 31) / 32; return (3 * L / 4 + 31) / 32 + 6 * full + (CONTRA ? 4 * full : 0); };
  uint32_t n_coop = 0;
  while (n_coop < buckets.size() && buckets[n_coop].mode == MODE_COOP) n_coop++;   // (sorted longest first: they lead)
  static const int force_coop_lanes = dev_env("RNA_COOP_LANES") ? atoi(dev_env("RNA_COOP_LANES")) : 0;
  int coop_lanes = 1;
  if (n_coop > 1) {
    const int sms_needed = std::max(1, (coop_warps(buckets[0].Lcap) + 7) / 8);
    // (measured: a grid 15 % short of its lanes costs nothing — 8 x 1024 nt on 4 grids, 4 x 2048 nt on 2 — a grid 40 % short does)
    coop_lanes = std::max(1, std::min<int>({(int)rna_handle::kLanes, (int)n_coop, (int)(h->sm_count / (0.85 * sms_needed))}));
    if (force_coop_lanes > 0) coop_lanes = std::min<int>({force_coop_lanes, (int)rna_handle::kLanes, (int)n_coop});
  }
  for (size_t k = 0; k < buckets.size(); k++) {
    const Bucket& bk = buckets[k];
    if (bk.mode == MODE_COOP && !no_streams) {
      // one sequence on the whole GPU: a single stream slot, as large as memory comfortably allows (else the kernel
      // scores on the fly)
      const size_t Tc = (size_t)bk.Lcap * (bk.Lcap + 1) / 2;
      meminfo();
      size_t cap = std::min<size_t>(96 * Tc, (size_t)0xfffffff0u);
      const size_t budget = (freeb + h->stream_ws.cap) / 3 / (size_t)coop_lanes;
      if (cap * 16 > budget) cap = budget / 16;
      tcap_of[k] = (uint32_t)cap;
      stream_stride_of[k] = fold2_stream_bytes(bk.Lcap, tcap_of[k]);
      stream_bytes = std::max(stream_bytes, stream_stride_of[k]);
    }
    if (bk.mode != MODE_SMEM) continue;
    // CTAs per SM: bounded by shared memory (C, log P and the small per-sequence tables) and by 1024 threads per SM
    // at 64 registers.  The chains are latency-bound, so residency is what fills the SM: take all the CTAs that
    // fit (<= 4, measured best) and give each 1024 / CTAs threads: roles first, the rest help in the all-thread phases.
    const size_t smem = smem_need(bk.Lcap);
    static const int occ_cap = dev_env("RNA_FOLD_OCC") ? atoi(dev_env("RNA_FOLD_OCC")) : 4;   // measured best on tRNA-length batches
    int occ_s = (int)std::min<size_t>((size_t)std::max(1, occ_cap), (size_t)233472 / (smem + 1024 + 64));
    occ_s = std::max(1, occ_s);
    occ_s = std::min(occ_s, std::max(1, 32 / fold2_min_warps(bk.Lcap, CONTRA)));
    int warps = std::min(16, 32 / occ_s);
    if (no_helpers) { const Roles fr = fold2_roles(bk.Lcap, CONTRA, 16); warps = std::min(warps, fr.nX + fr.nY + fr.nZ); }
    ro_of[k] = fold2_roles(bk.Lcap, CONTRA, warps);
    const int nt = 32 * warps;
    int occ = 1;
    if (fastnum) TRY(kern_prepare(h, rna_fastnum_fold_kernel(CONTRA, 0), nt, smem, &occ));
    else if (want_sums) TRY(kern_prepare(h, (const void*)fold_kernel2<CONTRA, MODE_SMEM, true>, nt, smem, &occ));
    else TRY(kern_prepare(h, (const void*)fold_kernel2<CONTRA, MODE_SMEM>, nt, smem, &occ));
    occ = std::max(1, std::min(occ, occ_cap));
    nt_of[k] = nt;
    grid_of[k] = (int)std::min<size_t>(bk.end - bk.begin, (size_t)occ * h->sm_count);
    stride_of[k] = 4 * ((size_t)bk.Lcap * (bk.Lcap + 1) / 2) + 32;   // R X E M1 of one CTA
    ws_floats = std::max(ws_floats, (size_t)grid_of[k] * stride_of[k]);
    if (!no_streams) {
      // room for 128 stream elements per cell (random sequences need ~50 incl. padding); sequences with longer
      // lists score on the fly
      tcap_of[k] = (uint32_t)(128 * ((size_t)bk.Lcap * (bk.Lcap + 1) / 2));
      stream_stride_of[k] = fold2_stream_bytes(bk.Lcap, tcap_of[k]);
      stream_bytes = std::max(stream_bytes, (size_t)grid_of[k] * stream_stride_of[k]);
    }
  }
  // Buckets run concurrently on up to kLanes streams (bucket k on lane k % kLanes; small buckets would otherwise
  // leave most SMs idle, and the tails of big ones overlap: the next bucket's CTAs take over SMs as the last CTAs of
  // the previous one drain).  Kernels of one lane are stream-ordered, so the scratch is one region per lane, each
  // sized for the largest bucket.
  static const bool dbg_roles = dev_env("RNA_FOLD_DBG") != nullptr;
  static const bool serial = dev_env("RNA_FOLD_SERIAL") != nullptr || dbg_roles;
  size_t batch_buckets = 0;
  for (size_t k = 0; k < buckets.size(); k++)
    if (buckets[k].mode != MODE_COOP) batch_buckets++;
  static const int force_lanes = dev_env("RNA_FOLD_LANES") ? atoi(dev_env("RNA_FOLD_LANES")) : 0;
  if (serial) coop_lanes = 1;
  const int nlanes = serial ? 1
                     : std::max(coop_lanes, (int)std::max<size_t>(1, std::min<size_t>(force_lanes > 0 ? (size_t)std::min(force_lanes, (int)rna_handle::kLanes)
                                                                                               : (size_t)rna_handle::kLanes, batch_buckets)));
  ws_floats = (ws_floats + 63) / 64 * 64;
  stream_bytes = (stream_bytes + 255) / 256 * 256;
  if (stream_bytes) TRY(ensure(h, h->stream_ws, stream_bytes * nlanes));
  if (ws_floats) TRY(ensure(h, h->ws, ws_floats * 4 * nlanes));

  FoldArgs a;
  memset(&a, 0, sizeof a);
  a.bases = b->d_bases;
  a.offsets = b->d_offsets;
  a.n_seqs = b->n_seqs;
  a.total_len = b->total_len;
  a.allows_short = b->allows_short_hairpins;
  a.tables = CONTRA ? (const void*)h->d_contra : (const void*)h->d_turner;
  a.gammas = b->d_gammas;
  a.n_gammas = b->n_gammas;
  a.out_logz = b->d_out_logz;
  a.out_bpp = b->d_out_bpp;
  a.bpp_offsets = b->d_bpp_offsets;
  a.out_structs = b->d_out_structs;
  a.out_ea = b->d_out_expect_acc;
  a.out_pairs = b->d_out_pairs;
  a.out_npairs = b->d_out_num_pairs;
  a.sums = b->d_out_sums;
  a.sums_offsets = reinterpret_cast<const unsigned long long*>(b->d_sums_offsets);
  a.inside_only = b->inside_only;
  a.workspace = (float*)h->ws.p;
  if (nlanes > 1) {
    CU(h, cudaEventRecord(h->ev_fork, st));
    for (int x = 0; x < nlanes; x++) CU(h, cudaStreamWaitEvent(h->aux[x], h->ev_fork, 0));
  }
  cudaStream_t st_main = st;

  long long* d_dbg = nullptr;
  if (dbg_roles) { cudaMalloc(&d_dbg, (size_t)(1 << 20) * 8); cudaMemset(d_dbg, 0, (size_t)(1 << 20) * 8); a.dbg = d_dbg; }
  for (size_t k = 0; k < buckets.size(); k++) {
    const Bucket& bk = buckets[k];
    // cooperative grids first (the long sequences lead the order), coop_lanes of them side by side; then a join, so
    // that the one-CTA buckets never hold SMs a cooperative grid is waiting for
    const int lane = nlanes <= 1 ? 0 : (bk.mode == MODE_COOP ? (int)(k % coop_lanes) : (int)(k % nlanes));
    if (nlanes > 1 && k == n_coop && n_coop > 0) {
      for (int x = 0; x < nlanes; x++) { CU(h, cudaEventRecord(h->ev_join[x], h->aux[x])); CU(h, cudaStreamWaitEvent(st_main, h->ev_join[x], 0)); }
      CU(h, cudaEventRecord(h->ev_fork, st_main));
      for (int x = 0; x < nlanes; x++) CU(h, cudaStreamWaitEvent(h->aux[x], h->ev_fork, 0));
    }
    st = nlanes > 1 ? h->aux[lane] : st_main;
    a.workspace = (float*)h->ws.p + (size_t)lane * ws_floats;
    a.order = (const uint32_t*)h->order.p + bk.begin;
    a.n_launch = bk.end - bk.begin;
    a.work_counter = (int*)h->counters.p + k;
    a.Lcap = bk.Lcap;
    a.ws_stride = stride_of[k];
    if (bk.mode != MODE_COOP) {
      Roles ro = (bk.mode == MODE_SMEM) ? ro_of[k] : fold2_roles_global(bk.Lcap, CONTRA, 16);
      if (bk.mode != MODE_SMEM && dev_env("RNA_FOLD_ROLES")) {   // A/B: "x,y,z" warps of the HBM-resident one-CTA mode
        int x = 0, y = 0, z = 0;
        if (sscanf(dev_env("RNA_FOLD_ROLES"), "%d,%d,%d", &x, &y, &z) == 3 && x > 0 && z > 0 && x + y + z <= 16 && (y > 0) == CONTRA) { ro.nX = x; ro.nY = y; ro.nZ = z; }
      }
      a.nXw = ro.nX; a.nYw = ro.nY; a.nZw = ro.nZ;
      a.nXw_out = 0;
      if (bk.mode != MODE_SMEM) {
        a.nXw_out = fold2_xwarps_outside_global(bk.Lcap, ro.nX + ro.nY + ro.nZ);
        if (dev_env("RNA_FOLD_XOUT")) a.nXw_out = std::max(1, std::min(atoi(dev_env("RNA_FOLD_XOUT")), ro.nX + ro.nY + ro.nZ - 1));
      }
      const int nt = (bk.mode == MODE_SMEM) ? nt_of[k] : 32 * (ro.nX + ro.nY + ro.nZ);
      a.stream_ws = nullptr;
      if (bk.mode == MODE_SMEM) {
        const size_t smem = smem_need(bk.Lcap);
        if (stream_stride_of[k]) {
          a.tcap = tcap_of[k];
          a.stream_stride = stream_stride_of[k];
          a.stream_ws = (unsigned char*)h->stream_ws.p + (size_t)lane * stream_bytes;
        }
        void* params[] = {(void*)&a};
        if (fastnum) CU(h, cudaLaunchKernel(rna_fastnum_fold_kernel(CONTRA, 0), dim3(grid_of[k]), dim3(nt), params, smem, st));
        else if (want_sums) fold_kernel2<CONTRA, MODE_SMEM, true><<<grid_of[k], nt, smem, st>>>(a);
        else fold_kernel2<CONTRA, MODE_SMEM><<<grid_of[k], nt, smem, st>>>(a);
      } else {
        static const bool no_ring = dev_env("RNA_NO_ZRING") != nullptr;
        a.ring_bytes = no_ring ? 0u : (unsigned)((size_t)std::max(std::max(ro.nX, a.nXw_out), std::max(ro.nY + ro.nZ, ro.nX + ro.nY + ro.nZ - a.nXw_out)) * 32 * RNA_Z_RING * sizeof(float4));   // a ring column per lane of the widest role
        const size_t smem = fold2_fixed_bytes<CONTRA>(bk.Lcap) + a.ring_bytes;
        if (fastnum) {
          void* params[] = {(void*)&a};
          TRY(kern_prepare(h, rna_fastnum_fold_kernel(CONTRA, 1), nt, smem, nullptr));
          CU(h, cudaLaunchKernel(rna_fastnum_fold_kernel(CONTRA, 1), dim3(grid_of[k]), dim3(nt), params, smem, st));
        } else if (want_sums) {
          TRY(kern_prepare(h, (const void*)fold_kernel2<CONTRA, MODE_GLOBAL, true>, nt, smem, nullptr));
          fold_kernel2<CONTRA, MODE_GLOBAL, true><<<grid_of[k], nt, smem, st>>>(a);
        } else {
          TRY(kern_prepare(h, (const void*)fold_kernel2<CONTRA, MODE_GLOBAL>, nt, smem, nullptr));
          fold_kernel2<CONTRA, MODE_GLOBAL><<<grid_of[k], nt, smem, st>>>(a);
        }
      }
    } else {
      // long sequence: cooperative grid, roles spread over the SMs
      const int nt = 256;
      const size_t smem = fold2_fixed_bytes<CONTRA>(bk.Lcap) + RNA_ML_RING_FLOATS * 4;
      static const bool no_split = dev_env("RNA_COOP_NOSPLIT") != nullptr;
      a.no_ml_split = no_split ? 1 : 0;
      int occ = 1;
      TRY(kern_prepare(h, (const void*)fold_kernel2_coop<CONTRA>, nt, smem, &occ));
      const int grid = std::max(1, std::max(1, std::min(occ, 2)) * h->sm_count / coop_lanes);
      const int W = grid * (nt / 32);
      // inside pair steps: X = closable cells of two diagonals, Y = all cells of two diagonals, Z = three chains per
      // cell of two diagonals (one lane each)
      const int full = (bk.Lcap + 31) / 32;
      int nX = (3 * bk.Lcap / 4 + 31) / 32, nZ = 6 * full, nY = CONTRA ? 4 * full : 0;
      while (nX + nY + nZ > W) { if (nZ > 1) nZ--; if (nY > 1) nY--; if (nX > 1 && nX + nY + nZ > W) nX--; if (nX + nY + nZ <= 3) break; }
      a.nXw = nX; a.nYw = nY; a.nZw = nZ;
      a.stream_ws = nullptr;
      if (stream_stride_of[k]) {
        a.tcap = tcap_of[k];
        a.stream_stride = stream_stride_of[k];
        a.stream_ws = (unsigned char*)h->stream_ws.p + (size_t)lane * stream_bytes;
      }
      void* params[] = {(void*)&a};
      CU(h, cudaLaunchCooperativeKernel((void*)fold_kernel2_coop<CONTRA>, dim3(grid), dim3(nt), params, smem, st));
    }
    CU(h, cudaGetLastError());
    h->stats.kernel_launches++;
#ifdef RNA_DEV_SWITCHES
    if (dbg_roles) fold_dbg_report(bk.mode, bk.Lcap, a, d_dbg, st, nt_of[k], grid_of[k], bk.mode == MODE_SMEM ? smem_need(bk.Lcap) : 0);
#endif
  }
  if (nlanes > 1) {
    for (int x = 0; x < nlanes; x++) { CU(h, cudaEventRecord(h->ev_join[x], h->aux[x])); CU(h, cudaStreamWaitEvent(st_main, h->ev_join[x], 0)); }
  }
  if (d_dbg) cudaFree(d_dbg);
  return RNA_OK;
}

// ---------------------------------------------------------------------------------------------------
// FAST numeric mode (fast_kernel.cuh): sequences up to RNA_FAST_COOP_MIN nt one CTA each, in launches by length class
// (the slot of a CTA is sized for the class); longer ones one at a time on a cooperative grid.  The centroid
// estimator then runs on the packed BPP matrices like rna_centroid_batch (it is exact max-plus either way).
// ---------------------------------------------------------------------------------------------------
#define RNA_FAST_COOP_MIN 700
#define RNA_FASTNUM_WAVE_MS 2300.0   // a full wave of one-CTA sequences of 1024 nt on the FAST batch kernel (measured)
template <bool CONTRA, class real>
static int launch_fast(rna_handle* h, const RnaFoldBatchDev* b, cudaStream_t st, float* d_bpp, const std::vector<uint32_t>* subset = nullptr) {
  const uint32_t* ho = b->h_offsets;
  std::vector<uint32_t> order;
  order_by_length(ho, b->n_seqs, order, subset);
  const uint32_t n = (uint32_t)order.size();   // sequences of this launch
  auto len_of = [&](uint32_t pos) { return (int)(ho[order[pos] + 1] - ho[order[pos]]); };
  std::vector<Bucket> buckets;
  for (uint32_t pos = 0; pos < n;) {
    const int L = len_of(pos);
    Bucket bk;
    bk.begin = pos;
    if (L >= RNA_FAST_COOP_MIN || subset) { bk.mode = MODE_COOP; bk.Lcap = L; bk.end = pos + 1; }   // (a subset = the sequences chosen for the cooperative grid)
    else {
      bk.mode = MODE_GLOBAL;
      bk.Lcap = L;
      const int lo = std::max(L / 2, (L > 96) ? 96 : 0);   // a class: lengths in (lo, L] (the shortest class takes all <= 96)
      uint32_t e = pos;
      while (e < n && len_of(e) > lo) e++;
      bk.end = e;
    }
    buckets.push_back(bk);
    pos = bk.end;
  }
  TRY(ensure(h, h->order, sizeof(uint32_t) * (size_t)std::max<uint32_t>(n, 1)));
  TRY(ensure(h, h->counters, sizeof(int) * buckets.size()));
  CU(h, cudaMemcpyAsync(h->order.p, order.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemsetAsync(h->counters.p, 0, sizeof(int) * buckets.size(), st));
  h->stats.h2d_bytes += sizeof(uint32_t) * n;
  int occ_b = 1, occ_c = 1;
  TRY(kern_prepare(h, (const void*)fast_fold_kernel<CONTRA, false, real>, RNA_FAST_NT_BATCH, 0, &occ_b));
  TRY(kern_prepare(h, (const void*)fast_fold_kernel<CONTRA, true, real>, RNA_FAST_NT_COOP, 0, &occ_c));
  size_t freeb = 0, totb = 0;
  cudaMemGetInfo(&freeb, &totb);
  const size_t budget = (freeb + h->fast_ws.cap) / 2;
  size_t ws_bytes = 0;
  std::vector<int> grid_of(buckets.size(), 0);
  std::vector<size_t> stride_of(buckets.size(), 0);
  for (size_t k = 0; k < buckets.size(); k++) {
    const Bucket& bk = buckets[k];
    const size_t per = (fast_slot_bytes(bk.Lcap, sizeof(real)) + 255) / 256 * 256;
    stride_of[k] = per;
    int g = 1;
    if (bk.mode != MODE_COOP) {
      g = (int)std::min<size_t>(bk.end - bk.begin, (size_t)std::max(1, occ_b) * h->sm_count);
      while (g > 1 && (size_t)g * per > budget) g--;
    }
    if ((size_t)g * per > budget) { h->err = "sequence too long for device memory"; return RNA_ERR_NOMEM; }
    grid_of[k] = g;
    ws_bytes = std::max(ws_bytes, (size_t)g * per);
  }
  TRY(ensure(h, h->fast_ws, ws_bytes));
  FastArgs a;
  memset(&a, 0, sizeof a);
  a.bases = b->d_bases;
  a.offsets = b->d_offsets;
  a.tables = CONTRA ? (const void*)h->d_contra : (const void*)h->d_turner;
  a.allows_short = b->allows_short_hairpins;
  a.out_logz = b->d_out_logz;
  a.out_bpp = d_bpp;
  a.bpp_offsets = b->d_bpp_offsets;
  a.workspace = h->fast_ws.p;
  for (size_t k = 0; k < buckets.size(); k++) {
    const Bucket& bk = buckets[k];
    a.order = (const uint32_t*)h->order.p + bk.begin;
    a.n_launch = bk.end - bk.begin;
    a.work_counter = (int*)h->counters.p + k;
    a.ws_stride = stride_of[k];
    if (bk.mode == MODE_COOP) {
      const int grid = std::max(1, std::min(occ_c, 1)) * h->sm_count;
      void* params[] = {(void*)&a};
      CU(h, cudaLaunchCooperativeKernel((void*)fast_fold_kernel<CONTRA, true, real>, dim3(grid), dim3(RNA_FAST_NT_COOP), params, 0, st));
    } else {
      fast_fold_kernel<CONTRA, false, real><<<grid_of[k], RNA_FAST_NT_BATCH, 0, st>>>(a);
    }
    CU(h, cudaGetLastError());
    h->stats.kernel_launches++;
  }
  return RNA_OK;
}

// FAST_F64 runs everything on fast_fold_kernel (f64 state: the accuracy mode).  FAST_F32 is the speed mode: the few
// long sequences that would otherwise each wait for one CTA go one at a time to fast_fold_kernel's cooperative grid
// (warp-shuffle reductions: 29 / 76 / 284 ms at 1 / 2 / 4 k nt), everything else to the FAST build of the batch kernel
// (fold_fastnum.cu), which is several times faster than fast_fold_kernel's one-CTA-per-sequence launch on batches.
static int launch_fast_mode(rna_handle* h, const RnaFoldBatchDev* b, cudaStream_t st) {
  const bool f64 = h->numeric_mode == RNA_NUMERIC_FAST_F64;
  const bool contra = b->model == RNA_MODEL_CONTRA;
  const uint32_t n = b->n_seqs;
  const uint32_t* ho = b->h_offsets;
  std::vector<uint32_t> sel_coop, sel_batch;
  if (!f64) {
    std::vector<uint32_t> order;
    order_by_length(ho, n, order);
    auto len_of = [&](uint32_t pos) { return (int)(ho[order[pos] + 1] - ho[order[pos]]); };
    uint32_t n_long = 0, forced = 0;
    while (n_long < n && len_of(n_long) > 220) n_long++;          // (beyond the shared-memory mode)
    while (forced < n_long && len_of(forced) > 1024) forced++;    // (beyond the one-CTA modes)
    // cost model from B200 measurements, like launch_fold_model's: the longest c long sequences go to the cooperative grid
    auto coop_ms = [](int L) { const double r = L / 1024.0; return 29.0 * (r <= 1.0 ? r * std::sqrt(r) : r * r); };
    auto wave_ms = [&](int Lmax, uint32_t cnt) {
      if (cnt == 0) return 0.0;
      const double r = Lmax / 1024.0, per_wave = 2.0 * h->sm_count;
      const double waves = cnt / per_wave;
      return RNA_FASTNUM_WAVE_MS * r * r * r * std::max(waves, std::min(1.0, 0.55 + 0.65 * waves));
    };
    double best = 1e300, prefix = 0;
    uint32_t best_c = forced;
    for (uint32_t c = 0; c <= std::min<uint32_t>(n_long, forced + 256); c++) {
      if (c >= forced) {
        const double cost = prefix + (c < n_long ? wave_ms(len_of(c), n_long - c) : 0.0);
        if (cost < best) { best = cost; best_c = c; }
      }
      if (c < n_long) prefix += coop_ms(len_of(c));
    }
    sel_coop.assign(order.begin(), order.begin() + best_c);
    sel_batch.assign(order.begin() + best_c, order.end());
  }
  const bool all_fast_kernel = f64;
  // the packed BPP matrices feed the centroid kernels of the fast_fold_kernel part: the caller's buffer, or one of the handle's
  float* d_bpp = b->d_out_bpp;
  RnaFoldBatchDev bb = *b;
  if (!d_bpp && b->n_gammas && (all_fast_kernel || !sel_coop.empty())) {
    if (!b->d_bpp_offsets) { h->err = "d_bpp_offsets required (FAST mode computes the centroid from the packed BPPs)"; return RNA_ERR_BAD_ARG; }
    uint64_t tot = 0;
    for (uint32_t s = 0; s < b->n_seqs; s++) tot += rna_bpp_len(b->h_offsets[s + 1] - b->h_offsets[s]);
    TRY(ensure(h, h->fast_bpp, tot * 4));
    d_bpp = (float*)h->fast_bpp.p;
  }
  bb.d_out_bpp = d_bpp;
  if (all_fast_kernel) {
    if (contra) TRY((launch_fast<true, double>(h, b, st, d_bpp))); else TRY((launch_fast<false, double>(h, b, st, d_bpp)));
    if (b->n_gammas) TRY(launch_centroid(h, &bb, st, d_bpp));
    return RNA_OK;
  }
  if (!sel_coop.empty()) {
    if (contra) TRY((launch_fast<true, float>(h, b, st, d_bpp, &sel_coop))); else TRY((launch_fast<false, float>(h, b, st, d_bpp, &sel_coop)));
    if (b->n_gammas) TRY(launch_centroid(h, &bb, st, d_bpp, &sel_coop));
  }
  if (!sel_batch.empty()) {
    if (contra) TRY(launch_fold_model<true>(h, b, st, &sel_batch, true)); else TRY(launch_fold_model<false>(h, b, st, &sel_batch, true));
  }
  return RNA_OK;
}

static int check_fold_args(rna_handle* h, const RnaFoldBatchDev* b) {
  if (!h || !b) return RNA_ERR_BAD_ARG;
  if (b->n_seqs == 0) return RNA_OK;
  if (!b->h_offsets || !b->d_bases || !b->d_offsets) { h->err = "null input pointer"; return RNA_ERR_BAD_ARG; }
  if (b->model != RNA_MODEL_TURNER && b->model != RNA_MODEL_CONTRA) { h->err = "bad model"; return RNA_ERR_BAD_ARG; }
  if (b->model == RNA_MODEL_TURNER && !h->has_turner) { h->err = "Turner tables not set"; return RNA_ERR_NO_TABLES; }
  if (b->model == RNA_MODEL_CONTRA && !h->has_contra) { h->err = "CONTRAfold tables not set"; return RNA_ERR_NO_TABLES; }
  if (b->d_out_bpp && !b->d_bpp_offsets) { h->err = "d_bpp_offsets required"; return RNA_ERR_BAD_ARG; }
  if (b->n_gammas && !b->d_gammas) { h->err = "d_gammas required"; return RNA_ERR_BAD_ARG; }
  if (b->d_out_sums && !b->d_sums_offsets) { h->err = "d_sums_offsets required"; return RNA_ERR_BAD_ARG; }
  if ((b->d_out_sums || b->inside_only) && h->numeric_mode != RNA_NUMERIC_REF_EXACT) {
    h->err = "the FoldSums / FoldScores outputs exist in the reference-exact numeric mode only";
    return RNA_ERR_BAD_ARG;
  }
  if (rna_validate_fold_lengths(b->h_offsets, b->n_seqs) != RNA_OK) { h->err = "sequence longer than RNA_MAX_FOLD_LEN (46340 nt)"; return RNA_ERR_TOO_LONG; }
  return RNA_OK;
}

extern "C" int rna_mccaskill_centroid_batch_dev(rna_handle* h, const RnaFoldBatchDev* b, void* stream) {
  TRY(check_fold_args(h, b));
  if (b->n_seqs == 0) return RNA_OK;
  CU(h, cudaSetDevice(h->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  TRY(call_begin(h, st));
  if (h->numeric_mode != RNA_NUMERIC_REF_EXACT) TRY(launch_fast_mode(h, b, st));
  else if (b->model == RNA_MODEL_CONTRA) TRY(launch_fold_model<true>(h, b, st));
  else TRY(launch_fold_model<false>(h, b, st));
  return call_end(h, st);
}

// ---------------------------------------------------------------------------------------------------
// host-buffer entry points
// ---------------------------------------------------------------------------------------------------
static int h2d(rna_handle* h, DevBuf& buf, const void* src, size_t bytes, cudaStream_t st) {
  TRY(ensure(h, buf, bytes));
  if (bytes) CU(h, cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, st));
  h->stats.h2d_bytes += bytes;
  return RNA_OK;
}
static int d2h(rna_handle* h, void* dst, const DevBuf& buf, size_t bytes, cudaStream_t st) {
  if (bytes) CU(h, cudaMemcpyAsync(dst, buf.p, bytes, cudaMemcpyDeviceToHost, st));
  h->stats.d2h_bytes += bytes;
  return RNA_OK;
}

extern "C" int rna_mccaskill_centroid_batch(rna_handle* h, const uint8_t* bases, const uint32_t* offsets,
                                            uint32_t n_seqs, int model, int allows_short_hairpins,
                                            const float* gammas, uint32_t n_gammas, float* out_logz,
                                            float* out_bpp, const uint64_t* bpp_offsets, uint8_t* out_structs,
                                            float* out_expect_acc) {
  if (!h) return RNA_ERR_BAD_ARG;
  h->stats = RnaCallStats{};
  if (n_seqs == 0) return RNA_OK;
  int rc = rna_validate_bases(bases, offsets, n_seqs);
  if (rc == RNA_OK) rc = rna_validate_fold_lengths(offsets, n_seqs);
  if (rc != RNA_OK) { h->err = rc == RNA_ERR_TOO_LONG ? "sequence longer than RNA_MAX_FOLD_LEN (46340 nt)" : "input validation failed"; return rc; }
  if (offsets[0] != 0) { h->err = "offsets[0] must be 0"; return RNA_ERR_BAD_ARG; }
  if (n_gammas && !gammas) return RNA_ERR_BAD_ARG;
  CU(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const uint32_t total = offsets[n_seqs];
  std::vector<uint64_t> own_off;
  uint32_t maxlen = 0;
  for (uint32_t s = 0; s < n_seqs; s++) maxlen = std::max(maxlen, offsets[s + 1] - offsets[s]);
  // (FAST numeric modes compute the centroid from the packed BPP matrices: offsets are needed even without out_bpp)
  const bool fast_cent = h->numeric_mode != RNA_NUMERIC_REF_EXACT && n_gammas;
  if ((out_bpp || fast_cent) && !bpp_offsets) {
    own_off.resize((size_t)n_seqs + 1);
    own_off[0] = 0;
    for (uint32_t s = 0; s < n_seqs; s++) own_off[s + 1] = own_off[s] + rna_bpp_len(offsets[s + 1] - offsets[s]);
    bpp_offsets = own_off.data();
  }
  RnaFoldBatchDev b;
  memset(&b, 0, sizeof b);
  b.h_offsets = offsets;
  TRY(h2d(h, h->b_bases, bases, total, st));
  TRY(h2d(h, h->b_offsets, offsets, sizeof(uint32_t) * ((size_t)n_seqs + 1), st));
  b.d_bases = (const uint8_t*)h->b_bases.p;
  b.d_offsets = (const uint32_t*)h->b_offsets.p;
  size_t bpp_total = 0;
  bool bpp_zero_copy = false;
  if (fast_cent && !out_bpp) {
    TRY(h2d(h, h->b_bppoff, bpp_offsets, sizeof(uint64_t) * ((size_t)n_seqs + 1), st));
    b.d_bpp_offsets = (const uint64_t*)h->b_bppoff.p;
  }
  if (out_bpp) {
    bpp_total = bpp_offsets[n_seqs];
    TRY(h2d(h, h->b_bppoff, bpp_offsets, sizeof(uint64_t) * ((size_t)n_seqs + 1), st));
    b.d_bpp_offsets = (const uint64_t*)h->b_bppoff.p;
    // The packed BPP matrices are by far the largest output (11 KB per tRNA) and the kernels only ever WRITE them, in
    // coalesced row segments.  When the caller's buffer is page-locked (cudaHostAlloc / cudaHostRegister) the kernels
    // write it directly over PCIe while they compute, instead of staging in HBM and copying after the last kernel;
    // a pageable buffer takes the staged path.  (RNA_NO_ZERO_COPY=1 forces staging.)
    static const bool no_zc = dev_env("RNA_NO_ZERO_COPY") != nullptr;
    cudaPointerAttributes pat;
    memset(&pat, 0, sizeof pat);
    // (the centroid kernels of the FAST modes re-read the BPPs: staged in HBM)
    if (!no_zc && !fast_cent && cudaPointerGetAttributes(&pat, out_bpp) == cudaSuccess && pat.type == cudaMemoryTypeHost && pat.devicePointer) {
      // (the whole range must be page-locked: probe its last byte too)
      cudaPointerAttributes pend;
      memset(&pend, 0, sizeof pend);
      if (cudaPointerGetAttributes(&pend, reinterpret_cast<const char*>(out_bpp) + bpp_total * 4 - 1) == cudaSuccess &&
          pend.type == cudaMemoryTypeHost && pend.devicePointer) {
        bpp_zero_copy = true;
        b.d_out_bpp = reinterpret_cast<float*>(pat.devicePointer);
      }
    }
    cudaGetLastError();   // (older drivers report an unregistered pointer as an error)
    if (!bpp_zero_copy) {
      TRY(ensure(h, h->b_bpp, bpp_total * 4));
      b.d_out_bpp = (float*)h->b_bpp.p;
    }
  }
  if (n_gammas) {
    TRY(h2d(h, h->b_gammas, gammas, sizeof(float) * n_gammas, st));
    b.d_gammas = (const float*)h->b_gammas.p;
    if (out_structs) { TRY(ensure(h, h->b_structs, (size_t)n_gammas * total)); b.d_out_structs = (uint8_t*)h->b_structs.p; }
    if (out_expect_acc) { TRY(ensure(h, h->b_ea, sizeof(float) * (size_t)n_gammas * n_seqs)); b.d_out_expect_acc = (float*)h->b_ea.p; }
  }
  if (out_logz) { TRY(ensure(h, h->b_logz, sizeof(float) * n_seqs)); b.d_out_logz = (float*)h->b_logz.p; }
  b.n_seqs = n_seqs;
  b.total_len = total;
  b.max_len = maxlen;
  b.model = model;
  b.allows_short_hairpins = allows_short_hairpins;
  b.n_gammas = n_gammas;
  TRY(rna_mccaskill_centroid_batch_dev(h, &b, st));
  if (out_logz) TRY(d2h(h, out_logz, h->b_logz, sizeof(float) * n_seqs, st));
  if (out_bpp && !bpp_zero_copy) TRY(d2h(h, out_bpp, h->b_bpp, bpp_total * 4, st));
  if (out_bpp && bpp_zero_copy) h->stats.d2h_bytes += bpp_total * 4;   // written to host memory by the kernels
  if (n_gammas && out_structs) TRY(d2h(h, out_structs, h->b_structs, (size_t)n_gammas * total, st));
  if (n_gammas && out_expect_acc) TRY(d2h(h, out_expect_acc, h->b_ea, sizeof(float) * (size_t)n_gammas * n_seqs, st));
  CU(h, cudaStreamSynchronize(st));
  return RNA_OK;
}

extern "C" int rna_fold_sums_batch(rna_handle* h, const uint8_t* bases, const uint32_t* offsets, uint32_t n_seqs, int model,
                                   int allows_short_hairpins, float* out_sums, const uint64_t* sums_offsets, float* out_logz) {
  if (!h || !out_sums) return RNA_ERR_BAD_ARG;
  h->stats = RnaCallStats{};
  if (n_seqs == 0) return RNA_OK;
  int rc = rna_validate_bases(bases, offsets, n_seqs);
  if (rc == RNA_OK) rc = rna_validate_fold_lengths(offsets, n_seqs);
  if (rc != RNA_OK) { h->err = rc == RNA_ERR_TOO_LONG ? "sequence longer than RNA_MAX_FOLD_LEN (46340 nt)" : "input validation failed"; return rc; }
  if (offsets[0] != 0) { h->err = "offsets[0] must be 0"; return RNA_ERR_BAD_ARG; }
  CU(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  std::vector<uint64_t> own;
  if (!sums_offsets) {
    own.resize((size_t)n_seqs + 1);
    own[0] = 0;
    for (uint32_t s = 0; s < n_seqs; s++) own[s + 1] = own[s] + (uint64_t)RNA_SUMS_PLANES * rna_sums_len(offsets[s + 1] - offsets[s]);
    sums_offsets = own.data();
  }
  const size_t tot = sums_offsets[n_seqs];
  RnaFoldBatchDev b;
  memset(&b, 0, sizeof b);
  b.h_offsets = offsets;
  TRY(h2d(h, h->b_bases, bases, offsets[n_seqs], st));
  TRY(h2d(h, h->b_offsets, offsets, sizeof(uint32_t) * ((size_t)n_seqs + 1), st));
  TRY(h2d(h, h->b_proboff, sums_offsets, sizeof(uint64_t) * ((size_t)n_seqs + 1), st));
  TRY(ensure(h, h->b_probs, tot * 4));
  if (out_logz) TRY(ensure(h, h->b_logz, sizeof(float) * n_seqs));
  b.d_bases = (const uint8_t*)h->b_bases.p;
  b.d_offsets = (const uint32_t*)h->b_offsets.p;
  b.d_sums_offsets = (const uint64_t*)h->b_proboff.p;
  b.d_out_sums = (float*)h->b_probs.p;
  b.d_out_logz = out_logz ? (float*)h->b_logz.p : nullptr;
  b.n_seqs = n_seqs;
  b.total_len = offsets[n_seqs];
  b.model = model;
  b.allows_short_hairpins = allows_short_hairpins;
  b.inside_only = 1;
  TRY(rna_mccaskill_centroid_batch_dev(h, &b, st));
  TRY(d2h(h, out_sums, h->b_probs, tot * 4, st));
  if (out_logz) TRY(d2h(h, out_logz, h->b_logz, sizeof(float) * n_seqs, st));
  CU(h, cudaStreamSynchronize(st));
  return RNA_OK;
}

// FoldScores::twoloop_scores of one sequence (twoloop_export.cuh): inside pass for the sums_close keys, count, scan, fill.
extern "C" int rna_twoloop_scores(rna_handle* h, const uint8_t* seq, uint32_t seq_len, int model, int allows_short_hairpins,
                                  RnaTwoloopScore* out, uint64_t capacity, uint64_t* out_count) {
  if (!h || !seq || !out_count || (capacity && !out)) return RNA_ERR_BAD_ARG;
  if (seq_len == 0) return RNA_ERR_EMPTY_SEQ;
  if (seq_len > RNA_MAX_SEQ_LEN) { h->err = "sequence longer than RNA_MAX_SEQ_LEN (positions are u16)"; return RNA_ERR_TOO_LONG; }
  if (h->numeric_mode != RNA_NUMERIC_REF_EXACT) { h->err = "the FoldSums / FoldScores outputs exist in the reference-exact numeric mode only"; return RNA_ERR_BAD_ARG; }
  const uint32_t offsets[2] = {0, seq_len};
  const int L = (int)seq_len;
  const size_t ncell = (size_t)L * ((size_t)L + 1) / 2;
  // inside pass: the planes stay on the device (h->b_probs), only sums_close is read
  {
    int rc = rna_validate_bases(seq, offsets, 1);
    if (rc == RNA_OK) rc = rna_validate_fold_lengths(offsets, 1);
    if (rc != RNA_OK) { h->err = rc == RNA_ERR_TOO_LONG ? "sequence longer than RNA_MAX_FOLD_LEN (46340 nt)" : "input validation failed"; return rc; }
  }
  h->stats = RnaCallStats{};
  CU(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const uint64_t sums_offsets[2] = {0, (uint64_t)RNA_SUMS_PLANES * rna_sums_len(seq_len)};
  RnaFoldBatchDev b;
  memset(&b, 0, sizeof b);
  b.h_offsets = offsets;
  TRY(h2d(h, h->b_bases, seq, seq_len, st));
  TRY(h2d(h, h->b_offsets, offsets, sizeof offsets, st));
  TRY(h2d(h, h->b_proboff, sums_offsets, sizeof sums_offsets, st));
  TRY(ensure(h, h->b_probs, sums_offsets[1] * 4));
  b.d_bases = (const uint8_t*)h->b_bases.p;
  b.d_offsets = (const uint32_t*)h->b_offsets.p;
  b.d_sums_offsets = (const uint64_t*)h->b_proboff.p;
  b.d_out_sums = (float*)h->b_probs.p;
  b.n_seqs = 1;
  b.total_len = seq_len;
  b.model = model;
  b.allows_short_hairpins = allows_short_hairpins;
  b.inside_only = 1;
  TRY(rna_mccaskill_centroid_batch_dev(h, &b, st));
  // codes, counts
  TRY(ensure(h, h->b_pairs, 3 * ((size_t)L + 16)));
  TRY(ensure(h, h->b_bppoff, sizeof(unsigned long long) * ncell));
  TwoloopArgs a;
  memset(&a, 0, sizeof a);
  uint8_t* codes = (uint8_t*)h->b_pairs.p;
  a.sq = codes + 4; a.RR = codes + (L + 16); a.LL = codes + 2 * ((size_t)L + 16);
  a.L = L;
  a.tables = model == RNA_MODEL_CONTRA ? (const void*)h->d_contra : (const void*)h->d_turner;
  a.allows_short = allows_short_hairpins;
  a.close = (const float*)h->b_probs.p + (size_t)RNA_SUMS_CLOSE * rna_sums_len(seq_len);
  a.offsets = (unsigned long long*)h->b_bppoff.p;
  const int nt = 128, grid = (int)std::min<size_t>((ncell + nt - 1) / nt, (size_t)h->sm_count * 16);
  twoloop_codes_kernel<<<std::max(1, (L + 8 + 255) / 256), 256, 0, st>>>((const uint8_t*)h->b_bases.p, L, codes, const_cast<uint8_t*>(a.RR), const_cast<uint8_t*>(a.LL));
  if (model == RNA_MODEL_CONTRA) twoloop_kernel<true, false><<<grid, nt, 0, st>>>(a); else twoloop_kernel<false, false><<<grid, nt, 0, st>>>(a);
  CU(h, cudaGetLastError());
  std::vector<unsigned long long> cnt(ncell);
  CU(h, cudaMemcpyAsync(cnt.data(), a.offsets, sizeof(unsigned long long) * ncell, cudaMemcpyDeviceToHost, st));
  CU(h, cudaStreamSynchronize(st));
  unsigned long long run = 0;
  for (size_t c = 0; c < ncell; c++) { const unsigned long long n = cnt[c]; cnt[c] = run; run += n; }
  *out_count = run;
  h->stats.kernel_launches += 2;
  const uint64_t nwrite = std::min<uint64_t>(run, capacity);
  if (nwrite == 0) return RNA_OK;
  CU(h, cudaMemcpyAsync(a.offsets, cnt.data(), sizeof(unsigned long long) * ncell, cudaMemcpyHostToDevice, st));
  TRY(ensure(h, h->b_bpp, nwrite * sizeof(RnaTwoloopScore)));
  a.out = (RnaTwoloopScore*)h->b_bpp.p;
  a.capacity = nwrite;
  if (model == RNA_MODEL_CONTRA) twoloop_kernel<true, true><<<grid, nt, 0, st>>>(a); else twoloop_kernel<false, true><<<grid, nt, 0, st>>>(a);
  CU(h, cudaGetLastError());
  h->stats.kernel_launches++;
  TRY(d2h(h, out, h->b_bpp, nwrite * sizeof(RnaTwoloopScore), st));
  CU(h, cudaStreamSynchronize(st));
  return RNA_OK;
}

extern "C" int rna_mccaskill_batch(rna_handle* h, const uint8_t* bases, const uint32_t* offsets, uint32_t n_seqs,
                                   int model, int allows_short_hairpins, float* out_logz, float* out_bpp,
                                   const uint64_t* bpp_offsets) {
  return rna_mccaskill_centroid_batch(h, bases, offsets, n_seqs, model, allows_short_hairpins, nullptr, 0,
                                      out_logz, out_bpp, bpp_offsets, nullptr, nullptr);
}

static int centroid_host(rna_handle* h, const float* bpp, const uint64_t* bpp_offsets, const uint32_t* offsets,
                         uint32_t n_seqs, const float* gammas, uint32_t n_gammas, uint8_t* out_structs,
                         float* out_expect_acc, uint16_t* out_pairs, uint32_t* out_num_pairs) {
  if (!h) return RNA_ERR_BAD_ARG;
  h->stats = RnaCallStats{};
  if (n_seqs == 0 || n_gammas == 0) return RNA_OK;
  if (!bpp || !offsets || !gammas) return RNA_ERR_BAD_ARG;
  if (offsets[0] != 0) { h->err = "offsets[0] must be 0"; return RNA_ERR_BAD_ARG; }
  for (uint32_t s = 0; s < n_seqs; s++) {
    if (offsets[s + 1] <= offsets[s]) return RNA_ERR_EMPTY_SEQ;
    if (offsets[s + 1] - offsets[s] > (uint32_t)RNA_MAX_FOLD_LEN) return RNA_ERR_TOO_LONG;
  }
  CU(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const uint32_t total = offsets[n_seqs];
  std::vector<uint64_t> own_off;
  if (!bpp_offsets) {
    own_off.resize((size_t)n_seqs + 1);
    own_off[0] = 0;
    for (uint32_t s = 0; s < n_seqs; s++) own_off[s + 1] = own_off[s] + rna_bpp_len(offsets[s + 1] - offsets[s]);
    bpp_offsets = own_off.data();
  }
  const size_t bpp_total = bpp_offsets[n_seqs];
  RnaFoldBatchDev b;
  memset(&b, 0, sizeof b);
  b.h_offsets = offsets;
  TRY(h2d(h, h->b_offsets, offsets, sizeof(uint32_t) * ((size_t)n_seqs + 1), st));
  TRY(h2d(h, h->b_bppoff, bpp_offsets, sizeof(uint64_t) * ((size_t)n_seqs + 1), st));
  TRY(h2d(h, h->b_bpp, bpp, bpp_total * 4, st));
  TRY(h2d(h, h->b_gammas, gammas, sizeof(float) * n_gammas, st));
  b.d_offsets = (const uint32_t*)h->b_offsets.p;
  b.d_bpp_offsets = (const uint64_t*)h->b_bppoff.p;
  b.d_gammas = (const float*)h->b_gammas.p;
  b.n_seqs = n_seqs;
  b.total_len = total;
  b.n_gammas = n_gammas;
  if (out_structs) { TRY(ensure(h, h->b_structs, (size_t)n_gammas * total)); b.d_out_structs = (uint8_t*)h->b_structs.p; }
  if (out_expect_acc) { TRY(ensure(h, h->b_ea, sizeof(float) * (size_t)n_gammas * n_seqs)); b.d_out_expect_acc = (float*)h->b_ea.p; }
  if (out_pairs) { TRY(ensure(h, h->b_pairs, sizeof(uint16_t) * 2 * (size_t)n_gammas * total)); b.d_out_pairs = (uint16_t*)h->b_pairs.p; }
  if (out_num_pairs) { TRY(ensure(h, h->b_logz, sizeof(uint32_t) * (size_t)n_gammas * n_seqs)); b.d_out_num_pairs = (uint32_t*)h->b_logz.p; }
  TRY(call_begin(h, st));
  TRY(launch_centroid(h, &b, st, (const float*)h->b_bpp.p));
  TRY(call_end(h, st));
  if (out_structs) TRY(d2h(h, out_structs, h->b_structs, (size_t)n_gammas * total, st));
  if (out_expect_acc) TRY(d2h(h, out_expect_acc, h->b_ea, sizeof(float) * (size_t)n_gammas * n_seqs, st));
  if (out_pairs) TRY(d2h(h, out_pairs, h->b_pairs, sizeof(uint16_t) * 2 * (size_t)n_gammas * total, st));
  if (out_num_pairs) TRY(d2h(h, out_num_pairs, h->b_logz, sizeof(uint32_t) * (size_t)n_gammas * n_seqs, st));
  CU(h, cudaStreamSynchronize(st));
  return RNA_OK;
}

extern "C" int rna_centroid_batch(rna_handle* h, const float* bpp, const uint64_t* bpp_offsets,
                                  const uint32_t* offsets, uint32_t n_seqs, const float* gammas, uint32_t n_gammas,
                                  uint8_t* out_structs, float* out_expect_acc) {
  return centroid_host(h, bpp, bpp_offsets, offsets, n_seqs, gammas, n_gammas, out_structs, out_expect_acc, nullptr,
                       nullptr);
}

extern "C" int rna_mccaskill_algo(rna_handle* h, const uint8_t* seq, uint32_t seq_len, int uses_contra_model,
                                  int allows_short_hairpins, float* out_bpp, float* out_logz) {
  const uint32_t off[2] = {0, seq_len};
  return rna_mccaskill_centroid_batch(h, seq, off, 1, uses_contra_model ? RNA_MODEL_CONTRA : RNA_MODEL_TURNER,
                                      allows_short_hairpins, nullptr, 0, out_logz, out_bpp, nullptr, nullptr, nullptr);
}

extern "C" int rna_centroid_fold(rna_handle* h, const float* bpp, uint32_t seq_len, float centroid_threshold,
                                 uint8_t* out_fold_str, uint16_t* out_pairs, uint32_t* out_num_pairs,
                                 float* out_expect_accuracy) {
  const uint32_t off[2] = {0, seq_len};
  return centroid_host(h, bpp, nullptr, off, 1, &centroid_threshold, 1, out_fold_str, out_expect_accuracy, out_pairs,
                       out_num_pairs);
}

// ---------------------------------------------------------------------------------------------------
// Durbin
// ---------------------------------------------------------------------------------------------------
extern "C" int rna_durbin_batch_dev(rna_handle* h, const RnaDurbinBatchDev* b, void* stream) {
  if (!h || !b) return RNA_ERR_BAD_ARG;
  if (b->n_pairs == 0) return RNA_OK;
  if (!h->has_align) { h->err = "align tables not set"; return RNA_ERR_NO_TABLES; }
  if (!b->h_offsets || !b->h_pairs || !b->d_bases || !b->d_offsets || !b->d_pairs || !b->d_prob_offsets ||
      !b->d_out_probs) { h->err = "null pointer"; return RNA_ERR_BAD_ARG; }
  CU(h, cudaSetDevice(h->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  TRY(call_begin(h, st));
  const uint32_t np = b->n_pairs;
  const uint32_t* ho = b->h_offsets;
  std::vector<uint32_t> order(np);
  int ncap = 0, mcap = 0;
  for (uint32_t p = 0; p < np; p++) {
    order[p] = p;
    const uint32_t sa = b->h_pairs[2 * p], sb = b->h_pairs[2 * p + 1];
    if (sa >= b->n_seqs || sb >= b->n_seqs) { h->err = "pair index out of range"; return RNA_ERR_BAD_ARG; }
    ncap = std::max(ncap, (int)(ho[sa + 1] - ho[sa]) + 2);
    mcap = std::max(mcap, (int)(ho[sb + 1] - ho[sb]) + 2);
  }
  auto cost = [&](uint32_t p) {
    const uint32_t sa = b->h_pairs[2 * p], sb = b->h_pairs[2 * p + 1];
    return (uint64_t)(ho[sa + 1] - ho[sa] + 2) * (uint64_t)(ho[sb + 1] - ho[sb] + 2);
  };
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return cost(x) > cost(y); });
  TRY(ensure(h, h->order, sizeof(uint32_t) * np));
  TRY(ensure(h, h->counters, sizeof(int)));
  CU(h, cudaMemcpyAsync(h->order.p, order.data(), sizeof(uint32_t) * np, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemsetAsync(h->counters.p, 0, sizeof(int), st));
  h->stats.h2d_bytes += sizeof(uint32_t) * np;
  DurbinArgs a;
  memset(&a, 0, sizeof a);
  a.bases = b->d_bases;
  a.offsets = b->d_offsets;
  a.pairs = b->d_pairs;
  a.order = (const uint32_t*)h->order.p;
  a.n_pairs = np;
  a.prob_offsets = b->d_prob_offsets;
  a.out_probs = b->d_out_probs;
  a.tables = h->d_align;
  a.work_counter = (int*)h->counters.p;
  a.ncap = ncap;
  a.mcap = mcap;
  a.roll_in_smem = durbin_smem_bytes(ncap, mcap, true) + 1024 <= h->smem_optin;
  // the n x m forward / posterior matrix on chip when at least 6 CTAs per SM still fit (tRNA-length pairs)
  // (on chip it would cap the residency at 6 CTAs of 3 warps per SM for tRNA-length pairs — measured 3.07 M pairs/s
  // against 4.4 M with a global slot per CTA, which stays L2-resident: 2 x SM-count x ~20 CTAs x 33 KB)
  static const bool park_smem = dev_env("RNA_DURBIN_PARK_SMEM") != nullptr;
  a.park_in_smem = park_smem && a.roll_in_smem && 6 * (durbin_smem_bytes(ncap, mcap, true, true) + 1024) <= (size_t)233472;
  const size_t smem = durbin_smem_bytes(ncap, mcap, a.roll_in_smem != 0, a.park_in_smem != 0);
  if (smem + 1024 > h->smem_optin) { h->err = "sequence too long for the Durbin kernel"; return RNA_ERR_TOO_LONG; }
  const int nt = std::min(256, std::max(32, (ncap + 31) / 32 * 32));
  int occ = 1;
  TRY(kern_prepare(h, (const void*)durbin_kernel, nt, smem, &occ));
  const int grid = (int)std::min<size_t>(np, (size_t)std::max(1, occ) * h->sm_count);
  {
    a.ws_stride = a.roll_in_smem ? 0 : (size_t)9 * ncap + 32;
    a.park_stride = a.park_in_smem ? 0 : ((size_t)ncap * mcap + 31) / 32 * 32;
    const size_t per = a.ws_stride + a.park_stride;
    if (per) {
      TRY(ensure(h, h->ws, (size_t)grid * per * 4));
      a.workspace = (float*)h->ws.p;
      a.park_ws = (float*)h->ws.p + (size_t)grid * a.ws_stride;
    }
  }
  durbin_kernel<<<grid, nt, smem, st>>>(a);
  CU(h, cudaGetLastError());
  h->stats.kernel_launches++;
  return call_end(h, st);
}

extern "C" int rna_durbin_batch(rna_handle* h, const uint8_t* bases, const uint32_t* offsets, uint32_t n_seqs,
                                const uint32_t* pairs, uint32_t n_pairs, float* out_probs,
                                const uint64_t* prob_offsets) {
  if (!h) return RNA_ERR_BAD_ARG;
  h->stats = RnaCallStats{};
  if (n_pairs == 0) return RNA_OK;
  int rc = rna_validate_bases(bases, offsets, n_seqs);
  if (rc != RNA_OK) { h->err = "input validation failed"; return rc; }
  if (!pairs || !out_probs) return RNA_ERR_BAD_ARG;
  if (offsets[0] != 0) { h->err = "offsets[0] must be 0"; return RNA_ERR_BAD_ARG; }
  for (uint32_t p = 0; p < 2 * n_pairs; p++)
    if (pairs[p] >= n_seqs) { h->err = "pair index out of range"; return RNA_ERR_BAD_ARG; }
  CU(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  std::vector<uint64_t> own;
  if (!prob_offsets) {
    own.resize((size_t)n_pairs + 1);
    own[0] = 0;
    for (uint32_t p = 0; p < n_pairs; p++) {
      const uint64_t la = offsets[pairs[2 * p] + 1] - offsets[pairs[2 * p]];
      const uint64_t lb = offsets[pairs[2 * p + 1] + 1] - offsets[pairs[2 * p + 1]];
      own[p + 1] = own[p] + (la + 2) * (lb + 2);
    }
    prob_offsets = own.data();
  }
  const size_t tot = prob_offsets[n_pairs];
  TRY(h2d(h, h->b_bases, bases, offsets[n_seqs], st));
  TRY(h2d(h, h->b_offsets, offsets, sizeof(uint32_t) * ((size_t)n_seqs + 1), st));
  TRY(h2d(h, h->b_pairs, pairs, sizeof(uint32_t) * 2 * (size_t)n_pairs, st));
  TRY(h2d(h, h->b_proboff, prob_offsets, sizeof(uint64_t) * ((size_t)n_pairs + 1), st));
  // the kernel writes the match-probability matrices once, whole rows at a time: straight into the caller's buffer when
  // it is page-locked (see rna_mccaskill_centroid_batch), else staged in HBM and copied afterwards
  bool zero_copy = false;
  float* d_target = nullptr;
  {
    static const bool no_zc = dev_env("RNA_NO_ZERO_COPY") != nullptr;
    cudaPointerAttributes p0, p1;
    memset(&p0, 0, sizeof p0); memset(&p1, 0, sizeof p1);
    if (!no_zc && tot && cudaPointerGetAttributes(&p0, out_probs) == cudaSuccess && p0.type == cudaMemoryTypeHost && p0.devicePointer &&
        cudaPointerGetAttributes(&p1, reinterpret_cast<const char*>(out_probs) + tot * 4 - 1) == cudaSuccess &&
        p1.type == cudaMemoryTypeHost && p1.devicePointer) {
      zero_copy = true;
      d_target = reinterpret_cast<float*>(p0.devicePointer);
    }
    cudaGetLastError();
  }
  if (!zero_copy) TRY(ensure(h, h->b_probs, tot * 4));
  RnaDurbinBatchDev b;
  memset(&b, 0, sizeof b);
  b.h_offsets = offsets;
  b.h_pairs = pairs;
  b.d_bases = (const uint8_t*)h->b_bases.p;
  b.d_offsets = (const uint32_t*)h->b_offsets.p;
  b.d_pairs = (const uint32_t*)h->b_pairs.p;
  b.d_prob_offsets = (const uint64_t*)h->b_proboff.p;
  b.n_seqs = n_seqs;
  b.n_pairs = n_pairs;
  b.d_out_probs = zero_copy ? d_target : (float*)h->b_probs.p;
  TRY(rna_durbin_batch_dev(h, &b, st));
  if (!zero_copy) TRY(d2h(h, out_probs, h->b_probs, tot * 4, st));
  else h->stats.d2h_bytes += tot * 4;   // written to host memory by the kernel
  CU(h, cudaStreamSynchronize(st));
  return RNA_OK;
}

extern "C" int rna_durbin_algo(rna_handle* h, const uint8_t* seq_a, uint32_t len_a, const uint8_t* seq_b,
                               uint32_t len_b, float* out_probs) {
  if (!seq_a || !seq_b) return RNA_ERR_BAD_ARG;
  if (len_a == 0 || len_b == 0) return RNA_ERR_EMPTY_SEQ;
  std::vector<uint8_t> cat((size_t)len_a + len_b);
  memcpy(cat.data(), seq_a, len_a);
  memcpy(cat.data() + len_a, seq_b, len_b);
  const uint32_t off[3] = {0, len_a, len_a + len_b};
  const uint32_t pr[2] = {0, 1};
  return rna_durbin_batch(h, cat.data(), off, 2, pr, 1, out_probs, nullptr);
}
