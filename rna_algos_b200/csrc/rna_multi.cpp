// rna_multi.cpp — every GPU of one box behind one object (include/rna_algos_b200.h "rna_multi").  Pure host code over
// the single-device C ABI: a longest-processing-time-first partition of the call's units (rna_partition_lpt, the cost
// model of SURVEY.md §8(e)), one host thread + one rna_handle per device, compact per-device sub-batches, results
// scattered into the caller's buffers (disjoint ranges).  No collective, no peer-to-peer traffic: the units are
// independent, exactly like the tasks of the reference's thread pool (src/bin/centroid_fold.rs:119-132).
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/rna_algos_b200.h"

struct rna_multi {
  std::vector<rna_handle*> h;
  std::vector<int> dev;
  std::string err;
  std::vector<double> busy;
  std::vector<uint32_t> units;
};

extern "C" int rna_multi_create(const int* devices, int n_devices, rna_multi** out) {
  if (!out || n_devices < 0 || (n_devices > 0 && !devices)) return RNA_ERR_BAD_ARG;
  *out = nullptr;
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible <= 0) { cudaGetLastError(); return RNA_ERR_NO_DEVICE; }
  rna_multi* m = new rna_multi();
  if (n_devices == 0) for (int d = 0; d < visible; d++) m->dev.push_back(d);
  else m->dev.assign(devices, devices + n_devices);
  for (int d : m->dev) {
    rna_handle* hh = nullptr;
    const int rc = rna_create(d, &hh);
    if (rc != RNA_OK) { rna_multi_destroy(m); return rc; }
    m->h.push_back(hh);
  }
  m->busy.assign(m->h.size(), 0.0);
  m->units.assign(m->h.size(), 0);
  *out = m;
  return RNA_OK;
}
extern "C" int rna_multi_destroy(rna_multi* m) {
  if (!m) return RNA_OK;
  for (rna_handle* hh : m->h) rna_destroy(hh);
  delete m;
  return RNA_OK;
}
extern "C" int rna_multi_num_devices(const rna_multi* m) { return m ? (int)m->h.size() : 0; }
extern "C" rna_handle* rna_multi_handle(rna_multi* m, int k) { return (m && k >= 0 && k < (int)m->h.size()) ? m->h[k] : nullptr; }
extern "C" const char* rna_multi_last_error(const rna_multi* m) { return m ? m->err.c_str() : "null rna_multi"; }

template <class F>
static int for_all(rna_multi* m, F f) {
  if (!m) return RNA_ERR_BAD_ARG;
  for (size_t k = 0; k < m->h.size(); k++) {
    const int rc = f(m->h[k]);
    if (rc != RNA_OK) { m->err = "device " + std::to_string(m->dev[k]) + ": " + rna_last_error(m->h[k]); return rc; }
  }
  return RNA_OK;
}
extern "C" int rna_multi_set_turner_tables(rna_multi* m, const RnaTurnerTables* t) { return for_all(m, [&](rna_handle* x) { return rna_set_turner_tables(x, t); }); }
extern "C" int rna_multi_set_contra_tables(rna_multi* m, const RnaContraTables* t) { return for_all(m, [&](rna_handle* x) { return rna_set_contra_tables(x, t); }); }
extern "C" int rna_multi_set_align_tables(rna_multi* m, const RnaAlignTables* t) { return for_all(m, [&](rna_handle* x) { return rna_set_align_tables(x, t); }); }
extern "C" int rna_multi_set_numeric_mode(rna_multi* m, int mode) { return for_all(m, [&](rna_handle* x) { return rna_set_numeric_mode(x, mode); }); }

extern "C" int rna_multi_last_shares(const rna_multi* m, double* busy_seconds, uint32_t* units) {
  if (!m) return RNA_ERR_BAD_ARG;
  for (size_t k = 0; k < m->h.size(); k++) {
    if (busy_seconds) busy_seconds[k] = m->busy[k];
    if (units) units[k] = m->units[k];
  }
  return RNA_OK;
}

// runs job(k) for every device with a non-empty share on its own host thread; first failing status wins
template <class F>
static int run_shares(rna_multi* m, const std::vector<std::vector<uint32_t>>& share, F job) {
  const size_t nd = m->h.size();
  std::vector<int> rcs(nd, RNA_OK);
  std::vector<std::thread> th;
  for (size_t k = 0; k < nd; k++) {
    m->units[k] = (uint32_t)share[k].size();
    m->busy[k] = 0.0;
    if (share[k].empty()) continue;
    th.emplace_back([&, k] {
      const auto t0 = std::chrono::steady_clock::now();
      rcs[k] = job(k);
      m->busy[k] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    });
  }
  for (std::thread& t : th) t.join();
  for (size_t k = 0; k < nd; k++)
    if (rcs[k] != RNA_OK) { m->err = "device " + std::to_string(m->dev[k]) + ": " + rna_last_error(m->h[k]); return rcs[k]; }
  return RNA_OK;
}

extern "C" int rna_multi_mccaskill_centroid_batch(rna_multi* m, const uint8_t* bases, const uint32_t* offsets, uint32_t n_seqs,
                                                  int model, int allows_short_hairpins, const float* gammas, uint32_t n_gammas,
                                                  float* out_logz, float* out_bpp, const uint64_t* bpp_offsets,
                                                  uint8_t* out_structs, float* out_expect_acc) {
  if (!m || m->h.empty()) return RNA_ERR_BAD_ARG;
  if (n_seqs == 0) return RNA_OK;
  int rc = rna_validate_bases(bases, offsets, n_seqs);
  if (rc == RNA_OK) rc = rna_validate_fold_lengths(offsets, n_seqs);
  if (rc != RNA_OK) { m->err = "input validation failed"; return rc; }
  if (offsets[0] != 0 || (n_gammas && !gammas)) { m->err = "bad arguments"; return RNA_ERR_BAD_ARG; }
  const size_t nd = m->h.size();
  const uint32_t total = offsets[n_seqs];
  // partition: LPT on L^3 + 500 L^2
  std::vector<uint64_t> cost(n_seqs);
  for (uint32_t s = 0; s < n_seqs; s++) { const uint64_t L = offsets[s + 1] - offsets[s]; cost[s] = L * L * L + 500 * L * L; }
  std::vector<uint32_t> part(n_seqs);
  rna_partition_lpt(cost.data(), n_seqs, (uint32_t)nd, part.data());
  std::vector<std::vector<uint32_t>> share(nd);
  for (uint32_t s = 0; s < n_seqs; s++) share[part[s]].push_back(s);
  std::vector<uint64_t> own_off;
  if (out_bpp && !bpp_offsets) {
    own_off.resize((size_t)n_seqs + 1);
    own_off[0] = 0;
    for (uint32_t s = 0; s < n_seqs; s++) own_off[s + 1] = own_off[s] + rna_bpp_len(offsets[s + 1] - offsets[s]);
    bpp_offsets = own_off.data();
  }
  return run_shares(m, share, [&](size_t k) -> int {
    const std::vector<uint32_t>& mine = share[k];
    const uint32_t n = (uint32_t)mine.size();
    // compact sub-batch
    std::vector<uint32_t> off(n + 1, 0);
    for (uint32_t x = 0; x < n; x++) off[x + 1] = off[x] + (offsets[mine[x] + 1] - offsets[mine[x]]);
    std::vector<uint8_t> b(off[n]);
    for (uint32_t x = 0; x < n; x++) memcpy(b.data() + off[x], bases + offsets[mine[x]], off[x + 1] - off[x]);
    std::vector<uint64_t> boff(n + 1, 0);
    for (uint32_t x = 0; x < n; x++) boff[x + 1] = boff[x] + rna_bpp_len(off[x + 1] - off[x]);
    std::vector<float> logz(out_logz ? n : 0), bpp(out_bpp ? boff[n] : 0), ea(out_expect_acc ? (size_t)n_gammas * n : 0);
    std::vector<uint8_t> st(out_structs ? (size_t)n_gammas * off[n] : 0);
    const int r = rna_mccaskill_centroid_batch(m->h[k], b.data(), off.data(), n, model, allows_short_hairpins, gammas, n_gammas,
                                               out_logz ? logz.data() : nullptr, out_bpp ? bpp.data() : nullptr, boff.data(),
                                               out_structs ? st.data() : nullptr, out_expect_acc ? ea.data() : nullptr);
    if (r != RNA_OK) return r;
    // scatter into the caller's buffers (ranges of different sequences are disjoint)
    for (uint32_t x = 0; x < n; x++) {
      const uint32_t s = mine[x], L = off[x + 1] - off[x];
      if (out_logz) out_logz[s] = logz[x];
      if (out_bpp) memcpy(out_bpp + bpp_offsets[s], bpp.data() + boff[x], sizeof(float) * (boff[x + 1] - boff[x]));
      for (uint32_t g = 0; g < n_gammas; g++) {
        if (out_structs) memcpy(out_structs + (size_t)g * total + offsets[s], st.data() + (size_t)g * off[n] + off[x], L);
        if (out_expect_acc) out_expect_acc[(size_t)g * n_seqs + s] = ea[(size_t)g * n + x];
      }
    }
    return RNA_OK;
  });
}

extern "C" int rna_multi_durbin_batch(rna_multi* m, const uint8_t* bases, const uint32_t* offsets, uint32_t n_seqs,
                                      const uint32_t* pairs, uint32_t n_pairs, float* out_probs, const uint64_t* prob_offsets) {
  if (!m || m->h.empty()) return RNA_ERR_BAD_ARG;
  if (n_pairs == 0) return RNA_OK;
  int rc = rna_validate_bases(bases, offsets, n_seqs);
  if (rc != RNA_OK) { m->err = "input validation failed"; return rc; }
  if (!pairs || !out_probs || offsets[0] != 0) { m->err = "bad arguments"; return RNA_ERR_BAD_ARG; }
  for (uint32_t p = 0; p < 2 * n_pairs; p++)
    if (pairs[p] >= n_seqs) { m->err = "pair index out of range"; return RNA_ERR_BAD_ARG; }
  const size_t nd = m->h.size();
  auto cells = [&](uint32_t p) {
    const uint64_t la = offsets[pairs[2 * p] + 1] - offsets[pairs[2 * p]], lb = offsets[pairs[2 * p + 1] + 1] - offsets[pairs[2 * p + 1]];
    return (la + 2) * (lb + 2);
  };
  std::vector<uint64_t> cost(n_pairs);
  for (uint32_t p = 0; p < n_pairs; p++) cost[p] = cells(p);
  std::vector<uint32_t> part(n_pairs);
  rna_partition_lpt(cost.data(), n_pairs, (uint32_t)nd, part.data());
  std::vector<std::vector<uint32_t>> share(nd);
  for (uint32_t p = 0; p < n_pairs; p++) share[part[p]].push_back(p);
  std::vector<uint64_t> own;
  if (!prob_offsets) {
    own.resize((size_t)n_pairs + 1);
    own[0] = 0;
    for (uint32_t p = 0; p < n_pairs; p++) own[p + 1] = own[p] + cost[p];
    prob_offsets = own.data();
  }
  return run_shares(m, share, [&](size_t k) -> int {
    const std::vector<uint32_t>& mine = share[k];
    const uint32_t n = (uint32_t)mine.size();
    std::vector<uint32_t> pr(2 * (size_t)n);
    std::vector<uint64_t> po(n + 1, 0);
    for (uint32_t x = 0; x < n; x++) {
      pr[2 * x] = pairs[2 * mine[x]]; pr[2 * x + 1] = pairs[2 * mine[x] + 1];
      po[x + 1] = po[x] + cost[mine[x]];
    }
    std::vector<float> probs(po[n]);
    const int r = rna_durbin_batch(m->h[k], bases, offsets, n_seqs, pr.data(), n, probs.data(), po.data());
    if (r != RNA_OK) return r;
    for (uint32_t x = 0; x < n; x++) memcpy(out_probs + prob_offsets[mine[x]], probs.data() + po[x], sizeof(float) * cost[mine[x]]);
    return RNA_OK;
  });
}
