// rna_queue.cpp — coalescing front end for single-sequence calls (include/rna_algos_b200.h "rna_queue").  Pure host
// code over the batched C ABI.  Leader / follower combining: a caller appends its request; if no launch is being
// prepared it becomes the leader, takes every compatible pending request, runs ONE rna_mccaskill_centroid_batch and
// distributes the results; callers that arrive meanwhile queue up and are served by the next round (the same leader
// keeps going until the queue is empty, so nobody is left waiting without a leader).
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <vector>

#include "../../include/rna_algos_b200.h"

namespace {
struct Request {
  const uint8_t* seq;
  uint32_t len;
  int contra, allows_short;
  bool want_centroid;
  float gamma;
  float* out_bpp;
  float* out_logz;
  uint8_t* out_str;
  float* out_ea;
  int rc = RNA_OK;
  bool done = false;
  bool compatible(const Request& o) const {
    return contra == o.contra && allows_short == o.allows_short && want_centroid == o.want_centroid &&
           (!want_centroid || memcmp(&gamma, &o.gamma, sizeof gamma) == 0);
  }
};
}  // namespace

struct rna_queue {
  rna_handle* h;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Request*> pending;
  bool leader_active = false;
  uint64_t requests = 0, launches = 0;
};

extern "C" int rna_queue_create(rna_handle* h, rna_queue** out) {
  if (!h || !out) return RNA_ERR_BAD_ARG;
  rna_queue* q = new rna_queue();
  q->h = h;
  *out = q;
  return RNA_OK;
}
extern "C" int rna_queue_destroy(rna_queue* q) {
  if (!q) return RNA_OK;
  {
    std::unique_lock<std::mutex> lk(q->mu);
    q->cv.wait(lk, [&] { return q->pending.empty() && !q->leader_active; });
  }
  delete q;
  return RNA_OK;
}
extern "C" int rna_queue_stats(const rna_queue* q, uint64_t* requests, uint64_t* launches) {
  if (!q) return RNA_ERR_BAD_ARG;
  if (requests) *requests = q->requests;
  if (launches) *launches = q->launches;
  return RNA_OK;
}

// one batched launch for `batch` (all compatible); fills every request's outputs and status
static void serve(rna_queue* q, std::vector<Request*>& batch) {
  const uint32_t n = (uint32_t)batch.size();
  std::vector<uint32_t> off(n + 1, 0);
  std::vector<uint64_t> boff(n + 1, 0);
  for (uint32_t x = 0; x < n; x++) {
    off[x + 1] = off[x] + batch[x]->len;
    boff[x + 1] = boff[x] + rna_bpp_len(batch[x]->len);
  }
  std::vector<uint8_t> bases(off[n]);
  for (uint32_t x = 0; x < n; x++) memcpy(bases.data() + off[x], batch[x]->seq, batch[x]->len);
  const Request& r0 = *batch[0];
  bool any_bpp = false;
  for (Request* r : batch) any_bpp = any_bpp || r->out_bpp;
  std::vector<float> logz(n), bpp(any_bpp ? boff[n] : 0), ea(r0.want_centroid ? n : 0);
  std::vector<uint8_t> st(r0.want_centroid ? off[n] : 0);
  const int rc = rna_mccaskill_centroid_batch(q->h, bases.data(), off.data(), n, r0.contra ? RNA_MODEL_CONTRA : RNA_MODEL_TURNER,
                                              r0.allows_short, r0.want_centroid ? &r0.gamma : nullptr, r0.want_centroid ? 1u : 0u,
                                              logz.data(), any_bpp ? bpp.data() : nullptr, boff.data(),
                                              r0.want_centroid ? st.data() : nullptr, r0.want_centroid ? ea.data() : nullptr);
  if (rc != RNA_OK && n > 1) {
    // a bad request must not fail its neighbours: serve them one by one
    for (Request* r : batch) { std::vector<Request*> one(1, r); serve(q, one); }
    return;
  }
  for (uint32_t x = 0; x < n; x++) {
    Request* r = batch[x];
    r->rc = rc;
    if (rc != RNA_OK) continue;
    if (r->out_logz) *r->out_logz = logz[x];
    if (r->out_bpp) memcpy(r->out_bpp, bpp.data() + boff[x], sizeof(float) * (boff[x + 1] - boff[x]));
    if (r->out_str) memcpy(r->out_str, st.data() + off[x], r->len);
    if (r->out_ea) *r->out_ea = ea[x];
  }
}

extern "C" int rna_queue_mccaskill_algo(rna_queue* q, const uint8_t* seq, uint32_t seq_len, int uses_contra_model,
                                        int allows_short_hairpins, float* out_bpp, float* out_logz, float centroid_threshold,
                                        uint8_t* out_fold_str, float* out_expect_accuracy) {
  if (!q || !seq) return RNA_ERR_BAD_ARG;
  if (seq_len == 0) return RNA_ERR_EMPTY_SEQ;
  Request me;
  me.seq = seq; me.len = seq_len; me.contra = uses_contra_model ? 1 : 0; me.allows_short = allows_short_hairpins ? 1 : 0;
  me.want_centroid = out_fold_str || out_expect_accuracy;
  me.gamma = centroid_threshold;
  me.out_bpp = out_bpp; me.out_logz = out_logz; me.out_str = out_fold_str; me.out_ea = out_expect_accuracy;
  std::unique_lock<std::mutex> lk(q->mu);
  q->pending.push_back(&me);
  q->requests++;
  for (;;) {
    if (me.done) return me.rc;
    if (!q->leader_active) break;     // become the leader
    q->cv.wait(lk);
  }
  q->leader_active = true;
  while (!q->pending.empty()) {
    // every pending request compatible with the oldest one
    std::vector<Request*> batch;
    Request* first = q->pending.front();
    for (auto it = q->pending.begin(); it != q->pending.end();) {
      if ((*it)->compatible(*first)) { batch.push_back(*it); it = q->pending.erase(it); } else ++it;
    }
    q->launches++;
    lk.unlock();
    serve(q, batch);
    lk.lock();
    for (Request* r : batch) r->done = true;
    q->cv.notify_all();
    if (me.done && !q->pending.empty()) {
      // my own request is served: hand the leadership to a waiting caller instead of serving strangers forever
      break;
    }
  }
  q->leader_active = false;
  q->cv.notify_all();
  return me.rc;
}
