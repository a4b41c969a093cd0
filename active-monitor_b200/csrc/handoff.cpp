// handoff.cpp — bounded FIFO between the tick loop and the workflow workers
// (the step after the sweep).  Host only.
//
// The reference arms one time.AfterFunc per HealthCheck (hcc.go:751) and every
// fire runs createSubmitWorkflow on its own goroutine; concurrency is bounded
// only for Reconcile (MaxConcurrentReconciles, hcc.go:298).  With one sweep per
// second producing the whole due list at once, the fan-out becomes a queue:
// one publisher (the ticker), MaxParallel poppers.  A tick's entries are
// enqueued all-or-nothing so that a worker never sees half a tick.
#include <cstdlib>
#include <mutex>
#include <new>

#include "../../include/amsweep.h"

static_assert(sizeof(am_work_item_t) == 24, "am_work_item_t is part of the ABI");

struct am_handoff {
  std::mutex mu;
  am_work_item_t* ring = nullptr;
  uint64_t cap = 0;        // slots
  uint64_t head = 0;       // next slot to pop   (monotonic, modulo cap on access)
  uint64_t tail = 0;       // next slot to fill  (monotonic)
  uint64_t rejected = 0;   // publish calls refused for lack of room
};

extern "C" int am_handoff_create(am_handoff_t** out, uint64_t capacity) {
  if (!out || capacity == 0 || capacity > (1ull << 40)) return AM_E_INVAL;
  am_handoff* h = new (std::nothrow) am_handoff;
  if (!h) return AM_E_NOMEM;
  h->ring = static_cast<am_work_item_t*>(std::malloc(capacity * sizeof(am_work_item_t)));
  if (!h->ring) {
    delete h;
    return AM_E_NOMEM;
  }
  h->cap = capacity;
  *out = h;
  return AM_OK;
}

extern "C" void am_handoff_destroy(am_handoff_t* h) {
  if (!h) return;
  std::free(h->ring);
  delete h;
}

extern "C" int am_handoff_publish(am_handoff_t* h, int64_t unix_sec, uint64_t n, const uint64_t* idx,
                                  const uint32_t* action, uint32_t action_mask, uint64_t* n_out) {
  if (!h || (n && (!idx || !action))) return AM_E_INVAL;
  // count outside the lock: the caller's arrays are not shared
  uint64_t want = 0;
  for (uint64_t i = 0; i < n; ++i) want += (action[i] & action_mask) != 0;
  std::lock_guard<std::mutex> g(h->mu);
  if (n_out) *n_out = want;
  if (want > h->cap - (h->tail - h->head)) {
    ++h->rejected;
    return AM_E_NOSPACE;
  }
  uint64_t t = h->tail;
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t a = action[i] & action_mask;
    if (!a) continue;
    am_work_item_t& w = h->ring[t % h->cap];
    w.idx = idx[i];
    w.unix_sec = unix_sec;
    w.action = a;
    w.reserved = 0;
    ++t;
  }
  h->tail = t;
  return AM_OK;
}

extern "C" int am_handoff_pop(am_handoff_t* h, uint64_t max, am_work_item_t* out, uint64_t* n_out) {
  if (!h || !n_out || (max && !out)) return AM_E_INVAL;
  std::lock_guard<std::mutex> g(h->mu);
  uint64_t k = h->tail - h->head;
  if (k > max) k = max;
  for (uint64_t i = 0; i < k; ++i) out[i] = h->ring[(h->head + i) % h->cap];
  h->head += k;
  *n_out = k;
  return AM_OK;
}

extern "C" int am_handoff_stats(am_handoff_t* h, uint64_t* pending, uint64_t* published,
                                uint64_t* popped, uint64_t* rejected_batches) {
  if (!h) return AM_E_INVAL;
  std::lock_guard<std::mutex> g(h->mu);
  if (pending) *pending = h->tail - h->head;
  if (published) *published = h->tail;
  if (popped) *popped = h->head;
  if (rejected_batches) *rejected_batches = h->rejected;
  return AM_OK;
}
