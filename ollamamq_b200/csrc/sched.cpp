// Fair-share scheduler: see sched.hpp for the contract and the reference citations.
#include "sched.hpp"
#include <chrono>
#include <cstdio>
#include "../../include/ollamamq_b200.h"
#include <algorithm>
#include <cstring>
#include <new>

namespace mq {

void set_last_error(const char* fmt, ...);

Scheduler::Scheduler(int n_backends, int capacity) : backends_(n_backends < 0 ? 0 : n_backends) {
  set_capacity(capacity);
}

void Scheduler::active_insert(User* u) {
  auto it = std::lower_bound(active_.begin(), active_.end(), u,
                             [this](const User* a, const User* b) { return key_less(a, b); });
  active_.insert(it, u);
}

void Scheduler::active_erase(User* u) {
  auto it = std::lower_bound(active_.begin(), active_.end(), u,
                             [this](const User* a, const User* b) { return key_less(a, b); });
  if (it != active_.end() && *it == u) active_.erase(it);
}

uint64_t Scheduler::enqueue(const std::string& user) {
  auto it = users_.find(user);
  if (it == users_.end()) {
    it = users_.emplace(user, User{}).first;
    it->second.name = user;
  }
  User& u = it->second;
  const bool was_active = is_active(&u);
  const uint64_t id = next_task_id_++;
  u.queue.push_back(id);
  u.next_seq++;
  pending_++;
  if (!was_active) active_insert(&u);
  return id;
}

bool Scheduler::next(SchedDispatch* out) {
  // 1. eligible backends (:202-206).  Checked before anything else: with none, no state changes.
  uint64_t min_conns = UINT64_MAX;
  bool any = false;
  for (const Backend& b : backends_)
    if (b.online && b.active < (uint64_t)capacity_) {
      any = true;
      min_conns = std::min(min_conns, b.active);
    }
  if (!any || active_.empty()) return false;

  // 2. user selection (:230-240).  active_ is sorted, so the first member met is the winner of a multi-member set.
  User* target = nullptr;
  if (!vip_.empty())
    for (User* u : active_)
      if (std::find(vip_.begin(), vip_.end(), u->name) != vip_.end()) { target = u; break; }
  if (!target && !boost_.empty() && counter_ % (uint64_t)boost_mod_ == 0)
    for (User* u : active_)
      if (std::find(boost_.begin(), boost_.end(), u->name) != boost_.end()) { target = u; break; }
  if (!target) {
    if (current_idx_ >= active_.size()) current_idx_ = 0;
    target = active_[current_idx_];
    current_idx_ += 1;
  }

  // 3. pop + counter (:244-245)
  const uint64_t task = target->queue.front();
  target->queue.pop_front();
  const uint64_t seq = target->popped++;
  pending_--;
  if (target->queue.empty()) active_erase(target);
  counter_ += 1;

  // 4. least-connections, round-robin tie-break: first candidate index > last, else the first (:248-254)
  int chosen = -1, first = -1;
  for (int i = 0; i < (int)backends_.size(); ++i) {
    const Backend& b = backends_[i];
    if (!(b.online && b.active < (uint64_t)capacity_ && b.active == min_conns)) continue;
    if (first < 0) first = i;
    if ((size_t)i > last_backend_) { chosen = i; break; }
  }
  if (chosen < 0) chosen = first;
  last_backend_ = (size_t)chosen;
  backends_[chosen].active += 1;

  out->task_id = task;
  out->user_seq = seq;
  out->backend = chosen;
  out->user = target->name;
  return true;
}

void Scheduler::complete(int backend, const std::string& user, int outcome) {
  auto it = users_.find(user);
  if (it == users_.end()) {
    it = users_.emplace(user, User{}).first;  // entry().or_insert(0) (:316,:319)
    it->second.name = user;
  }
  User& u = it->second;
  if (outcome == MQ_DONE_PROCESSED) {
    // the sort key changes: reposition inside the active list
    const bool act = is_active(&u);
    if (act) active_erase(&u);
    u.processed += 1;
    if (act) active_insert(&u);
  } else if (outcome == MQ_DONE_DROPPED) {
    u.dropped += 1;
  }
  if (backend >= 0 && backend < (int)backends_.size()) {
    Backend& b = backends_[backend];
    if (b.active > 0) b.active -= 1;  // saturating_sub (:338)
    b.processed_count += 1;           // (:339)
  }
}

void Scheduler::processing(const std::string& user, int delta) {
  auto it = users_.find(user);
  if (it == users_.end()) {
    if (delta <= 0) return;  // get_mut on a missing key is a no-op (:332)
    it = users_.emplace(user, User{}).first;
    it->second.name = user;
  }
  User& u = it->second;
  if (delta > 0) u.processing += (uint64_t)delta;
  else u.processing = u.processing >= (uint64_t)(-delta) ? u.processing - (uint64_t)(-delta) : 0;
}

static void erase_name(std::vector<std::string>& v, const std::string& n) {
  v.erase(std::remove(v.begin(), v.end(), n), v.end());
}
void Scheduler::set_vip(const char* user) {
  vip_.clear();
  if (user) add_vip(user);
}
void Scheduler::set_boost(const char* user) {
  boost_.clear();
  if (user) add_boost(user);
}
void Scheduler::add_vip(const std::string& user) {
  erase_name(boost_, user);  // a user holds at most one flag (tui.rs:142-148)
  if (std::find(vip_.begin(), vip_.end(), user) == vip_.end()) vip_.push_back(user);
}
void Scheduler::add_boost(const std::string& user) {
  erase_name(vip_, user);  // tui.rs:169-175
  if (std::find(boost_.begin(), boost_.end(), user) == boost_.end()) boost_.push_back(user);
}
void Scheduler::set_online(int backend, bool online) {
  if (backend >= 0 && backend < (int)backends_.size()) backends_[backend].online = online;
}

const Scheduler::User* Scheduler::find_user(const std::string& name) const {
  auto it = users_.find(name);
  return it == users_.end() ? nullptr : &it->second;
}

std::vector<const Scheduler::User*> Scheduler::users_tui_order() const {
  std::vector<const User*> v;
  v.reserve(users_.size());
  for (const auto& kv : users_) v.push_back(&kv.second);
  std::stable_sort(v.begin(), v.end(), [](const User* a, const User* b) {
    const uint64_t aa = a->queue.size() + a->processing, ba = b->queue.size() + b->processing;
    if (aa != ba) return aa > ba;
    const uint64_t ad = a->processed + a->dropped, bd = b->processed + b->dropped;
    if (ad != bd) return ad > bd;
    return a->name < b->name;
  });
  return v;
}

}  // namespace mq

// ------------------------------------------------------------------------------------------------ C ABI
using mq::Scheduler;

extern "C" {

mq_sched* mq_sched_new(int32_t n_backends, int32_t capacity) {
  if (n_backends < 1) {
    mq::set_last_error("mq_sched_new: n_backends must be >= 1");
    return nullptr;
  }
  return new (std::nothrow) mq_sched(n_backends, capacity);
}
void mq_sched_free(mq_sched* s) { delete s; }

int mq_sched_enqueue(mq_sched* s, const char* user, uint64_t* task_id_out) {
  if (!s) return MQ_ERR_INVAL;
  const char* u = user ? user : "anonymous";  // :364-368
  if (strlen(u) >= MQ_USER_MAX) {
    mq::set_last_error("user id longer than %d bytes", MQ_USER_MAX - 1);
    return MQ_ERR_INVAL;
  }
  const uint64_t id = s->s.enqueue(u);
  if (task_id_out) *task_id_out = id;
  return MQ_OK;
}

int mq_sched_next(mq_sched* s, mq_dispatch* out) {
  if (!s || !out) return MQ_ERR_INVAL;
  mq::SchedDispatch d;
  if (!s->s.next(&d)) return 0;
  out->task_id = d.task_id;
  out->user_seq = d.user_seq;
  out->backend = d.backend;
  out->reserved = 0;
  strncpy(out->user, d.user.c_str(), MQ_USER_MAX - 1);
  out->user[MQ_USER_MAX - 1] = 0;
  return 1;
}

int mq_sched_complete(mq_sched* s, int32_t backend, const char* user, int32_t outcome) {
  if (!s || !user || backend < 0 || backend >= s->s.n_backends()) return MQ_ERR_INVAL;
  s->s.complete(backend, user, outcome);
  return MQ_OK;
}
int mq_sched_processing(mq_sched* s, const char* user, int32_t delta) {
  if (!s || !user) return MQ_ERR_INVAL;
  s->s.processing(user, delta);
  return MQ_OK;
}
int mq_sched_set_vip(mq_sched* s, const char* user) {
  if (!s) return MQ_ERR_INVAL;
  s->s.set_vip(user);
  return MQ_OK;
}
int mq_sched_set_boost(mq_sched* s, const char* user) {
  if (!s) return MQ_ERR_INVAL;
  s->s.set_boost(user);
  return MQ_OK;
}
int mq_sched_add_vip(mq_sched* s, const char* user) {
  if (!s || !user) return MQ_ERR_INVAL;
  s->s.add_vip(user);
  return MQ_OK;
}
int mq_sched_add_boost(mq_sched* s, const char* user) {
  if (!s || !user) return MQ_ERR_INVAL;
  s->s.add_boost(user);
  return MQ_OK;
}
int mq_sched_set_online(mq_sched* s, int32_t backend, int32_t online) {
  if (!s || backend < 0 || backend >= s->s.n_backends()) return MQ_ERR_INVAL;
  s->s.set_online(backend, online != 0);
  return MQ_OK;
}
int mq_sched_set_capacity(mq_sched* s, int32_t capacity) {
  if (!s || capacity < 1) return MQ_ERR_INVAL;
  s->s.set_capacity(capacity);
  return MQ_OK;
}
int mq_sched_set_boost_mod(mq_sched* s, int32_t mod) {
  if (!s || mod < 1) return MQ_ERR_INVAL;
  s->s.set_boost_mod(mod);
  return MQ_OK;
}
int mq_sched_user_stats(mq_sched* s, const char* user, mq_user_stats* out) {
  if (!s || !user || !out) return MQ_ERR_INVAL;
  const Scheduler::User* u = s->s.find_user(user);
  if (!u) return MQ_ERR_NOENT;
  out->queued = u->queue.size();
  out->processing = u->processing;
  out->processed = u->processed;
  out->dropped = u->dropped;
  return MQ_OK;
}
int mq_sched_backend_stats(mq_sched* s, int32_t backend, mq_backend_stats* out) {
  if (!s || !out) return MQ_ERR_INVAL;
  const Scheduler::Backend* b = s->s.backend(backend);
  if (!b) return MQ_ERR_INVAL;
  out->active_requests = b->active;
  out->processed_count = b->processed_count;
  out->is_online = b->online ? 1 : 0;
  out->reserved = 0;
  return MQ_OK;
}
int32_t mq_sched_user_count(mq_sched* s) { return s ? s->s.user_count() : 0; }
int mq_sched_user_name(mq_sched* s, int32_t index, char* out, size_t cap) {
  if (!s || !out || cap == 0) return MQ_ERR_INVAL;
  auto v = s->s.users_tui_order();
  if (index < 0 || index >= (int)v.size()) return MQ_ERR_NOENT;
  strncpy(out, v[index]->name.c_str(), cap - 1);
  out[cap - 1] = 0;
  return MQ_OK;
}
uint64_t mq_sched_counter(mq_sched* s) { return s ? s->s.counter() : 0; }

// Decision micro-benchmark ("decisions/s of the C++ scheduler at U = 64 / 256"): U users x R requests
// enqueued user-major, B backends of the given capacity, unit service time, event model of SURVEY.md 3.2 (drain to
// quiescence, complete the in-flight request on the lowest backend index, repeat) - the same driver loop as
// oracle/dispatch_oracle.c:orc_bench, so the two dispatch counts must agree.
int mq_debug_sched_bench(int32_t n_users, int32_t reqs_per_user, int32_t n_backends, int32_t capacity,
                         uint64_t* dispatches_out, double* seconds_out) {
  if (n_users < 1 || reqs_per_user < 1 || n_backends < 1 || capacity < 1 || !dispatches_out || !seconds_out) return MQ_ERR_INVAL;
  mq::Scheduler sc(n_backends, capacity);
  char name[16];
  for (int u = 0; u < n_users; ++u) {
    snprintf(name, sizeof name, "user%03d", u);
    for (int r = 0; r < reqs_per_user; ++r) sc.enqueue(name);
  }
  std::vector<std::pair<int, std::string>> inflight;
  uint64_t total = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    mq::SchedDispatch d;
    while (sc.next(&d)) { inflight.emplace_back(d.backend, d.user); ++total; }
    if (inflight.empty()) break;
    size_t best = 0;
    for (size_t i = 1; i < inflight.size(); ++i) if (inflight[i].first < inflight[best].first) best = i;
    sc.complete(inflight[best].first, inflight[best].second, 0);
    inflight[best] = inflight.back();
    inflight.pop_back();
  }
  *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  *dispatches_out = total;
  return MQ_OK;
}

}  // extern "C"
