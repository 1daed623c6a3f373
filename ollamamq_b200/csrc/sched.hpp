// Fair-share scheduler state machine (product implementation).
//
// Behaviour contract: /root/reference/src/dispatcher.rs:195-262 (run_worker loop body) and :314-341
// (executor epilogue).  The reference rebuilds and sorts the active-user list on every dispatch
// (:216-228); this implementation keeps the list sorted incrementally and repositions one user when its
// processed count changes, so a dispatch is O(log U + U-memmove) with no allocation.  The observable
// sequence of (user, backend) decisions must stay bit-identical to the oracle (oracle/dispatch_oracle.c).
#pragma once
#include <cstdint>
#include <deque>
#include <algorithm>
#include <map>
#include <string>
#include <vector>

namespace mq {

struct SchedDispatch {
  uint64_t task_id;
  uint64_t user_seq;
  int backend;
  std::string user;
};

class Scheduler {
 public:
  Scheduler(int n_backends, int capacity);

  uint64_t enqueue(const std::string& user);
  bool next(SchedDispatch* out);
  void complete(int backend, const std::string& user, int outcome);
  void processing(const std::string& user, int delta);

  void set_vip(const char* user);    // reference semantics: ONE slot (nullptr clears every VIP)
  void set_boost(const char* user);
  // EXTENSION (BASELINE config 3, "2 VIP + 4 Boost"; no reference semantics): sets of users.  The
  // winner is the first member in the reference sort order (:224-228); with <= 1 member this is the reference.
  void add_vip(const std::string& user);
  void add_boost(const std::string& user);
  void set_online(int backend, bool online);
  void set_capacity(int c) { capacity_ = c < 1 ? 1 : c; }
  void set_boost_mod(int m) { boost_mod_ = m < 1 ? 1 : m; }

  struct User {
    std::string name;
    std::deque<uint64_t> queue;  // task ids, FIFO (:244, :399-402)
    uint64_t next_seq = 0;       // tasks ever enqueued
    uint64_t popped = 0;         // tasks ever dispatched
    uint64_t processing = 0, processed = 0, dropped = 0;
  };
  struct Backend {
    uint64_t active = 0, processed_count = 0;
    bool online = true;  // :75
  };

  const User* find_user(const std::string& name) const;
  const Backend* backend(int i) const { return i >= 0 && i < (int)backends_.size() ? &backends_[i] : nullptr; }
  int n_backends() const { return (int)backends_.size(); }
  int user_count() const { return (int)users_.size(); }
  std::vector<const User*> users_tui_order() const;  // tui.rs:70-80
  uint64_t counter() const { return counter_; }
  const std::vector<std::string>& vips() const { return vip_; }
  const std::vector<std::string>& boosts() const { return boost_; }
  uint64_t pending() const { return pending_; }

 private:
  // sort key of the reference comparator (:224-228): processed count ascending, then name (byte order)
  bool key_less(const User* a, const User* b) const {
    if (a->processed != b->processed) return a->processed < b->processed;
    return a->name < b->name;
  }
  void active_insert(User* u);
  void active_erase(User* u);
  bool is_active(const User* u) const { return !u->queue.empty(); }

  std::map<std::string, User> users_;  // entries are never removed (:399-402; TUI list only grows)
  std::vector<User*> active_;          // users with a non-empty queue, sorted by key_less
  std::vector<Backend> backends_;
  std::vector<std::string> vip_, boost_;  // sets (tiny): the reference has one Option<String> each (:57-58)
  uint64_t counter_ = 0;       // global_counter (:89)
  size_t current_idx_ = 0;     // run_worker local (:169)
  size_t last_backend_ = 0;    // last_backend_idx (:93)
  int capacity_ = 1;           // :204
  int boost_mod_ = 2;          // :233
  uint64_t next_task_id_ = 1;
  uint64_t pending_ = 0;       // queued tasks over all users
};

}  // namespace mq

// C handle (include/ollamamq_b200.h)
struct mq_sched {
  mq::Scheduler s;
  mq_sched(int n, int c) : s(n, c) {}
};
