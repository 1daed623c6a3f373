/* dispatch_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into libollamamq_b200.so).
 *
 * Plain-C restatement of ollamaMQ's scheduling decision, written to follow the reference line by line
 * rather than to be fast:
 *     /root/reference/src/dispatcher.rs:67-96    initial state (all online, last_backend_idx 0, counter 0)
 *     /root/reference/src/dispatcher.rs:164-169  loop-local current_idx
 *     /root/reference/src/dispatcher.rs:195-262  one iteration of the run_worker loop
 *     /root/reference/src/dispatcher.rs:314-341  executor epilogue (processed/dropped, backend release)
 *     /root/reference/src/dispatcher.rs:364-405  enqueue ("anonymous" default)
 *
 * PARITY UNPINNED: the reference ships no golden vectors, unit tests or fixtures for dispatch order
 *  and cannot be compiled here (no Rust toolchain), so this oracle is pinned only
 * against the three traces SURVEY.md 3.2 derives by hand from the source (tests/golden/dispatch_seed.json)
 * and against an independent Python restatement (oracle/dispatch_oracle.py).
 *
 * Generalisations, all defaulting to the reference's constants: capacity (1, :204), boost_mod (2, :233).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NAME_MAX 256

typedef struct {
  char name[ORC_NAME_MAX];
  long queued;    /* VecDeque length (:216-219) */
  long popped;    /* tasks dispatched so far = seq of the next pop */
  long processed; /* processed_counts[u] */
  long dropped;   /* dropped_counts[u] */
} orc_user;

typedef struct {
  long active_requests, processed_count;
  int is_online;
} orc_backend;

typedef struct orc {
  orc_user* users;
  int n_users, cap_users;
  orc_backend* backends;
  int n_backends;
  /* EXTENSION (BASELINE config 3): up to ORC_SET_MAX VIP / Boost users; the reference has one slot each (:57-58) */
  char vip[8][ORC_NAME_MAX], boost[8][ORC_NAME_MAX];
  int n_vip, n_boost;
  unsigned long global_counter; /* :89 */
  size_t current_idx;           /* :169 */
  size_t last_backend_idx;      /* :93 */
  int capacity, boost_mod;
} orc;

orc* orc_new(int n_backends, int capacity, int boost_mod) {
  orc* o = (orc*)calloc(1, sizeof(orc));
  o->backends = (orc_backend*)calloc((size_t)n_backends, sizeof(orc_backend));
  o->n_backends = n_backends;
  for (int i = 0; i < n_backends; ++i) o->backends[i].is_online = 1; /* :75 */
  o->capacity = capacity;
  o->boost_mod = boost_mod;
  return o;
}
void orc_free(orc* o) {
  if (!o) return;
  free(o->users);
  free(o->backends);
  free(o);
}

static orc_user* find_user(orc* o, const char* name, int create) {
  for (int i = 0; i < o->n_users; ++i)
    if (strcmp(o->users[i].name, name) == 0) return &o->users[i];
  if (!create) return NULL;
  if (o->n_users == o->cap_users) {
    o->cap_users = o->cap_users ? o->cap_users * 2 : 16;
    o->users = (orc_user*)realloc(o->users, (size_t)o->cap_users * sizeof(orc_user));
  }
  orc_user* u = &o->users[o->n_users++];
  memset(u, 0, sizeof(*u));
  strncpy(u->name, name, ORC_NAME_MAX - 1);
  return u;
}

void orc_enqueue(orc* o, const char* user) { /* :364-368, :397-403 */
  find_user(o, user ? user : "anonymous", 1)->queued += 1;
}
static int set_find(char (*set)[ORC_NAME_MAX], int n, const char* u) {
  for (int i = 0; i < n; ++i)
    if (strcmp(set[i], u) == 0) return i;
  return -1;
}
static void set_del(char (*set)[ORC_NAME_MAX], int* n, const char* u) {
  int i = set_find(set, *n, u);
  if (i < 0) return;
  for (; i + 1 < *n; ++i) memcpy(set[i], set[i + 1], ORC_NAME_MAX);
  *n -= 1;
}
void orc_add_vip(orc* o, const char* u) {
  set_del(o->boost, &o->n_boost, u); /* tui.rs:142-148 */
  if (set_find(o->vip, o->n_vip, u) < 0 && o->n_vip < 8) strncpy(o->vip[o->n_vip++], u, ORC_NAME_MAX - 1);
}
void orc_add_boost(orc* o, const char* u) {
  set_del(o->vip, &o->n_vip, u); /* tui.rs:169-175 */
  if (set_find(o->boost, o->n_boost, u) < 0 && o->n_boost < 8) strncpy(o->boost[o->n_boost++], u, ORC_NAME_MAX - 1);
}
void orc_set_vip(orc* o, const char* u) {
  o->n_vip = 0;
  memset(o->vip, 0, sizeof(o->vip));
  if (u) orc_add_vip(o, u);
}
void orc_set_boost(orc* o, const char* u) {
  o->n_boost = 0;
  memset(o->boost, 0, sizeof(o->boost));
  if (u) orc_add_boost(o, u);
}
void orc_set_online(orc* o, int b, int online) { o->backends[b].is_online = online; }
void orc_set_capacity(orc* o, int c) { o->capacity = c; }

static orc* g_sort_ctx; /* qsort has no context argument */
static int cmp_users(const void* pa, const void* pb) { /* :224-228 */
  const orc_user* a = &g_sort_ctx->users[*(const int*)pa];
  const orc_user* b = &g_sort_ctx->users[*(const int*)pb];
  if (a->processed != b->processed) return a->processed < b->processed ? -1 : 1;
  return strcmp(a->name, b->name); /* String Ord = byte-wise */
}

/* One iteration of the loop body.  Returns 1 when a task was dispatched. */
int orc_next(orc* o, char* user_out, int cap, long* seq_out, int* backend_out) {
  /* :202-206 online_indices */
  int* online = (int*)malloc(sizeof(int) * (size_t)(o->n_backends + 1));
  int n_online = 0;
  for (int i = 0; i < o->n_backends; ++i)
    if (o->backends[i].is_online && o->backends[i].active_requests < o->capacity) online[n_online++] = i;
  if (n_online == 0) { free(online); return 0; } /* :208-209 */

  /* :216-219 active_users */
  int* active = (int*)malloc(sizeof(int) * (size_t)(o->n_users + 1));
  int n_active = 0;
  for (int i = 0; i < o->n_users; ++i)
    if (o->users[i].queued > 0) active[n_active++] = i;
  if (n_active == 0) { free(online); free(active); return 0; } /* :221-222 */
  g_sort_ctx = o;
  qsort(active, (size_t)n_active, sizeof(int), cmp_users);

  int target = -1;
  /* :230 — with one VIP exactly `active_users.contains(v)`; with several, the first in sorted order wins */
  for (int i = 0; i < n_active && target < 0; ++i)
    if (set_find(o->vip, o->n_vip, o->users[active[i]].name) >= 0) target = active[i];
  if (target < 0 && o->n_boost > 0 && o->global_counter % (unsigned long)o->boost_mod == 0) /* :231-235 */
    for (int i = 0; i < n_active && target < 0; ++i)
      if (set_find(o->boost, o->n_boost, o->users[active[i]].name) >= 0) target = active[i];
  if (target < 0) { /* :236-240 */
    if (o->current_idx >= (size_t)n_active) o->current_idx = 0;
    target = active[o->current_idx];
    o->current_idx += 1;
  }

  orc_user* u = &o->users[target];
  u->queued -= 1; /* pop_front :244 */
  *seq_out = u->popped++;
  o->global_counter += 1; /* :245 */

  /* :248-254 */
  long min_conns = -1;
  for (int k = 0; k < n_online; ++k)
    if (min_conns < 0 || o->backends[online[k]].active_requests < min_conns)
      min_conns = o->backends[online[k]].active_requests;
  int* cand = (int*)malloc(sizeof(int) * (size_t)n_online);
  int n_cand = 0;
  for (int k = 0; k < n_online; ++k)
    if (o->backends[online[k]].active_requests == min_conns) cand[n_cand++] = online[k];
  int pos = 0;
  for (int k = 0; k < n_cand; ++k)
    if ((size_t)cand[k] > o->last_backend_idx) { pos = k; break; }
  const int sel = cand[pos];
  o->last_backend_idx = (size_t)sel;
  o->backends[sel].active_requests += 1;

  strncpy(user_out, u->name, (size_t)cap - 1);
  user_out[cap - 1] = 0;
  *backend_out = sel;
  free(online); free(active); free(cand);
  return 1;
}

/* outcome: 0 processed (:314-316), 1 dropped (:280,:318-319,:326-327), 2 uncounted (:299) */
void orc_complete(orc* o, int backend, const char* user, int outcome) {
  orc_user* u = find_user(o, user, 1);
  if (outcome == 0) u->processed += 1;
  else if (outcome == 1) u->dropped += 1;
  orc_backend* b = &o->backends[backend];
  if (b->active_requests > 0) b->active_requests -= 1; /* saturating_sub :338 */
  b->processed_count += 1;                             /* :339 */
}

long orc_user_processed(orc* o, const char* user) { orc_user* u = find_user(o, user, 0); return u ? u->processed : -1; }
long orc_user_dropped(orc* o, const char* user) { orc_user* u = find_user(o, user, 0); return u ? u->dropped : -1; }
long orc_user_queued(orc* o, const char* user) { orc_user* u = find_user(o, user, 0); return u ? u->queued : -1; }
long orc_backend_active(orc* o, int b) { return o->backends[b].active_requests; }
long orc_backend_processed(orc* o, int b) { return o->backends[b].processed_count; }

/* Bulk decision benchmark for bench.py's cpu_baseline of the dispatch row: U users x R requests, B
 * backends, unit service time, event model of SURVEY.md 3.2.  Returns the number of dispatches. */
long orc_bench(int n_users, int reqs_per_user, int n_backends, int capacity) {
  orc* o = orc_new(n_backends, capacity, 2);
  char name[32];
  for (int u = 0; u < n_users; ++u) {
    memset(name, 0, sizeof(name));
    name[0] = 'u'; name[1] = 's'; name[2] = 'e'; name[3] = 'r';
    name[4] = (char)('0' + (u / 100) % 10); name[5] = (char)('0' + (u / 10) % 10); name[6] = (char)('0' + u % 10);
    for (int r = 0; r < reqs_per_user; ++r) orc_enqueue(o, name);
  }
  long total = 0, seq;
  int be;
  char who[ORC_NAME_MAX];
  char (*inflight)[ORC_NAME_MAX] = (char (*)[ORC_NAME_MAX])calloc((size_t)n_backends * (size_t)capacity, ORC_NAME_MAX);
  int* inflight_b = (int*)calloc((size_t)n_backends * (size_t)capacity, sizeof(int));
  int n_inflight = 0;
  for (;;) {
    while (orc_next(o, who, ORC_NAME_MAX, &seq, &be)) {
      memcpy(inflight[n_inflight], who, ORC_NAME_MAX);
      inflight_b[n_inflight++] = be;
      ++total;
    }
    if (n_inflight == 0) break;
    /* complete the in-flight request on the lowest backend index, then reschedule */
    int best = 0;
    for (int i = 1; i < n_inflight; ++i) if (inflight_b[i] < inflight_b[best]) best = i;
    orc_complete(o, inflight_b[best], inflight[best], 0);
    --n_inflight;
    memcpy(inflight[best], inflight[n_inflight], ORC_NAME_MAX);
    inflight_b[best] = inflight_b[n_inflight];
  }
  free(inflight); free(inflight_b);
  orc_free(o);
  return total;
}
