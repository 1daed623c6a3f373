// HTTP/1.1 ingress in front of the dispatcher — SURVEY.md 8(f) rank 1 ("next" row).
//
// Restates, over plain sockets, what the reference does with axum:
//   /root/reference/src/main.rs:89-121       route table: GET /health -> "OK" un-queued; 20 explicit routes, any
//                                            method, all queued through proxy_handler; optional fallback
//                                            (--allow-all-routes); 1 GiB body limit
//   /root/reference/src/dispatcher.rs:354-428  proxy_handler: X-User-ID (default "anonymous"), 403 "IP blocked" /
//                                            "User blocked", enqueue, first ResponsePart decides the status line,
//                                            streamed body, 500 "Backend error: ..." / "Worker failed to respond"
//   /root/reference/src/tui.rs:126-237       the dashboard's control keys (VIP / Boost / block / unblock), exposed
//                                            here as loopback-only POST /admin/* so a headless box can be driven
//
// One epoll loop (edge-free, level-triggered) owns every socket: non-blocking accept / recv / send, a small state
// machine per connection (read head -> read body -> in flight -> next request).  The worker's callbacks never touch a
// socket: they frame bytes (status line, chunked encoding) into the request's own buffer under its own lock and ring
// the loop's eventfd; the loop moves the bytes.  That buffer is bounded, which plays the role of the reference's
// mpsc::channel(32) (:385): a client that stops reading counts as gone instead of stalling a GPU worker.
// Round 1 ran one detached thread per connection; this loop is joined on stop and holds 1000+ idle connections.
#include "../../include/ollamamq_b200.h"
#include <arpa/inet.h>
#include <atomic>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <memory>
#include <mutex>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <string>
#include <sys/epoll.h>
#include <sys/eventfd.h>
#include <sys/socket.h>
#include <thread>
#include <unistd.h>
#include <unordered_map>
#include <vector>

namespace mq {
void set_last_error(const char* fmt, ...);
}

namespace {

const char* kRoutes[] = {"/",           "/api/generate", "/api/chat",           "/api/embed",      "/api/embeddings",
                         "/api/tags",   "/api/show",     "/api/create",         "/api/copy",       "/api/delete",
                         "/api/pull",   "/api/push",     "/api/ps",             "/api/version",    "/v1/chat/completions",
                         "/v1/completions", "/v1/embeddings", "/v1/models"};
const char* kPrefixRoutes[] = {"/api/blobs/", "/v1/models/"};  // "/api/blobs/{digest}", "/v1/models/{model}"
constexpr size_t kBodyLimit = 1024ull * 1024 * 1024;           // main.rs:120
constexpr size_t kHeadLimit = 64 * 1024;
constexpr size_t kQueueLimit = 4u << 20;                       // bytes buffered per request before the client counts as gone

bool route_known(const std::string& path) {
  for (const char* r : kRoutes)
    if (path == r) return true;
  for (const char* r : kPrefixRoutes) {
    const size_t n = strlen(r);
    if (path.size() > n && path.compare(0, n, r) == 0 && path.find('/', n) == std::string::npos) return true;
  }
  return false;
}

int endpoint_of(const std::string& path) {
  if (path == "/api/generate") return MQ_EP_API_GENERATE;
  if (path == "/api/chat") return MQ_EP_API_CHAT;
  if (path == "/v1/chat/completions") return MQ_EP_V1_CHAT;
  if (path == "/v1/completions") return MQ_EP_V1_COMPLETIONS;
  if (path == "/api/embed" || path == "/api/embeddings" || path == "/v1/embeddings") return MQ_EP_EMBED;
  return MQ_EP_OTHER;
}

const char* reason_of(int s) {
  switch (s) {
    case 200: return "OK";
    case 400: return "Bad Request";
    case 403: return "Forbidden";
    case 404: return "Not Found";
    case 405: return "Method Not Allowed";
    case 413: return "Payload Too Large";
    case 500: return "Internal Server Error";
    case 501: return "Not Implemented";
    default: return "Status";
  }
}

std::string simple_response(int status, const char* ctype, const std::string& body, bool keep) {
  char h[256];
  const int n = snprintf(h, sizeof(h), "HTTP/1.1 %d %s\r\nContent-Type: %s\r\nContent-Length: %zu\r\nConnection: %s\r\n\r\n",
                         status, reason_of(status), ctype, body.size(), keep ? "keep-alive" : "close");
  return std::string(h, (size_t)n) + body;
}

}  // namespace

struct mq_http_server;

namespace {

// One in-flight request.  Shared by the loop (through its connection) and by the dispatcher's callbacks; freed by
// whoever drops the second reference.
struct Pending {
  std::mutex mu;
  std::atomic<int> refs{2};
  mq_http_server* s = nullptr;  // valid while !orphan
  uint64_t conn_id = 0;
  bool keep = true;
  std::string out;              // framed bytes for the socket
  bool head_sent = false, done = false, overflow = false;
  bool orphan = false;          // connection (or server) is gone: drop everything, report "client gone"
  bool signaled = false;        // already on the loop's ready list
};
void pending_unref(Pending* p) {
  if (p->refs.fetch_sub(1) == 1) delete p;
}

struct Conn {
  int fd = -1;
  uint64_t id = 0;
  std::string ip;
  std::string in, out;
  enum { READING, INFLIGHT } state = READING;
  Pending* cur = nullptr;
  uint64_t task = 0;
  bool close_after_flush = false;
  bool want_out = false;  // EPOLLOUT registered
};

struct Request {
  std::string method, path, user;
  size_t body_off = 0, body_len = 0;  // into Conn::in
  bool keep_alive = true;
  bool has_user = false;
};

}  // namespace

struct mq_http_server {
  mq_dispatcher* d = nullptr;
  int listen_fd = -1, ep = -1, evfd = -1;
  int port = 0;
  bool allow_all = false;
  std::atomic<bool> stop{false};
  std::thread loop;
  std::mutex mu;                     // guards `ready`
  std::vector<uint64_t> ready;       // connections whose request has new bytes / finished
  std::unordered_map<uint64_t, std::unique_ptr<Conn>> conns;  // loop thread only
  uint64_t next_id = 16;             // ids below 16 are reserved epoll tags (listen = 1, eventfd = 2)
};

namespace {

void ring(mq_http_server* s, Pending* p) {  // caller holds p->mu and has checked !p->orphan
  if (p->signaled) return;
  p->signaled = true;
  {
    std::lock_guard<std::mutex> g(s->mu);
    s->ready.push_back(p->conn_id);
  }
  const uint64_t one = 1;
  if (::write(s->evfd, &one, sizeof(one)) < 0) { /* counter saturated: the loop is awake anyway */ }
}

// ---- dispatcher callbacks (worker / scheduler threads): frame into the request's buffer, ring the loop
void cb_status(void* u, int32_t status, const char* ctype) {
  auto* p = (Pending*)u;
  std::lock_guard<std::mutex> g(p->mu);
  if (p->orphan || p->head_sent) return;
  char h[320];
  const int n = snprintf(h, sizeof(h), "HTTP/1.1 %d %s\r\nContent-Type: %.128s\r\nTransfer-Encoding: chunked\r\nConnection: %s\r\n\r\n",
                         status, reason_of(status), ctype ? ctype : "application/octet-stream", p->keep ? "keep-alive" : "close");
  p->out.append(h, (size_t)n);
  p->head_sent = true;
  ring(p->s, p);
}
int32_t cb_chunk(void* u, const uint8_t* data, size_t len) {
  auto* p = (Pending*)u;
  std::lock_guard<std::mutex> g(p->mu);
  if (p->orphan || p->overflow) return 1;
  if (len == 0) return 0;
  if (p->out.size() + len > kQueueLimit) {  // client is not reading: same outcome as a failed send (:305-308)
    p->overflow = true;
    ring(p->s, p);
    return 1;
  }
  char h[24];
  const int n = snprintf(h, sizeof(h), "%zx\r\n", len);
  p->out.append(h, (size_t)n);
  p->out.append((const char*)data, len);
  p->out.append("\r\n", 2);
  ring(p->s, p);
  return 0;
}
void cb_done(void* u, int32_t rc, const char* msg) {
  auto* p = (Pending*)u;
  {
    std::lock_guard<std::mutex> g(p->mu);
    if (!p->orphan) {
      if (p->head_sent) {
        p->out.append("0\r\n\r\n", 5);
      } else {
        // no Status part ever arrived: Error -> "Backend error: ..." (:423-425), otherwise "Worker failed to respond" (:427)
        const std::string err = msg ? msg : "";
        const std::string body = rc == MQ_ERR_BLOCKED || err.empty() ? "Worker failed to respond"
                                 : (err.rfind("Backend error:", 0) == 0 ? err : "Backend error: " + err);
        p->out += simple_response(500, "text/plain; charset=utf-8", body, p->keep);
      }
      p->done = true;
      ring(p->s, p);
    }
  }
  pending_unref(p);
}

// ---- loop-side helpers
void epoll_mod(mq_http_server* s, Conn* c) {
  epoll_event ev;
  memset(&ev, 0, sizeof(ev));
  // while a request is in flight only hang-ups matter (further pipelined bytes stay in the kernel's buffer)
  ev.events = EPOLLRDHUP | (c->state == Conn::READING && !c->close_after_flush ? EPOLLIN : 0) | (c->want_out ? EPOLLOUT : 0);
  ev.data.u64 = c->id;
  epoll_ctl(s->ep, EPOLL_CTL_MOD, c->fd, &ev);
}

void close_conn(mq_http_server* s, Conn* c) {
  if (c->cur) {
    Pending* p = c->cur;
    bool finished;
    {
      std::lock_guard<std::mutex> g(p->mu);
      p->orphan = true;
      finished = p->done;
    }
    if (!finished) mq_dispatcher_client_gone(s->d, c->task);  // responder closed (:278, :305-308)
    pending_unref(p);
    c->cur = nullptr;
  }
  epoll_ctl(s->ep, EPOLL_CTL_DEL, c->fd, nullptr);
  ::close(c->fd);
  s->conns.erase(c->id);  // frees c
}

// returns false when the connection was closed
bool flush_out(mq_http_server* s, Conn* c) {
  while (!c->out.empty()) {
    const ssize_t w = ::send(c->fd, c->out.data(), c->out.size(), MSG_NOSIGNAL);
    if (w > 0) { c->out.erase(0, (size_t)w); continue; }
    if (w < 0 && errno == EINTR) continue;
    if (w < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) break;
    close_conn(s, c);
    return false;
  }
  const bool want = !c->out.empty();
  if (c->out.empty() && c->close_after_flush && !c->cur) { close_conn(s, c); return false; }
  if (want != c->want_out) { c->want_out = want; epoll_mod(s, c); }
  return true;
}

bool parse_head(const std::string& in, size_t hdr_end, Request* rq, size_t* content_len, int* err) {
  const std::string head = in.substr(0, hdr_end);
  const size_t line_end = head.find("\r\n");
  const std::string rl = head.substr(0, line_end);
  const size_t s1 = rl.find(' '), s2 = rl.rfind(' ');
  if (s1 == std::string::npos || s2 <= s1) { *err = 400; return false; }
  rq->method = rl.substr(0, s1);
  const std::string target = rl.substr(s1 + 1, s2 - s1 - 1);
  const size_t qm = target.find('?');  // only uri.path() is used; the query string is dropped (:362)
  rq->path = qm == std::string::npos ? target : target.substr(0, qm);
  rq->keep_alive = rl.substr(s2 + 1) != "HTTP/1.0";
  *content_len = 0;
  size_t pos = line_end == std::string::npos ? head.size() : line_end + 2;
  while (pos < head.size()) {
    size_t e = head.find("\r\n", pos);
    if (e == std::string::npos) e = head.size();
    const std::string line = head.substr(pos, e - pos);
    pos = e + 2;
    const size_t c = line.find(':');
    if (c == std::string::npos) continue;
    std::string k = line.substr(0, c), v = line.substr(c + 1);
    for (auto& ch : k) ch = (char)tolower((unsigned char)ch);
    while (!v.empty() && (v.front() == ' ' || v.front() == '\t')) v.erase(v.begin());
    while (!v.empty() && (v.back() == ' ' || v.back() == '\t')) v.pop_back();
    if (k == "content-length") {
      char* end = nullptr;
      const unsigned long long n = strtoull(v.c_str(), &end, 10);
      if (end == v.c_str() || *end != 0) { *err = 400; return false; }
      if (n > kBodyLimit) { *err = 413; return false; }
      *content_len = (size_t)n;
    } else if (k == "x-user-id") {
      rq->user = v;
      rq->has_user = true;
    } else if (k == "connection") {
      for (auto& ch : v) ch = (char)tolower((unsigned char)ch);
      if (v == "close") rq->keep_alive = false;
      if (v == "keep-alive") rq->keep_alive = true;
    } else if (k == "transfer-encoding") {
      *err = 400;  // chunked request bodies are not accepted
      return false;
    }
  }
  return true;
}

// tiny JSON field reader for the admin bodies: {"user": "...", "ip": "...", "mode": "..."}
bool json_field(const std::string& body, const char* key, std::string* out) {
  const std::string pat = std::string("\"") + key + "\"";
  size_t k = body.find(pat);
  if (k == std::string::npos) return false;
  k = body.find(':', k + pat.size());
  if (k == std::string::npos) return false;
  ++k;
  while (k < body.size() && isspace((unsigned char)body[k])) ++k;
  if (k >= body.size() || body[k] != '"') return false;
  out->clear();
  for (++k; k < body.size() && body[k] != '"'; ++k) {
    if (body[k] == '\\' && k + 1 < body.size()) ++k;
    out->push_back(body[k]);
  }
  return k < body.size();
}

// POST /admin/{vip,boost,block,unblock} and GET /admin/state: the dashboard's control keys for a headless box
// (tui.rs:126-237).  Loopback only.
std::string admin_response(mq_http_server* s, Conn* c, const Request& rq, const std::string& body, bool keep) {
  if (c->ip != "127.0.0.1") return simple_response(403, "text/plain; charset=utf-8", "admin routes are loopback-only", keep);
  if (rq.path == "/admin/state") {
    const long long need = mq_dispatcher_snapshot_json(s->d, nullptr, 0);
    std::string js((size_t)(need > 0 ? need : 1), '\0');
    mq_dispatcher_snapshot_json(s->d, &js[0], js.size());
    js.resize(strlen(js.c_str()));
    return simple_response(200, "application/json", js, keep);
  }
  if (rq.method != "POST") return simple_response(405, "text/plain; charset=utf-8", "POST only", keep);
  std::string user, ip, mode;
  const bool has_user = json_field(body, "user", &user), has_ip = json_field(body, "ip", &ip);
  json_field(body, "mode", &mode);
  const char* action = nullptr;
  if (rq.path == "/admin/vip") action = mode == "add" ? "vip_add" : mode == "clear" ? "vip_clear" : "vip";
  else if (rq.path == "/admin/boost") action = mode == "add" ? "boost_add" : mode == "clear" ? "boost_clear" : "boost";
  else if (rq.path == "/admin/block") action = has_user && !has_ip ? (mode == "ip" ? "block_ip" : "block_user") : "block_ip";
  else if (rq.path == "/admin/unblock") action = has_user && !has_ip ? (mode == "user" ? "unblock_user" : "unblock") : "unblock_ip";
  if (!action) return simple_response(404, "text/plain", "", keep);
  const int rc = mq_dispatcher_control(s->d, action, has_user ? user.c_str() : nullptr, has_ip ? ip.c_str() : nullptr);
  if (rc != MQ_OK) return simple_response(400, "text/plain; charset=utf-8", mq_last_error(), keep);
  return simple_response(200, "application/json", "{\"ok\":true}", keep);
}

// one complete request sits at the front of c->in; returns false when the connection was closed
bool handle_request(mq_http_server* s, Conn* c, const Request& rq, size_t total_len) {
  const bool keep = rq.keep_alive;
  auto reply = [&](const std::string& r) {
    c->in.erase(0, total_len);
    c->out += r;
    if (!keep) c->close_after_flush = true;
    return true;
  };
  if (rq.path == "/health" && rq.method == "GET")  // main.rs:90 — not queued
    return reply(simple_response(200, "text/plain; charset=utf-8", "OK", keep));
  if (rq.path.compare(0, 7, "/admin/") == 0)
    return reply(admin_response(s, c, rq, c->in.substr(rq.body_off, rq.body_len), keep));
  if (!route_known(rq.path) && !s->allow_all) return reply(simple_response(404, "text/plain", "", keep));

  mq_request q;
  memset(&q, 0, sizeof(q));
  q.endpoint = endpoint_of(rq.path);
  q.body = rq.body_len ? (const uint8_t*)c->in.data() + rq.body_off : nullptr;
  q.body_len = rq.body_len;
  q.path = rq.path.c_str();
  // "stream" defaults: Ollama endpoints stream unless told otherwise, OpenAI endpoints do not
  const bool mentions_stream = rq.body_len && c->in.find("\"stream\"", rq.body_off) != std::string::npos;
  q.stream = mentions_stream ? -1 : ((q.endpoint == MQ_EP_V1_CHAT || q.endpoint == MQ_EP_V1_COMPLETIONS) ? 0 : 1);
  q.ignore_eos = 0;  // generation ends at the model's EOS when it has one (cfg.eos_token_id); random-init models have none
  Pending* p = new Pending();
  p->s = s;
  p->conn_id = c->id;
  p->keep = keep;
  mq_callbacks cb{cb_status, cb_chunk, cb_done};
  uint64_t task = 0;
  // (the body is parsed inside this call, on this thread - never on the scheduler's)
  const int rc = mq_dispatcher_submit(s->d, rq.has_user ? rq.user.c_str() : nullptr, c->ip.c_str(), &q, &cb, p, &task);
  if (rc != MQ_OK) {
    delete p;  // never handed to the dispatcher
    if (rc == MQ_ERR_BLOCKED) return reply(simple_response(403, "text/plain; charset=utf-8", mq_last_error(), keep));  // :370-378
    c->in.erase(0, total_len);
    c->out += simple_response(500, "text/plain; charset=utf-8", mq_last_error(), false);
    c->close_after_flush = true;
    return true;
  }
  c->in.erase(0, total_len);
  c->cur = p;
  c->task = task;
  c->state = Conn::INFLIGHT;
  if (!keep) c->close_after_flush = true;
  epoll_mod(s, c);
  return true;
}

// parse as many complete requests as `in` holds (one at a time: a request in flight parks the rest)
bool pump_requests(mq_http_server* s, Conn* c) {
  while (c->state == Conn::READING && !c->close_after_flush) {
    const size_t hdr_end = c->in.find("\r\n\r\n");
    if (hdr_end == std::string::npos) {
      if (c->in.size() > kHeadLimit) { close_conn(s, c); return false; }
      break;  // head still arriving
    }
    Request rq;
    size_t content_len = 0;
    int err = 0;
    if (!parse_head(c->in, hdr_end, &rq, &content_len, &err)) {
      c->in.clear();
      c->out += simple_response(err, "text/plain", err == 413 ? "body too large" : "bad request", false);
      c->close_after_flush = true;
      break;
    }
    if (c->in.size() < hdr_end + 4 + content_len) break;  // body still arriving
    rq.body_off = hdr_end + 4;
    rq.body_len = content_len;
    if (!handle_request(s, c, rq, hdr_end + 4 + content_len)) return false;
  }
  return flush_out(s, c);
}

void on_readable(mq_http_server* s, Conn* c) {
  char tmp[65536];
  for (;;) {
    const ssize_t r = ::recv(c->fd, tmp, sizeof(tmp), 0);
    if (r > 0) {
      c->in.append(tmp, (size_t)r);
      if (c->in.size() > kBodyLimit + kHeadLimit) { close_conn(s, c); return; }
      if ((size_t)r < sizeof(tmp)) break;
      continue;
    }
    if (r == 0) { close_conn(s, c); return; }
    if (errno == EINTR) continue;
    if (errno == EAGAIN || errno == EWOULDBLOCK) break;
    close_conn(s, c);
    return;
  }
  pump_requests(s, c);
}

void on_ready(mq_http_server* s, Conn* c) {  // the in-flight request has new bytes and / or finished
  Pending* p = c->cur;
  if (!p) return;
  bool done, overflow;
  {
    std::lock_guard<std::mutex> g(p->mu);
    p->signaled = false;
    c->out += p->out;
    p->out.clear();
    done = p->done;
    overflow = p->overflow;
  }
  if (overflow && !done) { close_conn(s, c); return; }  // the client stopped reading: gone (:305-308)
  if (done) {
    pending_unref(p);
    c->cur = nullptr;
    c->state = Conn::READING;
    epoll_mod(s, c);
    pump_requests(s, c);  // flushes, then serves whatever the client pipelined behind this request
    return;
  }
  flush_out(s, c);
}

void accept_all(mq_http_server* s) {
  for (;;) {
    sockaddr_in peer;
    socklen_t pl = sizeof(peer);
    const int fd = ::accept4(s->listen_fd, (sockaddr*)&peer, &pl, SOCK_NONBLOCK | SOCK_CLOEXEC);
    if (fd < 0) {
      if (errno == EINTR) continue;
      return;  // EAGAIN, or out of descriptors: try again on the next wake-up
    }
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    char ip[64] = "0.0.0.0";
    inet_ntop(AF_INET, &peer.sin_addr, ip, sizeof(ip));
    std::unique_ptr<Conn> c(new Conn());
    c->fd = fd;
    c->id = s->next_id++;
    c->ip = ip;
    epoll_event ev;
    memset(&ev, 0, sizeof(ev));
    ev.events = EPOLLIN | EPOLLRDHUP;
    ev.data.u64 = c->id;
    if (epoll_ctl(s->ep, EPOLL_CTL_ADD, fd, &ev) != 0) { ::close(fd); continue; }
    s->conns[c->id] = std::move(c);
  }
}

void event_loop(mq_http_server* s) {
  std::vector<epoll_event> evs(256);
  std::vector<uint64_t> ready;
  while (!s->stop.load()) {
    const int n = epoll_wait(s->ep, evs.data(), (int)evs.size(), 200);
    for (int i = 0; i < n; ++i) {
      const uint64_t id = evs[i].data.u64;
      if (id == 1) { accept_all(s); continue; }
      if (id == 2) {
        uint64_t cnt;
        if (::read(s->evfd, &cnt, sizeof(cnt)) < 0) { /* spurious */ }
        continue;
      }
      auto it = s->conns.find(id);
      if (it == s->conns.end()) continue;  // closed earlier in this batch
      Conn* c = it->second.get();
      const uint32_t e = evs[i].events;
      if (e & (EPOLLERR | EPOLLHUP)) { close_conn(s, c); continue; }
      if ((e & EPOLLRDHUP) && c->state == Conn::INFLIGHT) { close_conn(s, c); continue; }  // client went away mid-request
      if (e & EPOLLOUT) {
        if (!flush_out(s, c)) continue;
      }
      if (e & (EPOLLIN | EPOLLRDHUP)) on_readable(s, c);
    }
    {
      std::lock_guard<std::mutex> g(s->mu);
      ready.swap(s->ready);
    }
    for (uint64_t id : ready) {
      auto it = s->conns.find(id);
      if (it != s->conns.end()) on_ready(s, it->second.get());
    }
    ready.clear();
  }
  // shutdown: every connection goes, every request still in flight is told its client is gone
  while (!s->conns.empty()) close_conn(s, s->conns.begin()->second.get());
}

}  // namespace

extern "C" {

int mq_http_server_start(mq_dispatcher* d, const char* bind_addr, int32_t port, int32_t allow_all_routes,
                         mq_http_server** out) {
  if (!d || !out) return MQ_ERR_INVAL;
  int fd = ::socket(AF_INET, SOCK_STREAM | SOCK_NONBLOCK | SOCK_CLOEXEC, 0);
  if (fd < 0) { mq::set_last_error("socket: %s", strerror(errno)); return MQ_ERR_INVAL; }
  int one = 1;
  setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in a;
  memset(&a, 0, sizeof(a));
  a.sin_family = AF_INET;
  a.sin_port = htons((uint16_t)port);
  if (inet_pton(AF_INET, bind_addr ? bind_addr : "0.0.0.0", &a.sin_addr) != 1) {
    ::close(fd);
    mq::set_last_error("bad bind address");
    return MQ_ERR_INVAL;
  }
  if (::bind(fd, (sockaddr*)&a, sizeof(a)) != 0 || ::listen(fd, 4096) != 0) {
    mq::set_last_error("bind/listen: %s", strerror(errno));
    ::close(fd);
    return MQ_ERR_BUSY;
  }
  socklen_t al = sizeof(a);
  getsockname(fd, (sockaddr*)&a, &al);
  auto* s = new (std::nothrow) mq_http_server();
  if (!s) { ::close(fd); return MQ_ERR_NOMEM; }
  s->d = d;
  s->listen_fd = fd;
  s->port = ntohs(a.sin_port);
  s->allow_all = allow_all_routes != 0;
  s->ep = epoll_create1(EPOLL_CLOEXEC);
  s->evfd = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
  if (s->ep < 0 || s->evfd < 0) {
    mq::set_last_error("epoll/eventfd: %s", strerror(errno));
    if (s->ep >= 0) ::close(s->ep);
    if (s->evfd >= 0) ::close(s->evfd);
    ::close(fd);
    delete s;
    return MQ_ERR_INVAL;
  }
  epoll_event ev;
  memset(&ev, 0, sizeof(ev));
  ev.events = EPOLLIN;
  ev.data.u64 = 1;
  epoll_ctl(s->ep, EPOLL_CTL_ADD, s->listen_fd, &ev);
  ev.data.u64 = 2;
  epoll_ctl(s->ep, EPOLL_CTL_ADD, s->evfd, &ev);
  s->loop = std::thread(event_loop, s);
  *out = s;
  return MQ_OK;
}

int mq_http_server_port(mq_http_server* s) { return s ? s->port : 0; }

void mq_http_server_stop(mq_http_server* s) {
  if (!s) return;
  s->stop.store(true);
  const uint64_t one = 1;
  if (::write(s->evfd, &one, sizeof(one)) < 0) { /* the loop wakes on its timeout */ }
  if (s->loop.joinable()) s->loop.join();  // the loop orphans every in-flight request before it returns
  ::close(s->listen_fd);
  ::close(s->ep);
  ::close(s->evfd);
  delete s;
}

}  // extern "C"
