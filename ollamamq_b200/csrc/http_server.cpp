// HTTP/1.1 ingress in front of the dispatcher — SURVEY.md 8(f) rank 1 ("next" row).
//
// Restates, over plain sockets, what the reference does with axum:
//   /root/reference/src/main.rs:89-121       route table: GET /health -> "OK" un-queued; 20 explicit routes, any
//                                            method, all queued through proxy_handler; optional fallback
//                                            (--allow-all-routes); 1 GiB body limit
//   /root/reference/src/dispatcher.rs:354-428  proxy_handler: X-User-ID (default "anonymous"), 403 "IP blocked" /
//                                            "User blocked", enqueue, first ResponsePart decides the status line,
//                                            streamed body, 500 "Backend error: ..." / "Worker failed to respond"
// One thread per connection (keep-alive).  The worker's callbacks never touch the socket: they append to a bounded
// per-request queue that the connection thread drains, which plays the role of the reference's mpsc::channel(32)
// (:385) — a client that stops reading is treated as gone instead of stalling the GPU worker.
#include "../../include/ollamamq_b200.h"
#include <arpa/inet.h>
#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <string>
#include <sys/socket.h>
#include <thread>
#include <unistd.h>
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

struct Pending {  // one in-flight request of a connection
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::string> out;
  size_t queued = 0;
  int status = 0;
  std::string ctype;
  bool done = false;
  int rc = 0;
  std::string err;
  bool overflow = false;
};

void cb_status(void* u, int32_t status, const char* ctype) {
  auto* p = (Pending*)u;
  std::lock_guard<std::mutex> g(p->mu);
  p->status = status;
  p->ctype = ctype ? ctype : "application/octet-stream";
  p->cv.notify_all();
}
int32_t cb_chunk(void* u, const uint8_t* data, size_t len) {
  auto* p = (Pending*)u;
  std::lock_guard<std::mutex> g(p->mu);
  if (p->overflow) return 1;
  if (p->queued + len > kQueueLimit) {  // client is not reading: same outcome as a failed send (:305-308)
    p->overflow = true;
    return 1;
  }
  p->out.emplace_back((const char*)data, len);
  p->queued += len;
  p->cv.notify_all();
  return 0;
}
void cb_done(void* u, int32_t rc, const char* msg) {
  auto* p = (Pending*)u;
  std::lock_guard<std::mutex> g(p->mu);
  p->done = true;
  p->rc = rc;
  p->err = msg ? msg : "";
  p->cv.notify_all();
}

bool send_all(int fd, const char* b, size_t n) {
  while (n) {
    ssize_t w = ::send(fd, b, n, MSG_NOSIGNAL);
    if (w < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    b += w;
    n -= (size_t)w;
  }
  return true;
}
bool send_simple(int fd, int status, const char* reason, const char* ctype, const std::string& body, bool keep) {
  char h[256];
  int n = snprintf(h, sizeof(h), "HTTP/1.1 %d %s\r\nContent-Type: %s\r\nContent-Length: %zu\r\nConnection: %s\r\n\r\n",
                   status, reason, ctype, body.size(), keep ? "keep-alive" : "close");
  return send_all(fd, h, (size_t)n) && send_all(fd, body.data(), body.size());
}
const char* reason_of(int s) {
  switch (s) {
    case 200: return "OK";
    case 403: return "Forbidden";
    case 404: return "Not Found";
    case 413: return "Payload Too Large";
    case 500: return "Internal Server Error";
    case 501: return "Not Implemented";
    default: return "Status";
  }
}

struct Request {
  std::string method, path, user;
  std::vector<uint8_t> body;
  bool keep_alive = true;
  bool has_user = false;
};

// returns 1 ok, 0 clean EOF, -1 malformed / error, -2 body too large
int read_request(int fd, std::string& buf, Request* rq) {
  size_t hdr_end;
  while ((hdr_end = buf.find("\r\n\r\n")) == std::string::npos) {
    if (buf.size() > 64 * 1024) return -1;
    char tmp[8192];
    ssize_t r = ::recv(fd, tmp, sizeof(tmp), 0);
    if (r == 0) return buf.empty() ? 0 : -1;
    if (r < 0) {
      if (errno == EINTR) continue;
      return -1;
    }
    buf.append(tmp, (size_t)r);
  }
  const std::string head = buf.substr(0, hdr_end);
  size_t line_end = head.find("\r\n");
  const std::string rl = head.substr(0, line_end);
  const size_t s1 = rl.find(' '), s2 = rl.rfind(' ');
  if (s1 == std::string::npos || s2 <= s1) return -1;
  rq->method = rl.substr(0, s1);
  std::string target = rl.substr(s1 + 1, s2 - s1 - 1);
  const size_t qm = target.find('?');  // only uri.path() is used; the query string is dropped (:362)
  rq->path = qm == std::string::npos ? target : target.substr(0, qm);
  rq->keep_alive = rl.substr(s2 + 1) != "HTTP/1.0";
  size_t content_len = 0;
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
    if (k == "content-length") content_len = (size_t)strtoull(v.c_str(), nullptr, 10);
    else if (k == "x-user-id") { rq->user = v; rq->has_user = true; }
    else if (k == "connection") {
      for (auto& ch : v) ch = (char)tolower((unsigned char)ch);
      if (v == "close") rq->keep_alive = false;
      if (v == "keep-alive") rq->keep_alive = true;
    } else if (k == "transfer-encoding") return -1;  // chunked request bodies are not accepted
  }
  if (content_len > kBodyLimit) return -2;
  buf.erase(0, hdr_end + 4);
  while (buf.size() < content_len) {
    char tmp[65536];
    ssize_t r = ::recv(fd, tmp, sizeof(tmp), 0);
    if (r <= 0) {
      if (r < 0 && errno == EINTR) continue;
      return -1;
    }
    buf.append(tmp, (size_t)r);
  }
  rq->body.assign(buf.begin(), buf.begin() + (long)content_len);
  buf.erase(0, content_len);
  return 1;
}

bool peer_gone(int fd) {
  struct pollfd p = {fd, POLLRDHUP, 0};
  if (::poll(&p, 1, 0) > 0 && (p.revents & (POLLRDHUP | POLLHUP | POLLERR))) return true;
  return false;
}

}  // namespace

struct mq_http_server {
  mq_dispatcher* d = nullptr;
  int listen_fd = -1;
  int port = 0;
  bool allow_all = false;
  std::atomic<bool> stop{false};
  std::thread acceptor;
  std::mutex mu;
  std::vector<int> conns;
  std::atomic<int> live{0};
};

namespace {

void serve_connection(mq_http_server* s, int fd, std::string ip) {
  std::string buf;
  for (;;) {
    Request rq;
    const int r = read_request(fd, buf, &rq);
    if (r == -2) { send_simple(fd, 413, reason_of(413), "text/plain", "body too large", false); break; }
    if (r <= 0) break;
    bool keep = rq.keep_alive;
    if (rq.path == "/health" && rq.method == "GET") {  // main.rs:90 — not queued
      if (!send_simple(fd, 200, "OK", "text/plain; charset=utf-8", "OK", keep) || !keep) break;
      continue;
    }
    if (!route_known(rq.path) && !s->allow_all) {
      if (!send_simple(fd, 404, reason_of(404), "text/plain", "", keep) || !keep) break;
      continue;
    }
    Pending pend;
    mq_request q;
    memset(&q, 0, sizeof(q));
    q.endpoint = endpoint_of(rq.path);
    q.body = rq.body.empty() ? nullptr : rq.body.data();
    q.body_len = rq.body.size();
    q.path = rq.path.c_str();
    // "stream" defaults: Ollama endpoints stream unless told otherwise, OpenAI endpoints do not
    const bool mentions_stream =
        !rq.body.empty() && std::string((const char*)rq.body.data(), rq.body.size()).find("\"stream\"") != std::string::npos;
    q.stream = mentions_stream ? -1 : ((q.endpoint == MQ_EP_V1_CHAT || q.endpoint == MQ_EP_V1_COMPLETIONS) ? 0 : 1);
    q.ignore_eos = 0;  // generation ends at the model's EOS when it has one (cfg.eos_token_id); random-init models have none
    mq_callbacks cb{cb_status, cb_chunk, cb_done};
    uint64_t task = 0;
    const int rc = mq_dispatcher_submit(s->d, rq.has_user ? rq.user.c_str() : nullptr, ip.c_str(), &q, &cb, &pend, &task);
    if (rc == MQ_ERR_BLOCKED) {  // :370-378
      if (!send_simple(fd, 403, reason_of(403), "text/plain; charset=utf-8", mq_last_error(), keep) || !keep) break;
      continue;
    }
    if (rc != MQ_OK) {
      send_simple(fd, 500, reason_of(500), "text/plain; charset=utf-8", mq_last_error(), false);
      break;
    }
    // ---- relay: first part decides the status line (:408-427), then chunked body
    bool head_sent = false, sock_ok = true, told_gone = false;
    for (;;) {
      std::deque<std::string> batch;
      bool done;
      int status;
      std::string ctype;
      {
        std::unique_lock<std::mutex> lk(pend.mu);
        pend.cv.wait_for(lk, std::chrono::milliseconds(50),
                         [&] { return !pend.out.empty() || pend.done || (pend.status && !head_sent); });
        batch.swap(pend.out);
        pend.queued = 0;
        done = pend.done;
        status = pend.status;
        ctype = pend.ctype;
      }
      if (sock_ok && status && !head_sent) {
        char h[256];
        int n = snprintf(h, sizeof(h), "HTTP/1.1 %d %s\r\nContent-Type: %s\r\nTransfer-Encoding: chunked\r\nConnection: %s\r\n\r\n",
                         status, reason_of(status), ctype.c_str(), keep ? "keep-alive" : "close");
        sock_ok = send_all(fd, h, (size_t)n);
        head_sent = true;
      }
      for (auto& c : batch) {
        if (!sock_ok || c.empty()) continue;
        char h[32];
        int n = snprintf(h, sizeof(h), "%zx\r\n", c.size());
        sock_ok = send_all(fd, h, (size_t)n) && send_all(fd, c.data(), c.size()) && send_all(fd, "\r\n", 2);
      }
      if ((!sock_ok || peer_gone(fd)) && !told_gone) {  // client went away: tell the dispatcher (:278, :305-308)
        sock_ok = false;
        told_gone = true;
        mq_dispatcher_client_gone(s->d, task);
      }
      if (done) {
        std::lock_guard<std::mutex> g(pend.mu);
        if (pend.out.empty()) break;
      }
    }
    if (!sock_ok) break;
    if (!head_sent) {
      // no Status part ever arrived: Error -> "Backend error: ..." (:423-425), otherwise "Worker failed to respond" (:427)
      const std::string msg = pend.rc == MQ_ERR_BLOCKED || pend.err.empty() ? "Worker failed to respond"
                              : (pend.err.rfind("Backend error:", 0) == 0 ? pend.err : "Backend error: " + pend.err);
      if (!send_simple(fd, 500, reason_of(500), "text/plain; charset=utf-8", msg, keep) || !keep) break;
      continue;
    }
    if (!send_all(fd, "0\r\n\r\n", 5) || !keep) break;
  }
  ::shutdown(fd, SHUT_RDWR);
  ::close(fd);
  {
    std::lock_guard<std::mutex> g(s->mu);
    for (auto it = s->conns.begin(); it != s->conns.end(); ++it)
      if (*it == fd) { s->conns.erase(it); break; }
  }
  s->live.fetch_sub(1);
}

void accept_loop(mq_http_server* s) {
  while (!s->stop.load()) {
    struct pollfd p = {s->listen_fd, POLLIN, 0};
    if (::poll(&p, 1, 100) <= 0) continue;
    sockaddr_in peer;
    socklen_t pl = sizeof(peer);
    int fd = ::accept(s->listen_fd, (sockaddr*)&peer, &pl);
    if (fd < 0) continue;
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    char ip[64] = "0.0.0.0";
    inet_ntop(AF_INET, &peer.sin_addr, ip, sizeof(ip));
    {
      std::lock_guard<std::mutex> g(s->mu);
      s->conns.push_back(fd);
    }
    s->live.fetch_add(1);
    std::thread(serve_connection, s, fd, std::string(ip)).detach();
  }
}

}  // namespace

extern "C" {

int mq_http_server_start(mq_dispatcher* d, const char* bind_addr, int32_t port, int32_t allow_all_routes,
                         mq_http_server** out) {
  if (!d || !out) return MQ_ERR_INVAL;
  int fd = ::socket(AF_INET, SOCK_STREAM, 0);
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
  if (::bind(fd, (sockaddr*)&a, sizeof(a)) != 0 || ::listen(fd, 512) != 0) {
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
  s->acceptor = std::thread(accept_loop, s);
  *out = s;
  return MQ_OK;
}

int mq_http_server_port(mq_http_server* s) { return s ? s->port : 0; }

void mq_http_server_stop(mq_http_server* s) {
  if (!s) return;
  s->stop.store(true);
  if (s->acceptor.joinable()) s->acceptor.join();
  ::close(s->listen_fd);
  {
    std::lock_guard<std::mutex> g(s->mu);
    for (int fd : s->conns) ::shutdown(fd, SHUT_RDWR);  // wakes the connection threads
  }
  for (int i = 0; i < 500 && s->live.load() > 0; ++i) std::this_thread::sleep_for(std::chrono::milliseconds(10));
  delete s;
}

}  // extern "C"
