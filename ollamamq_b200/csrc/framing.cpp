#include "framing.hpp"
#include "../../include/ollamamq_b200.h"
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

namespace mq {

// ------------------------------------------------------------------ minimal JSON walker
namespace {
struct J {
  const char* p;
  const char* e;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool lit(const char* s) {
    size_t n = strlen(s);
    if ((size_t)(e - p) >= n && memcmp(p, s, n) == 0) { p += n; return true; }
    return false;
  }
  bool str(std::string* out) {
    ws();
    if (p >= e || *p != '"') return false;
    ++p;
    while (p < e && *p != '"') {
      if (*p == '\\' && p + 1 < e) {
        ++p;
        char c = *p++;
        switch (c) {
          case 'n': if (out) out->push_back('\n'); break;
          case 't': if (out) out->push_back('\t'); break;
          case 'r': if (out) out->push_back('\r'); break;
          case 'b': if (out) out->push_back('\b'); break;
          case 'f': if (out) out->push_back('\f'); break;
          case 'u': {
            unsigned cp = 0;
            for (int i = 0; i < 4 && p < e; ++i, ++p) cp = cp * 16 + (unsigned)(isdigit((unsigned char)*p) ? *p - '0' : (tolower(*p) - 'a' + 10));
            if (out) {
              if (cp < 0x80) out->push_back((char)cp);
              else if (cp < 0x800) { out->push_back((char)(0xC0 | (cp >> 6))); out->push_back((char)(0x80 | (cp & 0x3F))); }
              else { out->push_back((char)(0xE0 | (cp >> 12))); out->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out->push_back((char)(0x80 | (cp & 0x3F))); }
            }
            break;
          }
          default: if (out) out->push_back(c);
        }
      } else {
        if (out) out->push_back(*p);
        ++p;
      }
    }
    if (p >= e) return false;
    ++p;
    return true;
  }
  bool skip() {  // skip any value
    ws();
    if (p >= e) return false;
    if (*p == '"') return str(nullptr);
    if (*p == '{' || *p == '[') {
      const char open = *p, close = open == '{' ? '}' : ']';
      ++p;
      ws();
      if (p < e && *p == close) { ++p; return true; }
      for (;;) {
        if (open == '{') { if (!str(nullptr)) return false; ws(); if (p >= e || *p != ':') return false; ++p; }
        if (!skip()) return false;
        ws();
        if (p < e && *p == ',') { ++p; continue; }
        if (p < e && *p == close) { ++p; return true; }
        return false;
      }
    }
    while (p < e && *p != ',' && *p != '}' && *p != ']' && !isspace((unsigned char)*p)) ++p;
    return true;
  }
  bool number(double* v) {
    ws();
    char* end = nullptr;
    *v = strtod(p, &end);
    if (end == p) return false;
    p = end;
    return true;
  }
  bool int_array(std::vector<int32_t>* out) {
    ws();
    if (p >= e || *p != '[') return false;
    ++p;
    ws();
    if (p < e && *p == ']') { ++p; return true; }
    for (;;) {
      double v;
      if (!number(&v)) return false;
      out->push_back((int32_t)v);
      ws();
      if (p < e && *p == ',') { ++p; continue; }
      if (p < e && *p == ']') { ++p; return true; }
      return false;
    }
  }
};

void parse_content(J& j, std::string* text) {  // string or [{type:text,text:..},..]
  j.ws();
  if (j.p < j.e && *j.p == '"') { std::string s; if (j.str(&s)) *text += s; return; }
  if (j.p < j.e && *j.p == '[') {
    ++j.p;
    for (;;) {
      j.ws();
      if (j.p < j.e && *j.p == ']') { ++j.p; return; }
      if (j.p < j.e && *j.p == '{') {
        ++j.p;
        for (;;) {
          std::string k;
          j.ws();
          if (j.p < j.e && *j.p == '}') { ++j.p; break; }
          if (!j.str(&k)) return;
          j.ws(); if (j.p < j.e && *j.p == ':') ++j.p;
          if (k == "text") { std::string s; if (j.str(&s)) *text += s; } else if (!j.skip()) return;
          j.ws(); if (j.p < j.e && *j.p == ',') ++j.p;
        }
      } else if (!j.skip()) return;
      j.ws(); if (j.p < j.e && *j.p == ',') ++j.p;
    }
  }
  j.skip();
}

void parse_messages(J& j, std::string* text) {
  j.ws();
  if (j.p >= j.e || *j.p != '[') { j.skip(); return; }
  ++j.p;
  for (;;) {
    j.ws();
    if (j.p < j.e && *j.p == ']') { ++j.p; return; }
    if (j.p >= j.e || *j.p != '{') return;
    ++j.p;
    for (;;) {
      j.ws();
      if (j.p < j.e && *j.p == '}') { ++j.p; break; }
      std::string k;
      if (!j.str(&k)) return;
      j.ws(); if (j.p < j.e && *j.p == ':') ++j.p;
      if (k == "content") { parse_content(j, text); text->push_back('\n'); }
      else if (!j.skip()) return;
      j.ws(); if (j.p < j.e && *j.p == ',') ++j.p;
    }
    j.ws(); if (j.p < j.e && *j.p == ',') ++j.p;
  }
}
}  // namespace

bool parse_body(const std::string& body, int endpoint, ParsedBody* out) {
  (void)endpoint;
  J j{body.data(), body.data() + body.size()};
  j.ws();
  if (j.p >= j.e || *j.p != '{') return false;
  ++j.p;
  for (;;) {
    j.ws();
    if (j.p < j.e && *j.p == '}') return true;
    std::string k;
    if (!j.str(&k)) return false;
    j.ws();
    if (j.p >= j.e || *j.p != ':') return false;
    ++j.p;
    j.ws();
    if (k == "model") { if (!j.str(&out->model)) return false; }
    else if (k == "prompt") {
      if (j.p < j.e && *j.p == '[') { if (!j.int_array(&out->tokens)) return false; }
      else if (!j.str(&out->text)) { if (!j.skip()) return false; }
    }
    else if (k == "context") { if (!j.int_array(&out->tokens)) return false; }
    else if (k == "messages") parse_messages(j, &out->text);
    else if (k == "stream") {
      out->has_stream = true;
      if (j.lit("true")) out->stream = true; else if (j.lit("false")) out->stream = false; else if (!j.skip()) return false;
    }
    else if (k == "max_tokens" || k == "max_completion_tokens" || k == "num_predict") {
      double v; if (!j.number(&v)) { if (!j.skip()) return false; } else out->num_predict = (int)v;
    }
    else if (k == "temperature") { double v; if (j.number(&v)) { out->temperature = v; out->has_temperature = true; } else if (!j.skip()) return false; }
    else if (k == "top_p") { double v; if (j.number(&v)) { out->top_p = v; out->has_top_p = true; } else if (!j.skip()) return false; }
    else if (k == "seed") { double v; if (j.number(&v)) { out->seed = (unsigned long long)v; out->has_seed = true; } else if (!j.skip()) return false; }
    else if (k == "options") {
      j.ws();
      if (j.p < j.e && *j.p == '{') {
        ++j.p;
        for (;;) {
          j.ws();
          if (j.p < j.e && *j.p == '}') { ++j.p; break; }
          std::string ok;
          if (!j.str(&ok)) return false;
          j.ws(); if (j.p < j.e && *j.p == ':') ++j.p;
          if (ok == "num_predict") { double v; if (j.number(&v)) out->num_predict = (int)v; else if (!j.skip()) return false; }
          else if (ok == "temperature") { double v; if (j.number(&v)) { out->temperature = v; out->has_temperature = true; } else if (!j.skip()) return false; }
          else if (ok == "top_k") { double v; if (j.number(&v)) { out->top_k = (long long)v; out->has_top_k = true; } else if (!j.skip()) return false; }
          else if (ok == "top_p") { double v; if (j.number(&v)) { out->top_p = v; out->has_top_p = true; } else if (!j.skip()) return false; }
          else if (ok == "seed") { double v; if (j.number(&v)) { out->seed = (unsigned long long)v; out->has_seed = true; } else if (!j.skip()) return false; }
          else if (!j.skip()) return false;
          j.ws(); if (j.p < j.e && *j.p == ',') ++j.p;
        }
      } else if (!j.skip()) return false;
    }
    else if (!j.skip()) return false;
    j.ws();
    if (j.p < j.e && *j.p == ',') ++j.p;
  }
}

// deterministic byte-level tokenizer: random-init weights have no vocabulary
std::vector<int32_t> byte_tokenize(const std::string& text, int vocab) {
  std::vector<int32_t> t;
  t.reserve(text.size());
  for (unsigned char c : text) t.push_back((int32_t)(c % (unsigned)vocab));
  return t;
}

std::string token_text(int tok) {
  char b[24];
  snprintf(b, sizeof(b), " t%d", tok);
  return b;
}

const char* content_type_for(int endpoint, int stream) {
  if (endpoint == MQ_EP_RAW_TOKENS) return "application/octet-stream";
  if (!stream) return "application/json";
  return (endpoint == MQ_EP_V1_CHAT || endpoint == MQ_EP_V1_COMPLETIONS) ? "text/event-stream" : "application/x-ndjson";
}

static std::string json_escape(const std::string& s) {
  std::string o;
  for (unsigned char c : s) {
    if (c == '"' || c == '\\') { o.push_back('\\'); o.push_back((char)c); }
    else if (c < 0x20) { char b[8]; snprintf(b, sizeof(b), "\\u%04x", c); o += b; }
    else o.push_back((char)c);
  }
  return o;
}

static std::string now_iso() {
  char b[40];
  time_t t = time(nullptr);
  struct tm tmv;
  gmtime_r(&t, &tmv);
  strftime(b, sizeof(b), "%Y-%m-%dT%H:%M:%SZ", &tmv);
  return b;
}

std::string frame_token(int endpoint, const char* model, int tok) {
  const std::string m = json_escape(model), txt = json_escape(token_text(tok));
  char buf[512];
  switch (endpoint) {
    case MQ_EP_API_GENERATE:
      snprintf(buf, sizeof(buf), "{\"model\":\"%s\",\"created_at\":\"%s\",\"response\":\"%s\",\"done\":false}\n",
               m.c_str(), now_iso().c_str(), txt.c_str());
      break;
    case MQ_EP_API_CHAT:
      snprintf(buf, sizeof(buf),
               "{\"model\":\"%s\",\"created_at\":\"%s\",\"message\":{\"role\":\"assistant\",\"content\":\"%s\"},\"done\":false}\n",
               m.c_str(), now_iso().c_str(), txt.c_str());
      break;
    case MQ_EP_V1_CHAT:
      snprintf(buf, sizeof(buf),
               "data: {\"id\":\"chatcmpl-mq\",\"object\":\"chat.completion.chunk\",\"created\":%ld,\"model\":\"%s\","
               "\"choices\":[{\"index\":0,\"delta\":{\"content\":\"%s\"},\"finish_reason\":null}]}\n\n",
               (long)time(nullptr), m.c_str(), txt.c_str());
      break;
    default:
      snprintf(buf, sizeof(buf),
               "data: {\"id\":\"cmpl-mq\",\"object\":\"text_completion\",\"created\":%ld,\"model\":\"%s\","
               "\"choices\":[{\"text\":\"%s\",\"index\":0,\"finish_reason\":null}]}\n\n",
               (long)time(nullptr), m.c_str(), txt.c_str());
  }
  return buf;
}

void other_route_response(const std::string& path, const char* model, int* status, std::string* ctype, std::string* body) {
  const std::string m = json_escape(model);
  *status = 200;
  *ctype = "application/json";
  if (path == "/") {
    *ctype = "text/plain; charset=utf-8";
    *body = "Ollama is running";
  } else if (path == "/api/version") {
    *body = "{\"version\":\"0.0.0-ollamamq-b200\"}";
  } else if (path == "/api/tags" || path == "/api/ps") {
    *body = "{\"models\":[{\"name\":\"" + m + "\",\"model\":\"" + m + "\",\"details\":{\"format\":\"bf16\",\"family\":\"llama\"}}]}";
  } else if (path == "/api/show") {
    *body = "{\"details\":{\"format\":\"bf16\",\"family\":\"llama\"},\"model_info\":{\"general.name\":\"" + m + "\"}}";
  } else if (path == "/v1/models") {
    *body = "{\"object\":\"list\",\"data\":[{\"id\":\"" + m + "\",\"object\":\"model\",\"owned_by\":\"ollamamq-b200\"}]}";
  } else if (path.rfind("/v1/models/", 0) == 0) {
    *body = "{\"id\":\"" + json_escape(path.substr(11)) + "\",\"object\":\"model\",\"owned_by\":\"ollamamq-b200\"}";
  } else {
    *status = 501;  // embeddings and model management are not served by this worker
    *body = "{\"error\":\"" + json_escape(path) + " is not implemented by the B200 worker\"}";
  }
}

std::string frame_final(int endpoint, int stream, const char* model, const std::string& agg, int n_prompt, int n_gen,
                        bool stopped) {
  const char* why = stopped ? "stop" : "length";  // EOS reached vs generation budget spent
  if (endpoint == MQ_EP_RAW_TOKENS) return stream ? std::string() : agg;
  const std::string m = json_escape(model), txt = json_escape(agg);
  std::string o;
  char buf[512];
  if (endpoint == MQ_EP_API_GENERATE || endpoint == MQ_EP_API_CHAT) {
    o = "{\"model\":\"" + m + "\",\"created_at\":\"" + now_iso() + "\",";
    if (endpoint == MQ_EP_API_GENERATE) o += "\"response\":\"" + (stream ? std::string() : txt) + "\",";
    else o += "\"message\":{\"role\":\"assistant\",\"content\":\"" + (stream ? std::string() : txt) + "\"},";
    snprintf(buf, sizeof(buf), "\"done\":true,\"done_reason\":\"%s\",\"prompt_eval_count\":%d,\"eval_count\":%d}\n",
             why, n_prompt, n_gen);
    return o + buf;
  }
  const bool chat = endpoint == MQ_EP_V1_CHAT;
  if (stream) {
    snprintf(buf, sizeof(buf),
             chat ? "data: {\"id\":\"chatcmpl-mq\",\"object\":\"chat.completion.chunk\",\"created\":%ld,\"model\":\"%s\","
                    "\"choices\":[{\"index\":0,\"delta\":{},\"finish_reason\":\"%s\"}]}\n\ndata: [DONE]\n\n"
                  : "data: {\"id\":\"cmpl-mq\",\"object\":\"text_completion\",\"created\":%ld,\"model\":\"%s\","
                    "\"choices\":[{\"text\":\"\",\"index\":0,\"finish_reason\":\"%s\"}]}\n\ndata: [DONE]\n\n",
             (long)time(nullptr), m.c_str(), why);
    return buf;
  }
  snprintf(buf, sizeof(buf), "\"usage\":{\"prompt_tokens\":%d,\"completion_tokens\":%d,\"total_tokens\":%d}}", n_prompt,
           n_gen, n_prompt + n_gen);
  if (chat)
    o = "{\"id\":\"chatcmpl-mq\",\"object\":\"chat.completion\",\"model\":\"" + m +
        "\",\"choices\":[{\"index\":0,\"message\":{\"role\":\"assistant\",\"content\":\"" + txt +
        "\"},\"finish_reason\":\"" + std::string(why) + "\"}],";
  else
    o = "{\"id\":\"cmpl-mq\",\"object\":\"text_completion\",\"model\":\"" + m + "\",\"choices\":[{\"text\":\"" + txt +
        "\",\"index\":0,\"finish_reason\":\"" + std::string(why) + "\"}],";
  return o + buf;
}


// ------------------------------------------------------------------ embeddings
bool parse_embed_body(const std::string& body, ParsedEmbed* out) {
  J j{body.data(), body.data() + body.size()};
  j.ws();
  if (j.p >= j.e || *j.p != '{') return false;
  ++j.p;
  for (;;) {
    j.ws();
    if (j.p < j.e && *j.p == '}') return true;
    std::string k;
    if (!j.str(&k)) return false;
    j.ws();
    if (j.p >= j.e || *j.p != ':') return false;
    ++j.p;
    j.ws();
    if (k == "model") { if (!j.str(&out->model)) return false; }
    else if (k == "input" || k == "prompt") {
      if (j.p < j.e && *j.p == '"') { std::string s; if (!j.str(&s)) return false; out->texts.push_back(s); }
      else if (j.p < j.e && *j.p == '[') {
        const char* save = j.p;
        ++j.p;
        j.ws();
        if (j.p < j.e && *j.p == ']') { ++j.p; }
        else if (*j.p == '"') {                       // ["s", ...]
          for (;;) {
            std::string s;
            if (!j.str(&s)) return false;
            out->texts.push_back(s);
            j.ws();
            if (j.p < j.e && *j.p == ',') { ++j.p; continue; }
            if (j.p < j.e && *j.p == ']') { ++j.p; break; }
            return false;
          }
        } else if (*j.p == '[') {                     // [[ids], ...]
          for (;;) {
            std::vector<int32_t> t;
            if (!j.int_array(&t)) return false;
            out->token_seqs.push_back(t);
            j.ws();
            if (j.p < j.e && *j.p == ',') { ++j.p; continue; }
            if (j.p < j.e && *j.p == ']') { ++j.p; break; }
            return false;
          }
        } else {                                      // [ids]
          j.p = save;
          std::vector<int32_t> t;
          if (!j.int_array(&t)) return false;
          out->token_seqs.push_back(t);
        }
      } else if (!j.skip()) return false;
    }
    else if (!j.skip()) return false;
    j.ws();
    if (j.p < j.e && *j.p == ',') ++j.p;
  }
}

std::vector<int32_t> embed_tokenize(const std::string& text, int vocab, int max_len) {
  // BERT conventions where the vocabulary has room for them: [CLS] = 101, [SEP] = 102, bytes from 1000
  const int cls = vocab > 1300 ? 101 : 1, sep = vocab > 1300 ? 102 : 2, base = vocab > 1300 ? 1000 : 3;
  std::vector<int32_t> t;
  t.push_back(cls);
  for (unsigned char c : text) {
    if ((int)t.size() >= max_len - 1) break;
    t.push_back(base + (int)(c % (unsigned)(vocab - base)));
  }
  if ((int)t.size() < max_len) t.push_back(sep);
  return t;
}

static void append_vec(std::string& o, const float* v, int dim) {
  char b[32];
  o += '[';
  for (int i = 0; i < dim; ++i) {
    snprintf(b, sizeof b, i ? ",%.8g" : "%.8g", (double)v[i]);
    o += b;
  }
  o += ']';
}
std::string frame_embeddings(const std::string& path, const char* model, const float* emb, int n, int dim, int n_tokens) {
  const std::string m = json_escape(model);
  std::string o;
  o.reserve((size_t)n * dim * 14 + 256);
  char b[160];
  if (path == "/api/embeddings") {  // legacy Ollama route: one prompt, one vector
    o = "{\"embedding\":";
    append_vec(o, emb, n > 0 ? dim : 0);
    return o + "}";
  }
  if (path == "/v1/embeddings") {
    o = "{\"object\":\"list\",\"data\":[";
    for (int i = 0; i < n; ++i) {
      if (i) o += ',';
      o += "{\"object\":\"embedding\",\"embedding\":";
      append_vec(o, emb + (size_t)i * dim, dim);
      snprintf(b, sizeof b, ",\"index\":%d}", i);
      o += b;
    }
    snprintf(b, sizeof b, "],\"model\":\"%s\",\"usage\":{\"prompt_tokens\":%d,\"total_tokens\":%d}}", m.c_str(), n_tokens, n_tokens);
    return o + b;
  }
  o = "{\"model\":\"" + m + "\",\"embeddings\":[";
  for (int i = 0; i < n; ++i) {
    if (i) o += ',';
    append_vec(o, emb + (size_t)i * dim, dim);
  }
  snprintf(b, sizeof b, "],\"prompt_eval_count\":%d}", n_tokens);
  return o + b;
}

}  // namespace mq

// ------------------------------------------------------------------ host-only test ABI for the parsers / framers
extern "C" {
static long long emit_json(const std::string& o, char* out, size_t cap) {
  if (out && cap > o.size()) memcpy(out, o.c_str(), o.size() + 1);
  return (long long)o.size() + 1;
}
long long mq_debug_parse_body(int32_t endpoint, const uint8_t* body, size_t len, int32_t vocab, char* out, size_t cap) {
  using namespace mq;
  ParsedBody pb;
  const bool ok = parse_body(std::string((const char*)body, body ? len : 0), endpoint, &pb);
  std::string toks = "[";
  const std::vector<int32_t> t = !pb.tokens.empty() ? pb.tokens : byte_tokenize(pb.text, vocab > 0 ? vocab : 256);
  for (size_t i = 0; i < t.size(); ++i) toks += (i ? "," : "") + std::to_string(t[i]);
  toks += "]";
  char b[512];
  snprintf(b, sizeof b, "{\"ok\":%s,\"has_stream\":%s,\"stream\":%s,\"num_predict\":%d,\"has_temperature\":%s,\"temperature\":%.9g,"
           "\"has_top_k\":%s,\"top_k\":%lld,\"has_top_p\":%s,\"top_p\":%.9g,\"has_seed\":%s,\"seed\":%llu,\"n_text\":%zu,",
           ok ? "true" : "false", pb.has_stream ? "true" : "false", pb.stream ? "true" : "false", pb.num_predict,
           pb.has_temperature ? "true" : "false", pb.temperature, pb.has_top_k ? "true" : "false", pb.top_k,
           pb.has_top_p ? "true" : "false", pb.top_p, pb.has_seed ? "true" : "false", pb.seed, pb.text.size());
  return emit_json(std::string(b) + "\"model\":\"" + json_escape(pb.model) + "\",\"tokens\":" + toks + "}", out, cap);
}
long long mq_debug_parse_embed(const uint8_t* body, size_t len, int32_t vocab, int32_t max_len, char* out, size_t cap) {
  using namespace mq;
  ParsedEmbed pe;
  const bool ok = parse_embed_body(std::string((const char*)body, body ? len : 0), &pe);
  std::string o = std::string("{\"ok\":") + (ok ? "true" : "false") + ",\"model\":\"" + json_escape(pe.model) + "\",\"seqs\":[";
  bool first = true;
  auto put = [&](const std::vector<int32_t>& t) {
    o += first ? "[" : ",[";
    first = false;
    for (size_t i = 0; i < t.size(); ++i) o += (i ? "," : "") + std::to_string(t[i]);
    o += "]";
  };
  for (auto& t : pe.texts) put(embed_tokenize(t, vocab, max_len));
  for (auto& t : pe.token_seqs) put(t);
  return emit_json(o + "]}", out, cap);
}
long long mq_debug_frame_embeddings(const char* path, const char* model, const float* emb, int32_t n, int32_t dim,
                                    int32_t n_tokens, char* out, size_t cap) {
  return emit_json(mq::frame_embeddings(path ? path : "/api/embed", model ? model : "", emb, n, dim, n_tokens), out, cap);
}
long long mq_debug_frame_final(int32_t endpoint, int32_t stream, const char* model, const char* agg, int32_t n_prompt,
                               int32_t n_gen, int32_t stopped, char* out, size_t cap) {
  return emit_json(mq::frame_final(endpoint, stream, model ? model : "", agg ? agg : "", n_prompt, n_gen, stopped != 0), out, cap);
}
}  // extern "C"

namespace mq {
}  // namespace mq
