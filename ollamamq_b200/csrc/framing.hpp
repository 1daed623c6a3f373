// Request-body parsing and response framing for the endpoints the worker terminates.
// The reference relays opaque bytes (dispatcher.rs:287-312), so these formats follow the public Ollama /
// OpenAI wire shapes seen in the reference's own client script (test_dispatcher.sh:37-39,69,93) and README.
// Parity here is UNPINNED by the reference (SURVEY.md 7 "hard parts").
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mq {

struct ParsedBody {
  std::string model;
  std::string text;             // prompt / concatenated message contents
  std::vector<int32_t> tokens;  // "context": [ids] (Ollama /api/generate) or "prompt": [ids] (OpenAI completions)
  bool has_stream = false;
  bool stream = true;
  int num_predict = 0;          // options.num_predict | max_tokens | max_completion_tokens
  // sampling: options.{temperature, top_k, top_p, seed} (Ollama) or top-level temperature / top_p / seed (OpenAI)
  bool has_temperature = false, has_top_k = false, has_top_p = false, has_seed = false;
  double temperature = 0, top_p = 0;
  long long top_k = 0;
  unsigned long long seed = 0;
};

bool parse_body(const std::string& body, int endpoint, ParsedBody* out);
std::vector<int32_t> byte_tokenize(const std::string& text, int vocab);
std::string token_text(int tok);
const char* content_type_for(int endpoint, int stream);
std::string frame_token(int endpoint, const char* model, int tok);
// MQ_EP_OTHER: what the backend answers on the non-generation routes (main.rs:92-112) without touching the GPU
void other_route_response(const std::string& path, const char* model, int* status, std::string* ctype, std::string* body);
std::string frame_final(int endpoint, int stream, const char* model, const std::string& agg, int n_prompt, int n_gen,
                        bool stopped = false);  // stopped: done_reason / finish_reason "stop" (EOS) instead of "length"

// ---- embeddings (/api/embed, /api/embeddings, /v1/embeddings)
struct ParsedEmbed {
  std::string model;
  std::vector<std::string> texts;              // "input": "s" | ["s", ...]   or "prompt": "s" (/api/embeddings)
  std::vector<std::vector<int32_t>> token_seqs;  // "input": [ids] | [[ids], ...] (OpenAI token inputs)
};
bool parse_embed_body(const std::string& body, ParsedEmbed* out);
// [CLS] bytes... [SEP] over a byte-level vocabulary (random-init weights have no word pieces), truncated to max_len
std::vector<int32_t> embed_tokenize(const std::string& text, int vocab, int max_len);
// emb: n rows of `dim` floats; the reply shape follows the route (Ollama /api/embed, legacy /api/embeddings, OpenAI)
std::string frame_embeddings(const std::string& path, const char* model, const float* emb, int n, int dim, int n_tokens);

}  // namespace mq
