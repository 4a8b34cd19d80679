// TEST INFRASTRUCTURE ONLY -- driver around the REFERENCE's own Tokenizer class.
//
// oracle/build_ref.py cuts /root/reference/tuatara.cpp:25-117 (class Tokenizer, verbatim, never committed) into
// oracle/_ref/tokenizer_class.inc and compiles this file against the LibTorch headers / libraries of the torch wheel
// (the only third-party dependency that part of the reference has).  The binary is what pins the tokenizer:
//   tokenizer_ref table                      -> "<itos size> <eos> <bos> <pad>\n<itos as hex>\n"
//   tokenizer_ref decode N L C < floats.bin  -> one line per item: hex of Tokenizer::decode(dists)[i] after the
//                                               caller-side truncation at the first ']' (tuatara.cpp:497-502)
// Reading private members for the `table` command uses the -Dprivate=public trick on the included class only.
#include <torch/torch.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#define private public
#include "_ref/tokenizer_class.inc"
#undef private

static void print_hex(const std::string& s) {
  for (unsigned char c : s) std::printf("%02x", c);
  std::printf("\n");
}

int main(int argc, char** argv) {
  Tokenizer tok;
  if (argc >= 2 && std::strcmp(argv[1], "table") == 0) {
    std::printf("%zu %zu %zu %zu\n", tok.itos.size(), tok.eos_id, tok.bos_id, tok.pad_id);
    print_hex(tok.itos);
    return 0;
  }
  if (argc >= 5 && std::strcmp(argv[1], "decode") == 0) {
    const long N = std::atol(argv[2]), L = std::atol(argv[3]), C = std::atol(argv[4]);
    std::vector<float> buf(static_cast<size_t>(N) * L * C);
    if (std::fread(buf.data(), sizeof(float), buf.size(), stdin) != buf.size()) { std::fprintf(stderr, "short read\n"); return 2; }
    torch::Tensor logits = torch::from_blob(buf.data(), {N, L, C}, torch::kFloat32);
    torch::Tensor probs = logits.softmax(-1);                      // tuatara.cpp:486
    std::vector<std::string> tokens = tok.decode(probs, false);     // tuatara.cpp:492
    for (std::string token_str : tokens) {                          // tuatara.cpp:495-505
      size_t eos_pos = token_str.find(tok.EOS);
      if (eos_pos != std::string::npos) token_str = token_str.substr(0, eos_pos);
      print_hex(token_str);
    }
    return 0;
  }
  std::fprintf(stderr, "usage: tokenizer_ref table | decode N L C < float32.bin\n");
  return 2;
}
