#include "tokenizer.h"

#include <map>

namespace tt {

const TokenizerTable& tokenizer_table() {
  static const TokenizerTable table = [] {
    TokenizerTable t;
    // tuatara.cpp:32-34.  The second literal starts with an escaped backslash followed by an
    // apostrophe, so the charset holds a backslash before the apostrophe: 95 symbols, not 94.
    const std::string charset =
        "0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ!\"#$%&"
        "\\'()*+,-./:;<=>?@[\\]^_`{|}~";
    t.itos = std::string(1, ']') + charset + "[" + "P";  // :36-39
    std::map<char, int> stoi;
    for (size_t i = 0; i < t.itos.size(); ++i) stoi[t.itos[i]] = static_cast<int>(i);  // :41-43, later wins
    t.eos_id = stoi[']'];  // 88: the ']' inside the charset, not slot 0
    t.bos_id = stoi['['];  // 96
    t.pad_id = stoi['P'];  // 97
    return t;
  }();
  return table;
}

std::string decode_ids(const int32_t* ids, int len) {
  const TokenizerTable& t = tokenizer_table();
  std::string s;
  for (int i = 0; i < len; ++i) {
    const int id = ids[i];
    if (id == t.eos_id) continue;
    if (id < 0 || id >= static_cast<int>(t.itos.size())) continue;
    const char ch = t.itos[id];
    if (ch == ']') break;
    s.push_back(ch);
  }
  return s;
}

}  // namespace tt
