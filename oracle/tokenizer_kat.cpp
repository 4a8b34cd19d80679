// TEST INFRASTRUCTURE ONLY.  Known-answer generator for the reference tokenizer table:
// restates the constructor logic of /root/reference/tuatara.cpp:25-48 (same string-literal
// escapes, same "later duplicate wins" stoi loop) and prints the table, so the escapes are
// evaluated by a real C++ compiler rather than by eye.  Build: g++ -o _ref/tokenizer_kat tokenizer_kat.cpp
#include <cstdio>
#include <map>
#include <string>

int main() {
  const char BOS = '[', EOS = ']', PAD = 'P';
  const std::string charset =
      "0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ!\"#$%&"
      "\\'()*+,-./:;<=>?@[\\]^_`{|}~";
  std::string itos = charset;
  itos.insert(itos.begin(), EOS);
  itos.push_back(BOS);
  itos.push_back(PAD);
  std::map<char, size_t> stoi;
  for (size_t i = 0; i < itos.size(); ++i) stoi[itos[i]] = i;
  std::printf("%zu %zu %zu %zu\n", itos.size(), stoi[EOS], stoi[BOS], stoi[PAD]);
  for (unsigned char c : itos) std::printf("%02x", c);
  std::printf("\n");
  return 0;
}
