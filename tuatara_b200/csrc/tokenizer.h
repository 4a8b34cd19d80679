// The reference's Tokenizer (tuatara.cpp:25-117) and the truncation of :492-505, quirks included.
#pragma once
#include <stdint.h>

#include <string>

namespace tt {

struct TokenizerTable {
  std::string itos;  // 98 symbols: ']' + 95-char charset + '[' + 'P'
  int eos_id, bos_id, pad_id;
};
const TokenizerTable& tokenizer_table();

// ids: the per-position argmax over the 95 classes (first max wins, like at::max on CPU).
// Drops every id == eos_id (tuatara.cpp:108-116), maps through itos (:93-99), cuts at the first
// ']' character (:497-502).
std::string decode_ids(const int32_t* ids, int len);

}  // namespace tt
