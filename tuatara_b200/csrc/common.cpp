#include "common.h"

namespace tt {

static thread_local std::string t_last_error;
std::atomic<unsigned long long> g_launches{0};

void set_error(const std::string& msg) { t_last_error = msg; }
const char* last_error() { return t_last_error.c_str(); }

}  // namespace tt
