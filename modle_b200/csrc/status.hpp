// Thread-local error message plumbing of the C ABI (include/modle_b200.h).
#pragma once
#include <string>

namespace modle_b200 {
extern thread_local std::string g_last_error;
int fail(int code, const std::string& msg);
}  // namespace modle_b200
