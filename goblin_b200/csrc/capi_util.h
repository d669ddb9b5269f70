// Shared helpers for the extern "C" layer.
#pragma once
#include <string>

namespace gb {
void setLastError(const std::string& e);
int failWith(int code, const std::string& e);
} // namespace gb
