// Stand-in for glog: LOG(severity) swallows its operands.
#pragma once
#include <ostream>
namespace lk_shim {
struct NullLog {
    template <class T>
    NullLog& operator<<(const T&) { return *this; }
    NullLog& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
    NullLog& operator<<(std::ios_base& (*)(std::ios_base&)) { return *this; }
};
}  // namespace lk_shim
#define LOG(severity) ::lk_shim::NullLog()
#define LOG_IF(severity, cond) ::lk_shim::NullLog()
