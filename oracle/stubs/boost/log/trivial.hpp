// Test-infrastructure stub (oracle/): null-sink replacement for Boost.Log.
#pragma once
#include <ostream>
namespace oracle_stub {
struct NullLog {
    template <typename T> NullLog &operator<<(const T &) { return *this; }
    NullLog &operator<<(std::ostream &(*)(std::ostream &)) { return *this; }
};
}  // namespace oracle_stub
#define BOOST_LOG_TRIVIAL(lvl) ::oracle_stub::NullLog()
