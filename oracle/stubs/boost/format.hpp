// Test-infrastructure stub (oracle/): lets the reference's impl/util.hpp and
// impl/parallel_hash_array.hpp compile without Boost. Formatting is discarded.
#pragma once
#include <string>
#include <ostream>
namespace boost {
class format {
public:
    explicit format(const char *) {}
    explicit format(const std::string &) {}
    template <typename T> format &operator%(const T &) { return *this; }
};
inline std::ostream &operator<<(std::ostream &os, const format &) { return os; }
}  // namespace boost
