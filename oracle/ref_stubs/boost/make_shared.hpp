// Stub: boost::shared_ptr / make_shared as the std ones.  Test infrastructure only.
#pragma once
#include <memory>
namespace boost {
template <typename T> using shared_ptr = std::shared_ptr<T>;
template <typename T, typename... Args> std::shared_ptr<T> make_shared(Args&&... a) { return std::make_shared<T>(std::forward<Args>(a)...); }
}  // namespace boost
