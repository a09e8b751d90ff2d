// TEST INFRASTRUCTURE ONLY -- stand-in for <boost/container/flat_set.hpp> (Boost is not installed in the build
// container) so that the REFERENCE's own translation unit graphembed/pyx/impl/precision.cpp compiles unmodified from
// where it lies under /root/reference (recipe: oracle/Makefile, output: oracle/_ref/).  Only what that file uses:
// flat_multiset(comp), reserve, insert (after the existing equivalent elements, as Boost's insert_equal does: it
// inserts at upper_bound), lower_bound, cbegin -- on a sorted std::vector, random-access iterators.
#pragma once
#include <algorithm>
#include <vector>

namespace boost {
namespace container {

template <typename Key, typename Compare>
class flat_multiset {
 public:
  using container_type = std::vector<Key>;
  using iterator = typename container_type::const_iterator;
  using const_iterator = typename container_type::const_iterator;
  explicit flat_multiset(const Compare& comp) : comp_(comp) {}
  void reserve(std::size_t n) { data_.reserve(n); }
  const_iterator cbegin() const { return data_.cbegin(); }
  const_iterator cend() const { return data_.cend(); }
  const_iterator lower_bound(const Key& k) const { return std::lower_bound(data_.cbegin(), data_.cend(), k, comp_); }
  iterator insert(const Key& k) {
    auto pos = std::upper_bound(data_.cbegin(), data_.cend(), k, comp_);
    return data_.insert(pos, k);
  }

 private:
  container_type data_;
  Compare comp_;
};

}  // namespace container
}  // namespace boost
