/* Shim for boost/algorithm/string.hpp: split + is_any_of only, plus the std
 * headers the reference picks up transitively from Boost/casacore. */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <map>
#include <numeric>
#include <string>
#include <vector>
namespace boost {
struct gvref_any_of { std::string set; };
inline gvref_any_of is_any_of(const std::string& s) { return gvref_any_of{s}; }
template <typename Seq>
inline Seq& split(Seq& out, const std::string& in, const gvref_any_of& pred) {
  out.clear();
  std::string cur;
  for (char c : in) {
    if (pred.set.find(c) != std::string::npos) { out.push_back(cur); cur.clear(); }
    else cur.push_back(c);
  }
  out.push_back(cur);
  return out;
}
}  // namespace boost
