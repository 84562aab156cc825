/* Shim for boost/type_index.hpp: type_id<T>().pretty_name() for one error message. */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <map>
#include <numeric>
#include <string>
#include <typeinfo>
#include <vector>
namespace boost { namespace typeindex {
struct gvref_tid { const char* n; std::string pretty_name() const { return n; } };
template <typename T> inline gvref_tid type_id() { return gvref_tid{typeid(T).name()}; }
}}
