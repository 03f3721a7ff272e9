// oracle/shim/boost/variant.hpp -- TEST INFRASTRUCTURE ONLY.  Boost is not installed; the reference's Metrics.h:185 only
// needs a sum type to hold its metric values.
#pragma once
#include <variant>
namespace boost { template <class... T> using variant = std::variant<T...>; }
