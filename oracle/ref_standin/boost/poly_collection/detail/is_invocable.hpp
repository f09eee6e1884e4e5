// boost/poly_collection/detail/is_invocable.hpp STAND-IN (test infrastructure): C++17 has the trait in <type_traits>
#pragma once
#include <type_traits>
namespace boost { namespace poly_collection { namespace detail {
template <class F, class... Args>
using is_invocable = std::is_invocable<F, Args...>;
}}}
