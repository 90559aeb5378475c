// boost::bind / _1 / _2 — stand-in (TEST INFRASTRUCTURE): grid_map.cpp binds its synchronised callbacks with them; the
// result is handed to an inert registerCallback, so it only has to compile.
#pragma once
namespace boost {
struct bound_nothing {};
template <class... A> inline bound_nothing bind(A&&...) { return bound_nothing(); }
}  // namespace boost
struct pop_placeholder {};
static const pop_placeholder _1{}, _2{};
