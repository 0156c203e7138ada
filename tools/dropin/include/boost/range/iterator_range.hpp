// boost::make_iterator_range stand-in (int8_calibrator.cpp:34).
#pragma once
namespace boost {
template <class It>
struct iterator_range_stub {
    It b, e;
    It begin() const { return b; }
    It end() const { return e; }
};
template <class It>
iterator_range_stub<It> make_iterator_range(It b, It e) { return iterator_range_stub<It>{b, e}; }
}  // namespace boost
