// Test infrastructure — NOT product code, NOT the Boost library: a std-only stand-in for the few Boost 1.55 names the
// reference headers use, so that the unmodified headers under /root/reference compile here (see oracle/compat/README.md).
#ifndef ISL_COMPAT_BOOST_ARRAY
#define ISL_COMPAT_BOOST_ARRAY
#include <algorithm>
#include <cstddef>
#include <iterator>
#include <stdexcept>
namespace boost {
template <class T, std::size_t N>
class array {
public:
    T elems[N];
    typedef T value_type;
    typedef T* iterator;
    typedef const T* const_iterator;
    typedef T& reference;
    typedef const T& const_reference;
    typedef std::size_t size_type;
    typedef std::ptrdiff_t difference_type;
    typedef std::reverse_iterator<iterator> reverse_iterator;
    typedef std::reverse_iterator<const_iterator> const_reverse_iterator;
    enum { static_size = N };
    iterator begin() { return elems; }
    const_iterator begin() const { return elems; }
    const_iterator cbegin() const { return elems; }
    iterator end() { return elems + N; }
    const_iterator end() const { return elems + N; }
    const_iterator cend() const { return elems + N; }
    reverse_iterator rbegin() { return reverse_iterator(end()); }
    const_reverse_iterator rbegin() const { return const_reverse_iterator(end()); }
    reverse_iterator rend() { return reverse_iterator(begin()); }
    const_reverse_iterator rend() const { return const_reverse_iterator(begin()); }
    reference operator[](size_type i) { return elems[i]; }
    const_reference operator[](size_type i) const { return elems[i]; }
    reference at(size_type i) { if (i >= N) throw std::out_of_range("array"); return elems[i]; }
    const_reference at(size_type i) const { if (i >= N) throw std::out_of_range("array"); return elems[i]; }
    reference front() { return elems[0]; }
    const_reference front() const { return elems[0]; }
    reference back() { return elems[N - 1]; }
    const_reference back() const { return elems[N - 1]; }
    static size_type size() { return N; }
    static bool empty() { return false; }
    static size_type max_size() { return N; }
    T* data() { return elems; }
    const T* data() const { return elems; }
    T* c_array() { return elems; }
    void swap(array& o) { std::swap_ranges(begin(), end(), o.begin()); }
    void assign(const T& v) { std::fill_n(begin(), N, v); }
    void fill(const T& v) { std::fill_n(begin(), N, v); }
};
template <class T>
class array<T, 0> {
public:
    typedef T value_type;
    typedef T* iterator;
    typedef const T* const_iterator;
    typedef T& reference;
    typedef const T& const_reference;
    typedef std::size_t size_type;
    typedef std::ptrdiff_t difference_type;
    enum { static_size = 0 };
    iterator begin() { return nullptr; }
    const_iterator begin() const { return nullptr; }
    iterator end() { return nullptr; }
    const_iterator end() const { return nullptr; }
    reference operator[](size_type) { throw std::out_of_range("array<0>"); }
    const_reference operator[](size_type) const { throw std::out_of_range("array<0>"); }
    reference at(size_type) { throw std::out_of_range("array<0>"); }
    const_reference at(size_type) const { throw std::out_of_range("array<0>"); }
    static size_type size() { return 0; }
    static bool empty() { return true; }
    T* data() { return nullptr; }
    const T* data() const { return nullptr; }
    void swap(array&) {}
    void assign(const T&) {}
    void fill(const T&) {}
};
template <class T, std::size_t N>
bool operator==(const array<T, N>& a, const array<T, N>& b) { return std::equal(a.begin(), a.end(), b.begin()); }
template <class T, std::size_t N>
bool operator!=(const array<T, N>& a, const array<T, N>& b) { return !(a == b); }
template <class T, std::size_t N>
bool operator<(const array<T, N>& a, const array<T, N>& b) {
    return std::lexicographical_compare(a.begin(), a.end(), b.begin(), b.end());
}
template <class T, std::size_t N>
bool operator>(const array<T, N>& a, const array<T, N>& b) { return b < a; }
template <class T, std::size_t N>
bool operator<=(const array<T, N>& a, const array<T, N>& b) { return !(b < a); }
template <class T, std::size_t N>
bool operator>=(const array<T, N>& a, const array<T, N>& b) { return !(a < b); }
}  // namespace boost
#endif
