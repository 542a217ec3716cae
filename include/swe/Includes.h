// Includes.h — basic types of the C++17 host API (mirror of upstream include/Includes.h, without
// Eigen: the hot path runs on the GPU behind swe_b200.h, the host only needs small PODs).
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "../swe_b200.h"
#include "../swe_constants.h"

using Idx = int64_t;  // upstream: Eigen::Index (include/Includes.h:18)

// upstream include/Includes.h:21
enum class Boundaries : Idx { SOLID_WALL = -1, FREE_FLOW = -2, PERIODIC = -3, CUSTOM = -4 };

template <size_t k>
struct Array : std::array<double, k> {  // upstream: Eigen::Array<double, k, 1>
    Array() { this->fill(0.); }
    Array(std::initializer_list<double> l) { size_t i = 0; for (double x : l) if (i < k) (*this)[i++] = x; }
    Array &operator+=(const Array &o) { for (size_t i = 0; i < k; ++i) (*this)[i] += o[i]; return *this; }
    Array &operator-=(const Array &o) { for (size_t i = 0; i < k; ++i) (*this)[i] -= o[i]; return *this; }
    Array &operator*=(double s) { for (size_t i = 0; i < k; ++i) (*this)[i] *= s; return *this; }
    friend Array operator+(Array a, const Array &b) { return a += b; }
    friend Array operator-(Array a, const Array &b) { return a -= b; }
    friend Array operator*(double s, Array a) { return a *= s; }
    friend Array operator*(Array a, double s) { return a *= s; }
    friend Array operator/(Array a, double s) { for (size_t i = 0; i < k; ++i) a[i] /= s; return a; }
};

// upstream: Eigen::Array<double, k, Dynamic>, column-major => column i is contiguous
template <size_t k>
struct Storage {
    std::vector<double> data;
    Storage() = default;
    explicit Storage(size_t cols) : data(k * cols, 0.) {}
    size_t cols() const { return data.size() / k; }
    void resize(size_t cols) { data.assign(k * cols, 0.); }
    double &operator()(size_t r, size_t c) { return data[k * c + r]; }
    double operator()(size_t r, size_t c) const { return data[k * c + r]; }
    Array<k> col(size_t c) const { Array<k> a; for (size_t r = 0; r < k; ++r) a[r] = data[k * c + r]; return a; }
    void set_col(size_t c, const Array<k> &a) { for (size_t r = 0; r < k; ++r) data[k * c + r] = a[r]; }
};

constexpr inline double tol = SWE_TOL;  // upstream include/Includes.h:30
