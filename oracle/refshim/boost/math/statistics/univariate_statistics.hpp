// Stand-in for <boost/math/statistics/univariate_statistics.hpp> (see ../../../README.md): the three functions
// src/call_functions.tcc:51-53 calls on a std::vector<float>, restated from the published algorithms
// [Boost is not in this image]: mean = four interleaved running means combined at the end (Boost.Math >= 1.72,
// random-access real input); median / median_absolute_deviation by std::nth_element (they permute the input).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <limits>
namespace boost { namespace math { namespace statistics {

template <class Container> auto mean(const Container &v) {
    using Real = typename Container::value_type;
    const std::size_t n = v.size();
    Real mu0 = 0, mu1 = 0, mu2 = 0, mu3 = 0;
    Real i = 1;
    const std::size_t body = n - (n % 4);
    std::size_t k = 0;
    for (; k < body; k += 4) {
        const Real inv = Real(1) / i;
        const Real t0 = v[k] - mu0, t1 = v[k + 1] - mu1, t2 = v[k + 2] - mu2, t3 = v[k + 3] - mu3;
        mu0 += t0 * inv; mu1 += t1 * inv; mu2 += t2 * inv; mu3 += t3 * inv;
        i += 1;
    }
    const Real num1 = Real(body) / Real(4);
    const Real num2 = num1 + Real(n % 4);
    for (; k < n; ++k) { mu3 += (v[k] - mu3) / i; i += 1; }
    return (num1 * (mu0 + mu1 + mu2) + num2 * mu3) / Real(n);
}

template <class Container> auto median(Container &v) {
    using Real = typename Container::value_type;
    const std::size_t n = v.size();
    if (n % 2 == 0) {
        auto mid = v.begin() + (n / 2 - 1);
        std::nth_element(v.begin(), mid, v.end());
        auto next = std::min_element(mid + 1, v.end());
        return Real((*mid + *next) / 2);
    }
    auto mid = v.begin() + n / 2;
    std::nth_element(v.begin(), mid, v.end());
    return Real(*mid);
}

template <class Container>
auto median_absolute_deviation(Container &v, typename Container::value_type center = std::numeric_limits<typename Container::value_type>::quiet_NaN()) {
    using Real = typename Container::value_type;
    using std::abs;
    if (std::isnan(center)) center = median(v);
    const std::size_t n = v.size();
    auto cmp = [center](Real a, Real b) { return abs(a - center) < abs(b - center); };
    if (n % 2 == 0) {
        auto mid = v.begin() + (n / 2 - 1);
        std::nth_element(v.begin(), mid, v.end(), cmp);
        auto next = std::min_element(mid + 1, v.end(), cmp);
        return Real((abs(*mid - center) + abs(*next - center)) / abs(Real(2)));
    }
    auto mid = v.begin() + n / 2;
    std::nth_element(v.begin(), mid, v.end(), cmp);
    return Real(abs(*mid - center));
}

}}}  // namespace boost::math::statistics
