// Stand-in for Boost.Accumulators (see ../../README.md): accumulator_set<Sample, stats<mean, median, variance[, count]>>.
//
// Restated from the published algorithms [Boost is not in this image]:
//   sum_impl<Sample>            sum += sample, IN THE SAMPLE TYPE (an unsigned short sum wraps mod 65536)
//   mean_impl (lazy)            fdiv(sum, count): double for integral samples, the sample type for floating ones
//   median                      p_square_quantile_impl with p = 0.5: five markers, first five samples sorted when the
//                               fifth arrives, result = heights[2]
//   variance (immediate)        n > 1:  tmp = sample - mean;  var = var*(n-1)/n + tmp*tmp/(n-1)
// The sample passed to operator() keeps its own type (acc(item.protein_length) hands an unsigned int to an
// accumulator_set<unsigned short>): only the running sum is narrowed.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <type_traits>

namespace boost { namespace accumulators {

namespace tag { struct mean {}; struct median {}; struct variance {}; struct count {}; }
template <class... T> struct stats {};

template <class Sample, class Stats> class accumulator_set {
public:
    using float_type = typename std::conditional<std::is_integral<Sample>::value, double, Sample>::type;

    template <class A> void operator()(const A &x) {
        static const float_type inc[5] = {float_type(0), float_type(0.25), float_type(0.5), float_type(0.75), float_type(1)};
        n_ += 1;
        sum_ = static_cast<Sample>(sum_ + x);
        if (n_ <= 5) {
            h_[n_ - 1] = static_cast<float_type>(x);
            if (n_ == 5) std::sort(h_, h_ + 5);
        } else {
            const float_type xf = static_cast<float_type>(x);
            std::size_t k;
            if (xf < h_[0]) { h_[0] = xf; k = 1; }
            else if (h_[4] <= xf) { h_[4] = xf; k = 4; }
            else k = static_cast<std::size_t>(std::upper_bound(h_, h_ + 5, xf) - h_);
            for (std::size_t i = k; i < 5; ++i) pos_[i] += 1;
            for (std::size_t i = 0; i < 5; ++i) des_[i] += inc[i];
            for (std::size_t i = 1; i <= 3; ++i) {
                const float_type d = des_[i] - pos_[i];
                const float_type dp = pos_[i + 1] - pos_[i];
                const float_type dm = pos_[i - 1] - pos_[i];
                const float_type hp = (h_[i + 1] - h_[i]) / dp;
                const float_type hm = (h_[i - 1] - h_[i]) / dm;
                if ((d >= 1 && dp > 1) || (d <= -1 && dm < -1)) {
                    const short sign_d = static_cast<short>(d / std::abs(d));
                    const float_type h = h_[i] + sign_d / (dp - dm) * ((sign_d - dm) * hp + (dp - sign_d) * hm);
                    if (h_[i - 1] < h && h < h_[i + 1]) h_[i] = h;
                    else {
                        if (d > 0) h_[i] += hp;
                        if (d < 0) h_[i] -= hm;
                    }
                    pos_[i] += sign_d;
                }
            }
        }
        if (n_ > 1) {
            const float_type tmp = x - mean_value();
            var_ = (var_ * (n_ - 1)) / n_ + (tmp * tmp) / (n_ - 1);
        }
    }
    float_type mean_value() const { return static_cast<float_type>(sum_) / n_; }
    float_type median_value() const { return h_[2]; }
    float_type variance_value() const { return var_; }
    std::size_t count_value() const { return n_; }

private:
    std::size_t n_ = 0;
    Sample sum_ = Sample();
    float_type var_ = float_type();
    float_type h_[5] = {0, 0, 0, 0, 0};
    float_type pos_[5] = {1, 2, 3, 4, 5};
    float_type des_[5] = {1, 2, 3, 4, 5};
};

template <class S, class T> typename accumulator_set<S, T>::float_type mean(const accumulator_set<S, T> &a) { return a.mean_value(); }
template <class S, class T> typename accumulator_set<S, T>::float_type median(const accumulator_set<S, T> &a) { return a.median_value(); }
template <class S, class T> typename accumulator_set<S, T>::float_type variance(const accumulator_set<S, T> &a) { return a.variance_value(); }
template <class S, class T> std::size_t count(const accumulator_set<S, T> &a) { return a.count_value(); }

}}  // namespace boost::accumulators
