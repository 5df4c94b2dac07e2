// TEST INFRASTRUCTURE ONLY — stand-in for FSL's miscmaths/histogram.h (not installed, not in
// /root/reference). Used by the reference only for the optional intensity normalisation ("--IN",
// reg_tools.cpp:745-802), which is off the hot path and not enabled by any BASELINE config. Own
// implementation of a CDF histogram match with the same public surface; NOT claimed FSL-exact.
#ifndef ORACLE_SHIM_HISTOGRAM_H
#define ORACLE_SHIM_HISTOGRAM_H
#include <algorithm>
#include <vector>
#include "armawrap/newmat.h"

namespace MISCMATHS {

class Histogram {
    NEWMAT::ColumnVector src_, excl_;
    int bins_;
    double lo_ = 0, hi_ = 0;
    std::vector<double> hist_, cdf_;
    bool use(int i) const { return excl_.Nrows() != src_.Nrows() || excl_(i) != 0; }
    int bin_of(double v) const {
        if (hi_ <= lo_) return 1;
        int b = (int)((double)bins_ * (v - lo_) / (hi_ - lo_)) + 1;
        return std::max(1, std::min(b, bins_));
    }
    double value_of(double b) const { return lo_ + b * (hi_ - lo_) / (double)bins_; }
public:
    Histogram(const NEWMAT::ColumnVector& d, int nbins) : src_(d), bins_(nbins) {}
    void setexclusion(const NEWMAT::ColumnVector& e) { excl_ = e; }
    void generate() { NEWMAT::ColumnVector e; generate(e); }
    void generate(const NEWMAT::ColumnVector& e) {
        excl_ = e;
        bool first = true;
        for (int i = 1; i <= src_.Nrows(); ++i) if (use(i)) {
            if (first) { lo_ = hi_ = src_(i); first = false; }
            lo_ = std::min(lo_, src_(i)); hi_ = std::max(hi_, src_(i));
        }
        hist_.assign(bins_ + 1, 0.0);
        for (int i = 1; i <= src_.Nrows(); ++i) if (use(i)) hist_[bin_of(src_(i))] += 1.0;
    }
    void generateCDF() {
        cdf_.assign(bins_ + 1, 0.0);
        double tot = 0; for (int b = 1; b <= bins_; ++b) tot += hist_[b];
        double run = 0;
        for (int b = 1; b <= bins_; ++b) { run += hist_[b]; cdf_[b] = tot > 0 ? run / tot : 0.0; }
    }
    // map every non-excluded source value to the reference value with the same cumulative probability
    void match(Histogram& ref) {
        for (int i = 1; i <= src_.Nrows(); ++i) if (use(i)) {
            double p = cdf_[bin_of(src_(i))];
            int b = 1;
            while (b < ref.bins_ && ref.cdf_[b] < p) ++b;
            src_(i) = ref.value_of((double)b - 0.5);
        }
    }
    NEWMAT::ColumnVector getsourceData() const { return src_; }
};

}  // namespace MISCMATHS
#endif
