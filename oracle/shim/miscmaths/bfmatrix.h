// TEST INFRASTRUCTURE ONLY — stand-in for FSL's miscmaths/bfmatrix.h + SpMat.h (not installed, not in
// /root/reference). Own minimal implementation of the surface the reference's meshreg sources name
// (featurespace.h:48-58, similarities.{h,cpp}, reg_tools.cpp:745-866, rigid_costfunction.cpp:89-158):
// a 1-based "big matrix" interface with a dense and a sparse (column-map) implementation.
#ifndef ORACLE_SHIM_BFMATRIX_H
#define ORACLE_SHIM_BFMATRIX_H
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "miscmaths/miscmaths.h"

namespace MISCMATHS {

// Sparse matrix, column-compressed as one ordered map per column (row -> value), 1-based access.
template <class T> class SpMat {
    unsigned int m_ = 0, n_ = 0;
    std::vector<std::map<unsigned int, T>> cols_;
public:
    SpMat() = default;
    SpMat(unsigned int m, unsigned int n) : m_(m), n_(n), cols_(n) {}
    explicit SpMat(const NEWMAT::Matrix& M) : m_(M.Nrows()), n_(M.Ncols()), cols_(M.Ncols()) {
        for (unsigned int c = 1; c <= n_; ++c)
            for (unsigned int r = 1; r <= m_; ++r) if (M(r, c) != 0) cols_[c - 1][r] = (T)M(r, c);
    }
    // "row col value" triplet text; a final "nrows ncols 0" line fixes the shape (FSL's sparse ASCII convention)
    explicit SpMat(const std::string& fname) {
        NEWMAT::Matrix t = read_ascii_matrix(fname);
        if (t.Ncols() != 3) throw std::runtime_error("shim SpMat: expected 3-column triplet file");
        for (int i = 1; i <= t.Nrows(); ++i) { m_ = std::max(m_, (unsigned)t(i, 1)); n_ = std::max(n_, (unsigned)t(i, 2)); }
        cols_.resize(n_);
        for (int i = 1; i <= t.Nrows(); ++i) if (t(i, 3) != 0) cols_[(unsigned)t(i, 2) - 1][(unsigned)t(i, 1)] = (T)t(i, 3);
    }
    unsigned int Nrows() const { return m_; }
    unsigned int Ncols() const { return n_; }
    T Peek(unsigned int r, unsigned int c) const {
        auto it = cols_[c - 1].find(r);
        return it == cols_[c - 1].end() ? T(0) : it->second;
    }
    void Set(unsigned int r, unsigned int c, const T& v) { cols_[c - 1][r] = v; }
    void AddTo(unsigned int r, unsigned int c, const T& v) { cols_[c - 1][r] += v; }
    T& Here(unsigned int r, unsigned int c) { return cols_[c - 1][r]; }
    const std::map<unsigned int, T>& col(unsigned int c) const { return cols_[c - 1]; }
    NEWMAT::Matrix AsNEWMAT() const {
        NEWMAT::Matrix M(m_, n_);
        for (unsigned int c = 1; c <= n_; ++c) for (auto& kv : cols_[c - 1]) M(kv.first, c) = (double)kv.second;
        return M;
    }
    SpMat<T> t() const {
        SpMat<T> r(n_, m_);
        for (unsigned int c = 1; c <= n_; ++c) for (auto& kv : cols_[c - 1]) r.Set(c, kv.first, kv.second);
        return r;
    }
};

// Iterates the stored entries of one column in ascending row order.
class BFMatrixColumnIterator {
    std::vector<std::pair<unsigned int, double>> e_;
    std::size_t pos_ = 0;
public:
    BFMatrixColumnIterator() = default;
    BFMatrixColumnIterator(std::vector<std::pair<unsigned int, double>> e, bool end) : e_(std::move(e)), pos_(end ? e_.size() : 0) {}
    double operator*() const { return e_[pos_].second; }
    unsigned int Row() const { return e_[pos_].first; }
    BFMatrixColumnIterator& operator++() { ++pos_; return *this; }
    BFMatrixColumnIterator operator++(int) { BFMatrixColumnIterator t = *this; ++pos_; return t; }
    bool operator==(const BFMatrixColumnIterator& o) const { return pos_ == o.pos_; }
    bool operator!=(const BFMatrixColumnIterator& o) const { return pos_ != o.pos_; }
};

class BFMatrix {
public:
    virtual ~BFMatrix() = default;
    virtual unsigned int Nrows() const = 0;
    virtual unsigned int Ncols() const = 0;
    virtual double Peek(unsigned int r, unsigned int c) const = 0;
    virtual void Set(unsigned int r, unsigned int c, double v) = 0;
    virtual void AddTo(unsigned int r, unsigned int c, double v) = 0;
    virtual NEWMAT::ReturnMatrix AsMatrix() const = 0;
    virtual std::shared_ptr<BFMatrix> Transpose() const = 0;
    virtual void Print(const std::string& fname = std::string()) const {
        NEWMAT::Matrix M = AsMatrix();
        if (fname.empty()) { std::cout << M; return; }
        std::ofstream f(fname.c_str());
        f.precision(10);
        f << M;
    }
    virtual std::vector<std::pair<unsigned int, double>> column_entries(unsigned int c) const = 0;
    BFMatrixColumnIterator begin(unsigned int c) const { return BFMatrixColumnIterator(column_entries(c), false); }
    BFMatrixColumnIterator end(unsigned int c) const { return BFMatrixColumnIterator(column_entries(c), true); }
};

class FullBFMatrix : public BFMatrix {
    NEWMAT::Matrix M_;
public:
    FullBFMatrix() = default;
    FullBFMatrix(unsigned int m, unsigned int n) : M_(m, n) {}
    explicit FullBFMatrix(const NEWMAT::Matrix& M) : M_(M) {}
    unsigned int Nrows() const override { return M_.Nrows(); }
    unsigned int Ncols() const override { return M_.Ncols(); }
    double Peek(unsigned int r, unsigned int c) const override { return M_(r, c); }
    void Set(unsigned int r, unsigned int c, double v) override { M_(r, c) = v; }
    void AddTo(unsigned int r, unsigned int c, double v) override { M_(r, c) += v; }
    NEWMAT::ReturnMatrix AsMatrix() const override { return M_; }
    std::shared_ptr<BFMatrix> Transpose() const override { return std::make_shared<FullBFMatrix>(M_.t()); }
    std::vector<std::pair<unsigned int, double>> column_entries(unsigned int c) const override {
        std::vector<std::pair<unsigned int, double>> e;
        for (unsigned int r = 1; r <= Nrows(); ++r) e.emplace_back(r, M_(r, c));
        return e;
    }
};

template <class T> class SparseBFMatrix : public BFMatrix {
    SpMat<T> M_;
public:
    SparseBFMatrix() = default;
    SparseBFMatrix(unsigned int m, unsigned int n) : M_(m, n) {}
    explicit SparseBFMatrix(const SpMat<T>& M) : M_(M) {}
    explicit SparseBFMatrix(const NEWMAT::Matrix& M) : M_(M) {}
    unsigned int Nrows() const override { return M_.Nrows(); }
    unsigned int Ncols() const override { return M_.Ncols(); }
    double Peek(unsigned int r, unsigned int c) const override { return (double)M_.Peek(r, c); }
    void Set(unsigned int r, unsigned int c, double v) override { M_.Set(r, c, (T)v); }
    void AddTo(unsigned int r, unsigned int c, double v) override { M_.AddTo(r, c, (T)v); }
    NEWMAT::ReturnMatrix AsMatrix() const override { return M_.AsNEWMAT(); }
    std::shared_ptr<BFMatrix> Transpose() const override { return std::make_shared<SparseBFMatrix<T>>(M_.t()); }
    std::vector<std::pair<unsigned int, double>> column_entries(unsigned int c) const override {
        std::vector<std::pair<unsigned int, double>> e;
        for (auto& kv : M_.col(c)) e.emplace_back(kv.first, (double)kv.second);
        return e;
    }
};

}  // namespace MISCMATHS
#endif
