// dense_shim.h — the sliver of Eigen the sparse_gp public API needs at its edge (MatrixXd / VectorXd
// as argument and result containers; column-major like Eigen).  No arithmetic lives here.
#pragma once
#include <cstddef>
#include <vector>

namespace Eigen {

class MatrixXd {
    std::vector<double> d_;
    long r_ = 0, c_ = 0;
public:
    MatrixXd() {}
    MatrixXd(long r, long c) : d_((size_t)r * c, 0.0), r_(r), c_(c) {}
    void resize(long r, long c) { d_.assign((size_t)r * c, 0.0); r_ = r; c_ = c; }
    long rows() const { return r_; }
    long cols() const { return c_; }
    double& operator()(long i, long j) { return d_[(size_t)j * r_ + i]; }
    double operator()(long i, long j) const { return d_[(size_t)j * r_ + i]; }
    double* data() { return d_.data(); }
    const double* data() const { return d_.data(); }
};

class VectorXd {
    std::vector<double> d_;
public:
    VectorXd() {}
    explicit VectorXd(long n) : d_((size_t)n, 0.0) {}
    void resize(long n) { d_.assign((size_t)n, 0.0); }
    long rows() const { return (long)d_.size(); }
    long size() const { return (long)d_.size(); }
    double& operator()(long i) { return d_[(size_t)i]; }
    double operator()(long i) const { return d_[(size_t)i]; }
    double* data() { return d_.data(); }
    const double* data() const { return d_.data(); }
};

}  // namespace Eigen
