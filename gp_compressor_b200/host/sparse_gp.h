// sparse_gp.h — host shell with the reference's sparse_gp<Kernel, Noise> interface
// (/root/reference/src/sparse_gp.h:36-48) over the C ABI of include/gpc.h.  Only
// sparse_gp<rbf_kernel, gaussian_noise> exists on the device: there is no CPU fallback, other policy
// combinations fail to compile.
#pragma once
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/gpc.h"
#include "dense_shim.h"

// policy tags with the reference's constructor defaults (rbf_kernel.h:24, gaussian_noise.h:10)
struct rbf_kernel {
    double sigmaf_sq, l_sq;
    rbf_kernel(double sigmaf_sq = 100e-0f, double l_sq = 1 * 1) : sigmaf_sq(sigmaf_sq), l_sq(l_sq) {}
};
struct gaussian_noise {
    double s20;
    explicit gaussian_noise(double s20 = 1e-1f) : s20(s20) {}
};

template <class Kernel, class Noise>
class sparse_gp {
    static_assert(std::is_same<Kernel, rbf_kernel>::value && std::is_same<Noise, gaussian_noise>::value,
                  "only sparse_gp<rbf_kernel, gaussian_noise> is implemented on the B200 path");
    gpc_handle* h_ = nullptr;
    gpc_config cfg_;
    int size_ = 0;
    bool fitted_ = false;
    void check(int rc) const { if (rc != GPC_OK) throw std::runtime_error(std::string("gpc: ") + gpc_last_error(h_)); }
    void open() {
        if (!h_ && gpc_create(&cfg_, &h_) != GPC_OK) throw std::runtime_error("gpc_create failed: no usable sm_100 CUDA device");
    }
public:
    typedef Kernel kernel_type;
    typedef Noise noise_type;
    // sparse_gp(int capacity = 100, double s0 = 1e-1f), sparse_gp.h:48
    sparse_gp(int capacity = 100, double s0 = 1e-1f, const Kernel& kernel = Kernel()) {
        gpc_config_default(&cfg_);
        cfg_.capacity = capacity;
        cfg_.s0 = s0;
        cfg_.sigmaf_sq = kernel.sigmaf_sq;
        cfg_.l_sq = kernel.l_sq;
        cfg_.rgb_rand = 0;  // a stand-alone process consumes only its own shuffle draws
        cfg_.keep_state = 1;  // C is kept: predict_measurements returns sigma, likelihoods and derivatives need it
    }
    ~sparse_gp() { if (h_) gpc_destroy(h_); }
    sparse_gp(const sparse_gp&) = delete;
    sparse_gp& operator=(const sparse_gp&) = delete;
    gpc_config& config() { return cfg_; }

    // void add_measurements(const MatrixXd& X /* n x 2 */, const VectorXd& y), sparse_gp.hpp:59-86.  Like the reference,
    // a second call adds to the fitted process (gpc_add_measurements continues the kept state).
    void add_measurements(const Eigen::MatrixXd& X, const Eigen::VectorXd& y) {
        if (X.cols() != 2 || X.rows() != y.rows()) throw std::invalid_argument("sparse_gp::add_measurements: X must be n x 2, y n");
        open();
        const int64_t off[2] = {0, (int64_t)X.rows()};
        if (fitted_) check(gpc_add_measurements(h_, 1, off, X.data(), X.data() + X.rows(), y.data()));
        else check(gpc_fit_patches(h_, 1, off, X.data(), X.data() + X.rows(), y.data()));  // column-major: col 0 then col 1
        int32_t n = 0;
        check(gpc_get_params(h_, &n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
        size_ = n;
        fitted_ = true;
    }
    // void predict_measurements(VectorXd& f_star, const MatrixXd& X_star, VectorXd& sigconf, bool conf = false), :299-308
    void predict_measurements(Eigen::VectorXd& f_star, const Eigen::MatrixXd& X_star, Eigen::VectorXd& sigconf, bool conf = false) {
        if (!fitted_) throw std::runtime_error("sparse_gp::predict_measurements before add_measurements");
        const long m = X_star.rows();
        f_star.resize(m);
        sigconf.resize(m);
        const int64_t off[2] = {0, (int64_t)m};
        check(gpc_evaluate_patches(h_, 1, off, X_star.data(), X_star.data() + m, nullptr, conf ? 1 : 0, f_star.data(), sigconf.data(),
                                   nullptr, nullptr));
    }
    // void compute_likelihoods(VectorXd& l, const MatrixXd& X, const VectorXd& y), sparse_gp.hpp:414-420
    void compute_likelihoods(Eigen::VectorXd& l, const Eigen::MatrixXd& X, const Eigen::VectorXd& y) {
        if (!fitted_) throw std::runtime_error("sparse_gp::compute_likelihoods before add_measurements");
        if (X.cols() != 2 || X.rows() != y.rows()) throw std::invalid_argument("sparse_gp::compute_likelihoods: X must be n x 2, y n");
        const long m = X.rows();
        l.resize(m);
        const int64_t off[2] = {0, (int64_t)m};
        check(gpc_evaluate_patches(h_, 1, off, X.data(), X.data() + m, y.data(), 0, nullptr, nullptr, l.data(), nullptr));
    }
    // void compute_derivatives(MatrixXd& dX /* n x 3 */, const MatrixXd& X, const VectorXd& y), sparse_gp.hpp:459-468
    void compute_derivatives(Eigen::MatrixXd& dX, const Eigen::MatrixXd& X, const Eigen::VectorXd& y) {
        if (!fitted_) throw std::runtime_error("sparse_gp::compute_derivatives before add_measurements");
        if (X.cols() != 2 || X.rows() != y.rows()) throw std::invalid_argument("sparse_gp::compute_derivatives: X must be n x 2, y n");
        const long m = X.rows();
        std::vector<double> rows((size_t)3 * m);
        const int64_t off[2] = {0, (int64_t)m};
        check(gpc_evaluate_patches(h_, 1, off, X.data(), X.data() + m, y.data(), 0, nullptr, nullptr, nullptr, rows.data()));
        dX.resize(m, 3);
        for (long i = 0; i < m; i++)
            for (int c = 0; c < 3; c++) dX(i, c) = rows[3 * i + c];
    }
    int size() { return size_; }                       // sparse_gp.hpp:36-39
    void reset() { if (h_) { gpc_destroy(h_); h_ = nullptr; } size_ = 0; fitted_ = false; }
};
