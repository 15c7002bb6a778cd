// The reference's only driver for this path (src/test_gp_compress.cpp:11-25) against the host shells:
// build a cloud, gp_compressor comp(cloud, 0.15f, 20), save_compressed, load_compressed; then a
// stand-alone sparse_gp<rbf_kernel, gaussian_noise>.  Prints checksums that the Python test compares
// with the ctypes path.  The cloud is read from a raw file of 32-byte PointXYZRGB records (argv[1]).
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "gp_compressor.h"
#include "pointcloud_compressor.h"

int main(int argc, char** argv) {
    if (argc < 2) { std::cerr << "usage: test_host_api cloud.bin" << std::endl; return 2; }
    std::ifstream f(argv[1], std::ios::binary | std::ios::ate);
    const size_t bytes = (size_t)f.tellg();
    f.seekg(0);
    pcl::PointCloud<pcl::PointXYZRGB>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZRGB>());
    cloud->resize(bytes / 32);
    f.read(reinterpret_cast<char*>(cloud->points.data()), (std::streamsize)(bytes / 32 * 32));
    try {
        gp_compressor comp(cloud, 0.15f, 20);
        comp.save_compressed("test");
        pcl::PointCloud<pcl::PointXYZRGB>::Ptr out = comp.load_compressed();
        unsigned long long sum = 1469598103934665603ULL;  // FNV-1a over the decoded cloud
        const unsigned char* p = reinterpret_cast<const unsigned char*>(out->points.data());
        for (size_t i = 0; i < out->points.size() * 32; i++) { sum ^= p[i]; sum *= 1099511628211ULL; }
        std::printf("CLOUD %zu %016llx\n", out->points.size(), sum);
        if (argc > 2) { std::ofstream o(argv[2], std::ios::binary); o.write(reinterpret_cast<const char*>(out->points.data()), (std::streamsize)(out->points.size() * 32)); }

        sparse_gp<rbf_kernel, gaussian_noise> gp(2, 1e-1f);
        gp.config().shuffle = 0;
        Eigen::MatrixXd X(4, 2);
        Eigen::VectorXd y(4);
        const double xs[4][3] = {{0, 0, .01}, {.05, 0, .02}, {0, .05, -.01}, {.05, .05, 0}};
        for (int i = 0; i < 4; i++) { X(i, 0) = xs[i][0]; X(i, 1) = xs[i][1]; y(i) = xs[i][2]; }
        gp.add_measurements(X, y);
        Eigen::MatrixXd Xs(1, 2);
        Xs(0, 0) = 0.025; Xs(0, 1) = 0.025;
        Eigen::VectorXd fs, sg;
        gp.predict_measurements(fs, Xs, sg);
        std::printf("SOGP %d %.17g\n", gp.size(), fs(0));
        // rows N2 / N4 through the same shell: sigma, confidence, likelihood, likelihood gradient
        Eigen::VectorXd cf, lk, ys(1);
        Eigen::MatrixXd dX;
        ys(0) = 0.004;
        gp.predict_measurements(fs, Xs, cf, true);
        gp.compute_likelihoods(lk, Xs, ys);
        gp.compute_derivatives(dX, Xs, ys);
        std::printf("EVAL %.17g %.17g %.17g %.17g %.17g %.17g\n", sg(0), cf(0), lk(0), dX(0, 0), dX(0, 1), dX(0, 2));
        // a second add_measurements accumulates, as in the reference (sparse_gp.hpp:59-86)
        Eigen::MatrixXd X2(3, 2);
        Eigen::VectorXd y2(3);
        const double more[3][3] = {{.02, .03, .015}, {.04, .01, .005}, {.01, .04, -.002}};
        for (int i = 0; i < 3; i++) { X2(i, 0) = more[i][0]; X2(i, 1) = more[i][1]; y2(i) = more[i][2]; }
        gp.add_measurements(X2, y2);
        gp.predict_measurements(fs, Xs, sg);
        std::printf("SOGP2 %d %.17g %.17g\n", gp.size(), fs(0), sg(0));
        bool threw = false;
        try { pointcloud_decompressor d; d.load_compressed("x"); } catch (const std::exception&) { threw = true; }
        std::printf("KSVD_SHELL %d\n", (int)threw);
    } catch (const std::exception& e) {
        std::printf("ERROR %s\n", e.what());
        return 1;
    }
    return 0;
}
