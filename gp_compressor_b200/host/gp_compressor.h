// gp_compressor.h — host shell with the reference's gp_compressor interface
// (/root/reference/src/gp_compressor.h:65-67) over the C ABI of include/gpc.h.
#pragma once
#include <iostream>
#include <stdexcept>
#include <string>

#include "../../include/gpc.h"
#include "pcl_shim.h"
#include "sparse_gp.h"

class gp_compressor {
public:
    typedef pcl::PointXYZRGB point;
    typedef pcl::PointCloud<point> pointcloud;
protected:
    pointcloud::Ptr cloud;  // the reference copies its input (gp_compressor.cpp:15)
    double res;
    int sz;
    gpc_handle* gpu_ = nullptr;
    gpc_config cfg_;
    void check(int rc) const { if (rc != GPC_OK) throw std::runtime_error(std::string("gpc: ") + gpc_last_error(gpu_)); }
public:
    // gp_compressor(pointcloud::ConstPtr ncloud, double res = 0.1f, int sz = 10), gp_compressor.h:65
    gp_compressor(pointcloud::ConstPtr ncloud, double res = 0.1f, int sz = 10) : cloud(new pointcloud()), res(res), sz(sz) {
        cloud->points.insert(cloud->points.begin(), ncloud->begin(), ncloud->end());
        cloud->width = (uint32_t)cloud->points.size();
        gpc_config_default(&cfg_);
        cfg_.res = res;
        cfg_.sz = sz;
        cfg_.rgb = 1;  // the reference always fits the RGB field GP next to the height GP (gp_compressor.cpp:163)
    }
    ~gp_compressor() { if (gpu_) gpc_destroy(gpu_); }
    gp_compressor(const gp_compressor&) = delete;
    gp_compressor& operator=(const gp_compressor&) = delete;
    gpc_config& config() { return cfg_; }  // hyper-parameters the reference hard-codes (gp_compressor.cpp:126)
    gpc_handle* handle() { return gpu_; }

    // void save_compressed(const std::string& name), gp_compressor.cpp:21-27 (name is ignored there too)
    void save_compressed(const std::string&) {
        std::cout << "Size of original point cloud: " << cloud->width * cloud->height << std::endl;
        if (!gpu_ && gpc_create(&cfg_, &gpu_) != GPC_OK) throw std::runtime_error("gpc_create failed: no usable sm_100 CUDA device");
        check(gpc_compress(gpu_, cloud->points.data(), (int64_t)cloud->points.size()));
        gpc_sizes s;
        check(gpc_get_sizes(gpu_, &s));
        std::cout << "Number of patches: " << s.n_patches << std::endl;
        gpc_stats st;
        check(gpc_get_stats(gpu_, &st));
        // gp_compressor.cpp:164-174: running mean and max of gps[i].size() over the patches that received points
        uint64_t trained = 0;
        for (int k = 1; k < 33; k++) trained += st.bv_hist[k];
        std::cout << "Mean added: " << (trained ? (double)s.n_bv_total / (double)trained : 0.0) << std::endl;
        std::cout << "Max added: " << st.max_bv << std::endl;
    }
    // pointcloud::Ptr load_compressed(), gp_compressor.cpp:267-386
    pointcloud::Ptr load_compressed() {
        if (!gpu_) throw std::runtime_error("load_compressed before save_compressed");
        gpc_sizes s;
        check(gpc_get_sizes(gpu_, &s));
        pointcloud::Ptr out(new pointcloud());
        out->points.resize((size_t)(s.patch_hi - s.patch_lo) * sz * sz);
        int64_t n = 0;
        check(gpc_decompress(gpu_, out->points.data(), (int64_t)out->points.size(), &n));
        out->resize((size_t)n);
        std::cout << "Size of transformed point cloud: " << out->width * out->height << std::endl;
        return out;
    }
};
