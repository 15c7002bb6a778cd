// pointcloud_compressor.h / pointcloud_decompressor — API shells of the reference's K-SVD dictionary codec
// (/root/reference/src/pointcloud_compressor.h:35-38, pointcloud_decompressor.h:24-25).  The class names and
// signatures are kept so that code naming them still compiles; the K-SVD maths is a different codec and
// out of scope: calling it reports that, loudly.
#pragma once
#include <stdexcept>
#include <string>

#include "pcl_shim.h"

class pointcloud_compressor {
public:
    typedef pcl::PointXYZRGB point;
    typedef pcl::PointCloud<point> pointcloud;
    pointcloud_compressor(pointcloud::ConstPtr, double res, int sz, int dict_size, int words_max, double proj_error,
                          double stop_diff, int RGB_dict_size, int RGB_words_max, double RGB_proj_error, double RGB_stop_diff) {
        (void)res; (void)sz; (void)dict_size; (void)words_max; (void)proj_error; (void)stop_diff;
        (void)RGB_dict_size; (void)RGB_words_max; (void)RGB_proj_error; (void)RGB_stop_diff;
    }
    void save_compressed(const std::string&) {
        throw std::runtime_error("pointcloud_compressor (K-SVD dictionary codec) is outside the B200 GP path; use gp_compressor");
    }
};

class pointcloud_decompressor {
public:
    typedef pcl::PointXYZRGB point;
    typedef pcl::PointCloud<point> pointcloud;
    explicit pointcloud_decompressor(bool display = false) { (void)display; }
    pointcloud::Ptr load_compressed(const std::string&) {
        throw std::runtime_error("pointcloud_decompressor (K-SVD dictionary codec) is outside the B200 GP path; use gp_compressor");
    }
};
