// oracle/ref_frames_prelude.h -- TEST INFRASTRUCTURE.  First part of the translation unit that oracle/ref_build.py feeds to
// g++ on stdin:  this prelude  +  THE REFERENCE'S OWN TEXT of gp_compressor::compute_rotation and ::project_points (the line
// range of /root/reference/src/gp_compressor.cpp from "void gp_compressor::compute_rotation" to the comment after
// project_points, read where it lies and never written anywhere)  +  oracle/ref_frames_post.cpp.
// The class below declares just the members those two functions touch (gp_compressor.h:23-67); Eigen is oracle/eigen_shim.
#include <cmath>
#include <cstdlib>
#include <list>
#include <utility>
#include <vector>

#include <Eigen/Dense>

class gp_compressor {
public:
    typedef std::pair<Eigen::Vector3d, Eigen::Vector3d> point_pair;
    double res;
    int sz;
    std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d> > RGB_means;
    std::vector<std::list<point_pair, Eigen::aligned_allocator<point_pair> > > S;
    std::vector<std::list<point_pair, Eigen::aligned_allocator<point_pair> > > to_be_added;
    Eigen::Array<bool, Eigen::Dynamic, Eigen::Dynamic> W;
    void compute_rotation(Eigen::Matrix3d& R, const Eigen::MatrixXd& points);
    void project_points(Eigen::Vector3d& center, const Eigen::Matrix3d& R, Eigen::MatrixXd& points, const Eigen::MatrixXd& colors,
                        const std::vector<int>& index_search, int* occupied_indices, int i);
};

using namespace Eigen;
using std::fabs;

#line 1 "gp_compressor.cpp (slice)"
