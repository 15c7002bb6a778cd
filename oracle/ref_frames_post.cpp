
#line 1 "oracle/ref_frames_post.cpp"
// oracle/ref_frames_post.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's own compute_rotation and
// project_points (see oracle/ref_frames_prelude.h for how this translation unit is assembled).
extern "C" {

// pts: m points, 3 doubles each.  R9 out, row-major.
void ref_compute_rotation(int m, const double* pts, double* R9) {
    gp_compressor g;
    MatrixXd P(4, m);
    for (int j = 0; j < m; j++) { P(0, j) = pts[3 * j]; P(1, j) = pts[3 * j + 1]; P(2, j) = pts[3 * j + 2]; P(3, j) = 1.0; }
    Matrix3d R;
    g.compute_rotation(R, P);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R9[3 * i + j] = R(i, j);
}

// One call of project_points for one leaf: candidates pts / cols (m x 3 each) with their cloud indices index_search, the
// shared occupied array (updated), the frame R9 (row-major) and the voxel centre (updated: += mn R.col(0), :116).
// Outputs per claimed point, in claim order: candidate position, local coordinates (pt(0) already minus the patch mean,
// pt(1), pt(2)) and centred colour; the patch RGB mean.  Returns the number of claimed points.
int ref_project_points(double res, int sz, int m, const double* pts, const double* cols, const int* index_search, int* occupied,
                       const double* R9, double* center3, int* claimed_m, double* local3, double* colour3, double* rgb_mean3) {
    gp_compressor g;
    g.res = res;
    g.sz = sz;
    g.RGB_means.resize(1);
    g.S.resize(1);
    g.to_be_added.resize(1);
    MatrixXd P(4, m), Cc(3, m);
    for (int j = 0; j < m; j++) {
        P(0, j) = pts[3 * j]; P(1, j) = pts[3 * j + 1]; P(2, j) = pts[3 * j + 2]; P(3, j) = 1.0;
        for (int d = 0; d < 3; d++) Cc(d, j) = cols[3 * j + d];
    }
    Matrix3d R;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R(i, j) = R9[3 * i + j];
    Vector3d center(center3[0], center3[1], center3[2]);
    std::vector<int> idx(index_search, index_search + m);
    std::vector<char> before(m);
    for (int j = 0; j < m; j++) before[j] = (char)occupied[idx[j]];
    g.project_points(center, R, P, Cc, idx, occupied, 0);
    int n = 0;
    for (int j = 0; j < m; j++)
        if (!before[j] && occupied[idx[j]]) {
            // duplicates of one cloud index cannot occur in a candidate list; the j-th newly occupied index is the j-th claim
            claimed_m[n++] = j;
        }
    int k = 0;
    for (const gp_compressor::point_pair& p : g.to_be_added[0]) {
        for (int d = 0; d < 3; d++) { local3[3 * k + d] = p.first(d); colour3[3 * k + d] = p.second(d); }
        k++;
    }
    if (k != n) return -1;
    for (int d = 0; d < 3; d++) { center3[d] = center(d); rgb_mean3[d] = g.RGB_means[0](d); }
    return n;
}

}  // extern "C"
