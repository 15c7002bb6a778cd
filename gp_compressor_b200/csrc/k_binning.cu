// k_binning.cu — K1, K3, K4, K5: lattice replay, voxel keys, leaves, plane-fit rotation,
// point claiming and local-frame projection.
//
// Replaces gp_compressor::project_cloud / compute_rotation / project_points
// (/root/reference/src/gp_compressor.cpp:29-118,177-249) together with the PCL octree
// calls it makes (setInputCloud / addPointsFromInputCloud / leaf iterator /
// generate_voxel_center / radiusSearch; gp_octree.cpp:3-11).  The octree is never built:
// its observable behaviour is restated as data-parallel stages
//   lattice   : PCL's bounding-box growth replayed by "first violating point" searches
//   keys      : integer voxel key -> Morton code (child index x<<2|y<<1|z, gp_octree.cpp:138-140)
//   leaves    : run heads of the sorted codes; visiting order = reverse Morton (PCL 1.7/1.8)
//   rotation  : per leaf, canonical sums over the points of its <= 27 neighbour leaves that
//               lie within the search radius, then a 4x4 factorisation for the plane normal
//   claim     : owner(p) = first leaf in visiting order that has p in its candidate list
//               and accepts it in the box test (the occupied_indices greedy loop, :80-89)
//   group     : stable sort by owner; per-patch mean height / colour and frame outputs
#include <cooperative_groups.h>

#include "gpc_device.cuh"
#include "gpc_internal.h"

namespace cg = cooperative_groups;

namespace gpc {

namespace {

__device__ __forceinline__ bool finite3(float x, float y, float z) {
    return isfinite(x) && isfinite(y) && isfinite(z);
}

__device__ __forceinline__ uint64_t spread3(uint32_t v) {  // bit b -> bit 3b, 21 bits
    uint64_t x = v & 0x1fffffu;
    x = (x | (x << 32)) & 0x1f00000000ffffULL;
    x = (x | (x << 16)) & 0x1f0000ff0000ffULL;
    x = (x | (x << 8)) & 0x100f00f00f00f00fULL;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ULL;
    x = (x | (x << 2)) & 0x1249249249249249ULL;
    return x;
}
__device__ __forceinline__ uint32_t compact3(uint64_t x) {
    x &= 0x1249249249249249ULL;
    x = (x | (x >> 2)) & 0x10c30c30c30c30c3ULL;
    x = (x | (x >> 4)) & 0x100f00f00f00f00fULL;
    x = (x | (x >> 8)) & 0x1f0000ff0000ffULL;
    x = (x | (x >> 16)) & 0x1f00000000ffffULL;
    x = (x | (x >> 32)) & 0x1fffffULL;
    return (uint32_t)x;
}
__device__ __forceinline__ uint64_t morton(uint32_t kx, uint32_t ky, uint32_t kz) {
    return (spread3(kx) << 2) | (spread3(ky) << 1) | spread3(kz);
}

// ---- lattice replay (PCL OctreePointCloud::addPointsFromInputCloud -> adoptBoundingBoxToPoint) -------------
// The bounding box lives in device memory (LatticeState).  Per growth event: (1) every CTA scans its tile of
// the cloud for the first finite point at index >= state.start that violates [mn, mx) (or any finite point while
// the box is undefined); (2) one thread grows the box for that point exactly as PCL does [RECALLED PCL 1.7:
// adoptBoundingBoxToPoint / getKeyBitSize] and advances state.start.  The host only polls state.best.
__device__ void lattice_adopt(const uint8_t* __restrict__ cloud, LatticeState* __restrict__ st) {
    if (st->defined && st->lat.depth > 21) {  // the host rejects this cloud; do not grow further
        st->start = 0x7fffffffffffffffLL;
        st->found = 0;
        return;
    }
    const unsigned long long b = st->best;
    st->found = (b != ~0ull) ? 1 : 0;
    if (b == ~0ull) {  // every remaining point lies inside the box: later searches of the same batch exit at once
        st->start = 0x7fffffffffffffffLL;
        return;
    }
    const float4 pt = *reinterpret_cast<const float4*>(cloud + (int64_t)b * GPC_POINT_BYTES);
    const float p[3] = {pt.x, pt.y, pt.z};
    LatticeDev& L = st->lat;
    const float minValue = 1.1920929e-07f;  // std::numeric_limits<float>::epsilon()
    for (;;) {
        bool lo[3], up[3];
        for (int a = 0; a < 3; a++) { lo[a] = (double)p[a] < L.mn[a]; up[a] = (double)p[a] >= L.mx[a]; }
        if (st->defined && !(lo[0] || lo[1] || lo[2] || up[0] || up[1] || up[2])) break;
        if (st->defined) {
            double side = __dmul_rn((double)(1u << L.depth), L.res);
            for (int a = 0; a < 3; a++)
                if (!up[a]) L.mn[a] = __dadd_rn(L.mn[a], -side);
            L.depth++;
            side = __dadd_rn(__dmul_rn((double)(1u << L.depth), L.res), -(double)minValue);
            for (int a = 0; a < 3; a++) L.mx[a] = __dadd_rn(L.mn[a], side);
            if (L.depth > 30) break;
        } else {
            const double hr = __ddiv_rn(L.res, 2.0);
            for (int a = 0; a < 3; a++) { L.mn[a] = __dadd_rn((double)p[a], -hr); L.mx[a] = __dadd_rn((double)p[a], hr); }
            // getKeyBitSize(): max_voxels = max(keys, 2) = 2 for a box of one voxel, depth = ceil(log2(2) - eps) = 1
            unsigned maxv = 2u;
            for (int a = 0; a < 3; a++) {
                const unsigned mk = __double2uint_rz(__ddiv_rn(__dadd_rn(L.mx[a], -L.mn[a]), L.res));
                maxv = mk > maxv ? mk : maxv;
            }
            unsigned depth = 0;
            while ((1u << depth) < maxv) depth++;  // == ceil(log2(maxv) - float eps) for integer maxv >= 2
            L.depth = depth;
            const double side = __dadd_rn(__dmul_rn((double)(1u << L.depth), L.res), -(double)minValue);
            for (int a = 0; a < 3; a++) {
                const double over = __ddiv_rn(__dadd_rn(side, -__dadd_rn(L.mx[a], -L.mn[a])), 2.0);
                L.mn[a] = __dadd_rn(L.mn[a], -over);
                L.mx[a] = __dadd_rn(L.mx[a], over);
            }
            st->defined = 1;
        }
    }
    st->start = (int64_t)b + 1;
    st->best = ~0ull;
}

// ---- K1: voxel key -> Morton code (PCL genOctreeKeyforPoint) --------------------------------
// Also writes the 16-byte record (x, y, z, rgba bits) of every point when cpt != nullptr: the PointXYZRGB sector is in
// flight anyway, and the later gather into Morton order then reads a half-size array that stays in L2 for room-sized clouds.
__global__ void __launch_bounds__(256) point_keys_kernel(const uint8_t* __restrict__ cloud, int64_t n, LatticeDev lat,
                                                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                         float4* __restrict__ cpt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float4 p = *reinterpret_cast<const float4*>(cloud + i * GPC_POINT_BYTES);
        if (cpt) {
            float4 c = p;
            c.w = __uint_as_float(*reinterpret_cast<const uint32_t*>(cloud + i * GPC_POINT_BYTES + 16));  // b,g,r,a bytes
            cpt[i] = c;
        }
        uint64_t code = 1ull << (3 * lat.depth);  // non-finite points sort behind every voxel
        if (finite3(p.x, p.y, p.z)) {
            uint32_t kx = __double2uint_rz(__ddiv_rn(__dadd_rn((double)p.x, -lat.mn[0]), lat.res));
            uint32_t ky = __double2uint_rz(__ddiv_rn(__dadd_rn((double)p.y, -lat.mn[1]), lat.res));
            uint32_t kz = __double2uint_rz(__ddiv_rn(__dadd_rn((double)p.z, -lat.mn[2]), lat.res));
            code = morton(kx, ky, kz);
        }
        keys[i] = code;
        vals[i] = (uint32_t)i;
    }
}

// K3 fused: valid count, leaf heads, leaf numbering (chained scan with decoupled look-back over the blocks), leaf tables and
// the gather of the sorted points, in ONE pass over the sorted keys.  Replaces count_valid + mark_heads + a three-kernel
// int64 scan + fill_leaves (and one host round trip: n_valid and the leaf count come back together).
//   counts[0] = n_valid (keys below `invalid`, i.e. finite points), counts[1] = number of leaves
constexpr int LF_T = 256, LF_ITEMS = 8, LF_TILE = LF_T * LF_ITEMS;
__global__ void __launch_bounds__(LF_T) leaves_fused_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n,
                                                            uint64_t invalid, const float4* __restrict__ cpt,
                                                            int32_t* __restrict__ leaf_of, int64_t* __restrict__ leaf_start,
                                                            uint64_t* __restrict__ leaf_code, float4* __restrict__ spt,
                                                            unsigned long long* __restrict__ counts,
                                                            unsigned long long* __restrict__ status, unsigned int* __restrict__ counter) {
    __shared__ unsigned int tile_s;
    __shared__ unsigned int wsum[LF_T / 32];
    __shared__ unsigned long long prefix_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) tile_s = atomicAdd(counter, 1u);   // blocks start in the order of their tile numbers
    __syncthreads();
    const int64_t tile = tile_s;
    const int64_t s0 = tile * LF_TILE + (int64_t)t * LF_ITEMS;
    uint64_t k[LF_ITEMS];
    uint64_t prev = 0;
    bool have_prev = false;
    if (s0 > 0 && s0 - 1 < n) { prev = keys[s0 - 1]; have_prev = true; }
#pragma unroll
    for (int i = 0; i < LF_ITEMS; i++) k[i] = s0 + i < n ? keys[s0 + i] : ~0ull;
    // the point gathers do not depend on the scan: issue them now
    float4 pt[LF_ITEMS];
#pragma unroll
    for (int i = 0; i < LF_ITEMS; i++) {
        pt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s0 + i < n && k[i] < invalid) {
            pt[i] = cpt[vals[s0 + i]];
        }
    }
    unsigned int head[LF_ITEMS], c = 0, nvalid = 0;
#pragma unroll
    for (int i = 0; i < LF_ITEMS; i++) {
        const bool valid = s0 + i < n && k[i] < invalid;
        const uint64_t before = i ? k[i - 1] : prev;
        head[i] = (valid && (!(i || have_prev) || k[i] != before)) ? 1u : 0u;
        c += head[i];
        nvalid += valid ? 1u : 0u;
    }
    // block scan of the head counts (and the block's valid count)
    unsigned int x = c, v = nvalid;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 31) wsum[w] = x;
    if (lane == 0 && v) atomicAdd(counts + 0, (unsigned long long)v);
    __syncthreads();
    unsigned int wb = 0, total = 0;
#pragma unroll
    for (int ww = 0; ww < LF_T / 32; ww++) {
        if (ww < w) wb += wsum[ww];
        total += wsum[ww];
    }
    // chained scan over the tiles: status = flag (2 bits) | count.  Warp 0 looks back 32 tiles at a time.
    constexpr unsigned long long AGG = 1ull << 62, INCL = 2ull << 62, MASK = (1ull << 62) - 1;
    if (w == 0) {
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(&status[0]) = INCL | total;
        } else {
            if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(&status[tile]) = AGG | total;
            for (int64_t j = tile - 1;; j -= 32) {
                const int64_t idx = j - lane;
                unsigned long long sw = INCL;   // before the first tile: an inclusive prefix of zero
                if (idx >= 0) {
                    do { sw = *reinterpret_cast<volatile const unsigned long long*>(&status[idx]); } while ((sw >> 62) == 0ull);
                }
                const unsigned incl = __ballot_sync(0xffffffffu, (sw & INCL) != 0ull);
                const int stop = incl ? (__ffs(incl) - 1) : 31;            // nearest predecessor that holds an inclusive prefix
                unsigned long long v = (lane <= stop) ? (sw & MASK) : 0ull;
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                excl += v;
                if (incl) break;
            }
            if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(&status[tile]) = INCL | (excl + total);
        }
        if (lane == 0) {
            prefix_s = excl;
            if ((tile + 1) * LF_TILE >= n) counts[1] = excl + total;   // the last tile: number of leaves
        }
    }
    __syncthreads();
    int64_t a = (int64_t)prefix_s + wb + (x - c) - 1;   // leaf of the element before this thread's first one
#pragma unroll
    for (int i = 0; i < LF_ITEMS; i++) {
        const int64_t s = s0 + i;
        const bool valid = s < n && k[i] < invalid;
        if (valid) {
            a += head[i];
            leaf_of[s] = (int32_t)a;
            if (head[i]) { leaf_start[a] = s; leaf_code[a] = k[i]; }
            spt[s] = pt[i];
            // end of the valid keys: leaf_start[P] = n_valid
            const bool last = (s + 1 == n) || (i + 1 < LF_ITEMS ? !(k[i + 1] < invalid) : !(keys[s + 1] < invalid));
            if (last) leaf_start[a + 1] = s + 1;
        }
    }
}

// ---- neighbour table + voxel centres, one warp per leaf ------------------------------------
__global__ void __launch_bounds__(256) leaf_neighbours_kernel(const uint64_t* __restrict__ leaf_code, int64_t P, LatticeDev lat,
                                                              int32_t* __restrict__ nbr, int32_t* __restrict__ nnbr,
                                                              float* __restrict__ center) {
    const int64_t a = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (a >= P) return;
    const uint64_t code = leaf_code[a];
    const uint32_t kx = compact3(code >> 2), ky = compact3(code >> 1), kz = compact3(code);
    const int64_t kmax = ((int64_t)1 << lat.depth) - 1;
    int32_t id = -1;
    if (lane < 27) {
        const int64_t x = (int64_t)kx + (lane / 9) - 1, y = (int64_t)ky + ((lane / 3) % 3) - 1, z = (int64_t)kz + (lane % 3) - 1;
        if (x >= 0 && y >= 0 && z >= 0 && x <= kmax && y <= kmax && z <= kmax) {
            const uint64_t nc = morton((uint32_t)x, (uint32_t)y, (uint32_t)z);
            int64_t lo = 0, hi = P;  // first index with leaf_code >= nc
            while (lo < hi) {
                int64_t mid = (lo + hi) >> 1;
                if (leaf_code[mid] < nc) lo = mid + 1; else hi = mid;
            }
            if (lo < P && leaf_code[lo] == nc) id = (int32_t)lo;
        }
    }
    // ascending leaf index = ascending Morton = PCL radiusSearch child order 0..7
    int rank = 0;
    for (int l = 0; l < 27; l++) {
        int32_t o = __shfl_sync(0xffffffffu, id, l);
        if (o >= 0 && o < id) rank++;
    }
    const int cnt = __popc(__ballot_sync(0xffffffffu, id >= 0));
    if (id >= 0) nbr[a * 27 + rank] = id;
    if (lane >= cnt && lane < 27) nbr[a * 27 + lane] = -1;
    if (lane == 0) nnbr[a] = cnt;
    if (lane < 3) {
        const uint32_t k = lane == 0 ? kx : (lane == 1 ? ky : kz);
        // genLeafNodeCenterFromOctreeKey: float((double(k) + 0.5f) * res + min)
        center[a * 3 + lane] = (float)__dadd_rn(__dmul_rn(__dadd_rn((double)k, 0.5), lat.res), lat.mn[lane]);
    }
}

// ---- plane fit: smallest right singular vector of A = [x y z 1] from centred sums ----------
// G_c = A_c' A_c (A_c = [p - c, 1]); G_c = L D L'; M = sqrt(D) L' S with S = [[I,0],[c',1]];
// one-sided Jacobi on the 4x4 M.  Same operation sequence as the test oracle.
__device__ void smallest_right_singular_vector(const double* g, const double* c, double* v) {
    double G[4][4];
    G[0][0] = g[0]; G[0][1] = g[1]; G[0][2] = g[2]; G[0][3] = g[6];
    G[1][1] = g[3]; G[1][2] = g[4]; G[1][3] = g[7];
    G[2][2] = g[5]; G[2][3] = g[8];
    G[3][3] = g[9];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (j < i) G[i][j] = G[j][i];
    double Lm[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    double D[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        double d = G[j][j];
#pragma unroll
        for (int t = 0; t < 4; t++)
            if (t < j) d = d - (Lm[j][t] * Lm[j][t]) * D[t];
        if (!(d > 0.0)) d = 0.0;
        D[j] = d;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i > j) {
                double a = G[i][j];
#pragma unroll
                for (int t = 0; t < 4; t++)
                    if (t < j) a = a - (Lm[i][t] * Lm[j][t]) * D[t];
                Lm[i][j] = (d > 0.0) ? a / d : 0.0;
            }
    }
    double M[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double sd = sqrt(D[i]);
        double u[4];
#pragma unroll
        for (int j = 0; j < 4; j++) u[j] = (j >= i) ? sd * Lm[j][i] : 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++) M[i][j] = fma(u[3], c[j], u[j]);
        M[i][3] = u[3];
    }
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 30; sweep++) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
            for (int q = 1; q < 4; q++) {
                if (q <= p) continue;
                double a = 0, b = 0, gpq = 0;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    a = fma(M[t][p], M[t][p], a);
                    b = fma(M[t][q], M[t][q], b);
                    gpq = fma(M[t][p], M[t][q], gpq);
                }
                if (gpq == 0.0) continue;
                if (gpq * gpq <= 0x1p-106 * (a * b)) continue;
                rotated = true;
                double zeta = (b - a) / (2.0 * gpq);
                double t = 1.0 / (fabs(zeta) + sqrt(fma(zeta, zeta, 1.0)));
                if (zeta < 0.0) t = -t;
                double cs = 1.0 / sqrt(fma(t, t, 1.0));
                double sn = cs * t;
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    double mp = M[r][p], mq = M[r][q];
                    M[r][p] = cs * mp - sn * mq;
                    M[r][q] = sn * mp + cs * mq;
                    double vp = V[r][p], vq = V[r][q];
                    V[r][p] = cs * vp - sn * vq;
                    V[r][q] = sn * vp + cs * vq;
                }
            }
        if (!rotated) break;
    }
    int best = 0;
    double bestn = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        double nn = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) nn = fma(M[t][j], M[t][j], nn);
        if (j == 0 || nn < bestn) { bestn = nn; best = j; }
    }
#pragma unroll
    for (int r = 0; r < 4; r++) v[r] = (best == 0) ? V[r][0] : (best == 1) ? V[r][1] : (best == 2) ? V[r][2] : V[r][3];
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void normalize3(double* v) {
    double n2 = v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]);
    double n = sqrt(n2);
    v[0] = v[0] / n; v[1] = v[1] / n; v[2] = v[2] / n;
}
// gp_compressor.cpp:38-63
__device__ void rotation_from_normal(double* nrm, double* R) {
    normalize3(nrm);
    double ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, ez[3] = {0, 0, 1};
    double c1[3];
    const double ax = fabs(nrm[0]), ay = fabs(nrm[1]), az = fabs(nrm[2]);
    if (ax > ay && ax > az) {
        if (nrm[0] < 0) { nrm[0] *= -1; nrm[1] *= -1; nrm[2] *= -1; }
        cross3(ez, nrm, c1);
    } else if (ay > ax && ay > az) {
        if (nrm[1] < 0) { nrm[0] *= -1; nrm[1] *= -1; nrm[2] *= -1; }
        cross3(ex, nrm, c1);
    } else {
        if (nrm[2] < 0) { nrm[0] *= -1; nrm[1] *= -1; nrm[2] *= -1; }
        cross3(ey, nrm, c1);
    }
    normalize3(c1);
    double c2[3];
    cross3(nrm, c1, c2);
#pragma unroll
    for (int r = 0; r < 3; r++) { R[r * 3 + 0] = nrm[r]; R[r * 3 + 1] = c1[r]; R[r * 3 + 2] = c2[r]; }
}

// Eigen Quaterniond <- Matrix3d (gp_compressor.cpp:240), q = (x, y, z, w)
__device__ void rot_to_quat(const double* R, double* q) {
#define MR(r, c) R[(r) * 3 + (c)]
    double t = (MR(0, 0) + MR(1, 1)) + MR(2, 2);
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (MR(2, 1) - MR(1, 2)) * t;
        q[1] = (MR(0, 2) - MR(2, 0)) * t;
        q[2] = (MR(1, 0) - MR(0, 1)) * t;
    } else {
        int i = 0;
        if (MR(1, 1) > MR(0, 0)) i = 1;
        if (MR(2, 2) > MR(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(((MR(i, i) - MR(j, j)) - MR(k, k)) + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (MR(k, j) - MR(j, k)) * t;
        q[j] = (MR(j, i) + MR(i, j)) * t;
        q[k] = (MR(k, i) + MR(i, k)) * t;
    }
#undef MR
}

// ---- K4: rotation per leaf.  (a) one warp per leaf: canonical sums over the candidate points;
// (b) one thread per leaf: 4x4 factorisation and frame (the serial part runs once, not 32 times) ----------
__global__ void __launch_bounds__(256) leaf_sums_kernel(const float4* __restrict__ spt, const int64_t* __restrict__ leaf_start,
                                                        const int32_t* __restrict__ nbr, const int32_t* __restrict__ nnbr,
                                                        const float* __restrict__ center, int64_t P, double r2,
                                                        double* __restrict__ sums) {
    const int64_t a = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (a >= P) return;
    const float cx = center[a * 3], cy = center[a * 3 + 1], cz = center[a * 3 + 2];
    double s[10];
#pragma unroll
    for (int q = 0; q < 10; q++) s[q] = 0.0;
    const int cnt = nnbr[a];
    for (int t = 0; t < cnt; t++) {
        const int32_t v = nbr[a * 27 + t];
        const int64_t lo = leaf_start[v], hi = leaf_start[v + 1];
        for (int64_t i = lo + lane; i < hi; i += 32) {
            const float4 p = spt[i];
            // PCL pointSquaredDist: float (p - c).squaredNorm(); accept iff <= radius^2 (double)
            const float fx = __fsub_rn(p.x, cx), fy = __fsub_rn(p.y, cy), fz = __fsub_rn(p.z, cz);
            const float d2 = __fadd_rn(__fmul_rn(fx, fx), __fadd_rn(__fmul_rn(fy, fy), __fmul_rn(fz, fz)));
            if ((double)d2 > r2) continue;
            const double ux = __dadd_rn((double)p.x, -(double)cx), uy = __dadd_rn((double)p.y, -(double)cy),
                         uz = __dadd_rn((double)p.z, -(double)cz);
            s[0] = fma(ux, ux, s[0]); s[1] = fma(ux, uy, s[1]); s[2] = fma(ux, uz, s[2]);
            s[3] = fma(uy, uy, s[3]); s[4] = fma(uy, uz, s[4]); s[5] = fma(uz, uz, s[5]);
            s[6] = __dadd_rn(s[6], ux); s[7] = __dadd_rn(s[7], uy); s[8] = __dadd_rn(s[8], uz);
            s[9] = __dadd_rn(s[9], 1.0);
        }
    }
#pragma unroll
    for (int q = 0; q < 10; q++) s[q] = butterfly32(s[q]);
    if (lane < 10) {
        double val = s[0];
#pragma unroll
        for (int q = 1; q < 10; q++)
            if (lane == q) val = s[q];
        sums[(int64_t)lane * P + a] = val;  // SoA: the solve kernel reads coalesced
    }
}

__global__ void __launch_bounds__(128) leaf_solve_kernel(const double* __restrict__ sums, const float* __restrict__ center, int64_t P,
                                                         double* __restrict__ Rm, int32_t* __restrict__ ncand) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= P) return;
    double s[10];
#pragma unroll
    for (int q = 0; q < 10; q++) s[q] = sums[(int64_t)q * P + a];
    const int m = (int)s[9];
    double R[9];
    if (m < 4) {  // gp_compressor.cpp:31-34
#pragma unroll
        for (int q = 0; q < 9; q++) R[q] = (q % 4 == 0) ? 1.0 : 0.0;
    } else {
        const double cd[3] = {(double)center[a * 3], (double)center[a * 3 + 1], (double)center[a * 3 + 2]};
        double v[4];
        smallest_right_singular_vector(s, cd, v);
        double nrm[3] = {v[0], v[1], v[2]};
        rotation_from_normal(nrm, R);
    }
#pragma unroll
    for (int q = 0; q < 9; q++) Rm[a * 9 + q] = R[q];
    ncand[a] = m;
}

// ---- K5: owner and local coordinates, one thread per (Morton-sorted) point --------------------
__global__ void __launch_bounds__(256) claim_kernel(const float4* __restrict__ spt, const int32_t* __restrict__ leaf_of,
                                                    const int32_t* __restrict__ nbr, const int32_t* __restrict__ nnbr,
                                                    const float* __restrict__ center, const double* __restrict__ Rm,
                                                    const int32_t* __restrict__ ncand, int64_t n_valid, int64_t P, double r2,
                                                    double half, int leaf_order, uint64_t* __restrict__ okey,
                                                    uint32_t* __restrict__ oval, double* __restrict__ pt0,
                                                    double* __restrict__ pt1, double* __restrict__ pt2) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_valid) return;
    const float4 p = spt[s];
    const int32_t a = leaf_of[s];
    const int cnt = nnbr[a];
    uint64_t key = (uint64_t)P;  // unclaimed points sort behind every patch
    double q0 = 0, q1 = 0, q2 = 0;
    for (int t = 0; t < cnt; t++) {
        const int32_t v = (leaf_order == 0) ? nbr[(int64_t)a * 27 + (cnt - 1 - t)] : nbr[(int64_t)a * 27 + t];
        const float cx = center[(int64_t)v * 3], cy = center[(int64_t)v * 3 + 1], cz = center[(int64_t)v * 3 + 2];
        const float fx = __fsub_rn(p.x, cx), fy = __fsub_rn(p.y, cy), fz = __fsub_rn(p.z, cz);
        const float d2 = __fadd_rn(__fmul_rn(fx, fx), __fadd_rn(__fmul_rn(fy, fy), __fmul_rn(fz, fz)));
        if ((double)d2 > r2) continue;
        if (ncand[v] == 0) continue;
        const double* R = Rm + (int64_t)v * 9;
        const double d0 = __dadd_rn((double)p.x, -(double)cx), d1 = __dadd_rn((double)p.y, -(double)cy),
                     d2d = __dadd_rn((double)p.z, -(double)cz);
        // pt = R' (p - centre), gp_compressor.cpp:84
        const double a0 = __dadd_rn(__dadd_rn(__dmul_rn(R[0], d0), __dmul_rn(R[3], d1)), __dmul_rn(R[6], d2d));
        const double a1 = __dadd_rn(__dadd_rn(__dmul_rn(R[1], d0), __dmul_rn(R[4], d1)), __dmul_rn(R[7], d2d));
        const double a2 = __dadd_rn(__dadd_rn(__dmul_rn(R[2], d0), __dmul_rn(R[5], d1)), __dmul_rn(R[8], d2d));
        if (a1 > half || a1 < -half || a2 > half || a2 < -half) continue;  // :85
        key = (uint64_t)((leaf_order == 0) ? (P - 1 - v) : v);
        q0 = a0; q1 = a1; q2 = a2;
        break;
    }
    okey[s] = key;
    oval[s] = (uint32_t)s;
    pt0[s] = q0; pt1[s] = q1; pt2[s] = q2;
}

// patch_off[i] = first sorted position whose owner key >= i (i = 0..P)
__global__ void patch_bounds_kernel(const uint64_t* __restrict__ okey, int64_t n_valid, int64_t P, int64_t* __restrict__ patch_off) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > P) return;
    int64_t lo = 0, hi = n_valid;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (okey[mid] < (uint64_t)i) lo = mid + 1; else hi = mid;
    }
    patch_off[i] = lo;
}

__global__ void group_gather_kernel(const uint64_t* __restrict__ okey, const uint32_t* __restrict__ oval,
                                    const uint32_t* __restrict__ sorted_idx, const float4* __restrict__ spt,
                                    const double* __restrict__ pt0, const double* __restrict__ pt1,
                                    const double* __restrict__ pt2, int64_t n_claimed, int32_t* __restrict__ st_idx,
                                    double* __restrict__ h, double* __restrict__ x1, double* __restrict__ x2,
                                    uint32_t* __restrict__ rgb, int32_t* __restrict__ owner) {
    int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_claimed) return;
    const uint32_t s = oval[d];
    const uint32_t orig = sorted_idx[s];
    st_idx[d] = (int32_t)orig;
    h[d] = pt0[s];
    x1[d] = pt1[s];
    x2[d] = pt2[s];
    rgb[d] = __float_as_uint(spt[s].w);
    owner[orig] = (int32_t)okey[d];
}

// per patch (gp_index i): mean height, colour mean, y = height - mean, frame outputs; warp per patch
__global__ void __launch_bounds__(128) patch_frames_kernel(const int64_t* __restrict__ patch_off, int64_t P, int leaf_order,
                                                           const double* __restrict__ h, const uint32_t* __restrict__ rgb,
                                                           const uint64_t* __restrict__ leaf_code_a, const float* __restrict__ center_a,
                                                           const double* __restrict__ Rm_a, const int32_t* __restrict__ ncand_a,
                                                           double* __restrict__ y, uint64_t* __restrict__ code,
                                                           float* __restrict__ center, int32_t* __restrict__ ncand,
                                                           double* __restrict__ Rm, double* __restrict__ quat,
                                                           double* __restrict__ mean, double* __restrict__ rgbmean) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= P) return;
    const int64_t a = (leaf_order == 0) ? (P - 1 - i) : i;
    const int64_t lo = patch_off[i], hi = patch_off[i + 1];
    double sh = 0.0, sr = 0.0, sg = 0.0, sb = 0.0;
    for (int64_t d = lo + lane; d < hi; d += 32) {
        sh = __dadd_rn(sh, h[d]);
        const uint32_t c = rgb[d];
        sr = __dadd_rn(sr, (double)((c >> 16) & 255u));
        sg = __dadd_rn(sg, (double)((c >> 8) & 255u));
        sb = __dadd_rn(sb, (double)(c & 255u));
    }
    sh = butterfly32(sh); sr = butterfly32(sr); sg = butterfly32(sg); sb = butterfly32(sb);
    const double cnt = (double)(hi - lo);
    const double mn = sh / cnt;  // NaN for an empty patch, as in the reference (:101)
    for (int64_t d = lo + lane; d < hi; d += 32) y[d] = __dadd_rn(h[d], -mn);
    if (lane == 0) {
        code[i] = leaf_code_a[a];
        ncand[i] = ncand_a[a];
        double R[9];
        for (int q = 0; q < 9; q++) { R[q] = Rm_a[a * 9 + q]; Rm[i * 9 + q] = R[q]; }
        double qq[4];
        rot_to_quat(R, qq);
        for (int q = 0; q < 4; q++) quat[i * 4 + q] = qq[q];
        for (int dd = 0; dd < 3; dd++) {
            const float c = center_a[a * 3 + dd];
            center[i * 3 + dd] = c;
            mean[i * 3 + dd] = __dadd_rn((double)c, __dmul_rn(mn, R[dd * 3]));  // centre += mn*R.col(0), :116
        }
        rgbmean[i * 3 + 0] = sr / cnt; rgbmean[i * 3 + 1] = sg / cnt; rgbmean[i * 3 + 2] = sb / cnt;
    }
}

}  // namespace

// The whole replay in ONE cooperative launch: search for the first violator behind state.start (every CTA walks tiles of
// 4096 points in index order and stops behind the best hit so far), grid barrier, one thread grows the box, grid barrier,
// until no point violates the box.  A cloud costs about one pass over its points plus two grid barriers per growth event
// (~depth + 2 events); the host reads the final state once.
__global__ void __launch_bounds__(256) lattice_replay_kernel(const uint8_t* __restrict__ cloud, int64_t n, LatticeState* st) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned long long sbest;
    __shared__ int sskip;
    // Growth events cluster at the head of the cloud (the k-th event needs ~2^k points to show up): CTA 0 replays the first
    // LOCAL_W points on its own, event by event, without grid barriers; the grid-wide search handles the few events behind.
    constexpr int64_t LOCAL_W = 8192;
    if (blockIdx.x == 0) {
        const int64_t end = n < LOCAL_W ? n : LOCAL_W;
        for (;;) {
            const int64_t pos = *reinterpret_cast<volatile int64_t*>(&st->start);
            if (pos >= end) break;
            const int defined = *reinterpret_cast<volatile int32_t*>(&st->defined);
            const volatile double* vmn = st->lat.mn;
            const volatile double* vmx = st->lat.mx;
            const double mn0 = vmn[0], mn1 = vmn[1], mn2 = vmn[2], mx0 = vmx[0], mx1 = vmx[1], mx2 = vmx[2];
            if (threadIdx.x == 0) sbest = ~0ull;
            __syncthreads();
            for (int64_t i = pos + threadIdx.x; i < end; i += 256) {
                const float4 p = *reinterpret_cast<const float4*>(cloud + i * GPC_POINT_BYTES);
                if (!finite3(p.x, p.y, p.z)) continue;
                bool hit = !defined;
                if (defined) {
                    const double x = p.x, y = p.y, z = p.z;
                    hit = x < mn0 || y < mn1 || z < mn2 || x >= mx0 || y >= mx1 || z >= mx2;
                }
                if (hit) { atomicMin(&sbest, (unsigned long long)i); break; }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                if (sbest == ~0ull) {
                    st->start = end;            // the window is clean
                } else {
                    st->best = sbest;
                    lattice_adopt(cloud, st);   // grows the box, start = best + 1, best = none
                }
                __threadfence();
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) { st->best = ~0ull; __threadfence(); }
    }
    grid.sync();
    for (;;) {
        const int64_t start = *reinterpret_cast<volatile int64_t*>(&st->start);
        if (start >= n) break;   // nothing left (or the octree is too deep: the host rejects the cloud)
        const int defined = *reinterpret_cast<volatile int32_t*>(&st->defined);
        const volatile double* vmn = st->lat.mn;   // written by the adopting thread between two grid barriers
        const volatile double* vmx = st->lat.mx;
        const double mn0 = vmn[0], mn1 = vmn[1], mn2 = vmn[2];
        const double mx0 = vmx[0], mx1 = vmx[1], mx2 = vmx[2];
        for (int64_t tile = (int64_t)blockIdx.x * 4096; start + tile < n; tile += (int64_t)gridDim.x * 4096) {
            // one thread decides for the CTA (best changes under our feet: every thread must take the same branch)
            if (threadIdx.x == 0) {
                sskip = (unsigned long long)(start + tile) >= *reinterpret_cast<volatile unsigned long long*>(&st->best);  // an earlier hit exists
                sbest = ~0ull;
            }
            __syncthreads();
            if (sskip) break;
            unsigned long long mine = ~0ull;
            for (int r = 0; r < 16; r++) {
                const int64_t i = start + tile + r * 256 + threadIdx.x;
                if (i >= n) break;
                const float4 p = *reinterpret_cast<const float4*>(cloud + i * GPC_POINT_BYTES);
                if (!finite3(p.x, p.y, p.z)) continue;
                bool hit = !defined;
                if (defined) {
                    const double x = p.x, y = p.y, z = p.z;
                    hit = x < mn0 || y < mn1 || z < mn2 || x >= mx0 || y >= mx1 || z >= mx2;
                }
                if (hit) { mine = (unsigned long long)i; break; }
            }
            if (mine != ~0ull) atomicMin(&sbest, mine);
            __syncthreads();
            const bool found = sbest != ~0ull;
            if (threadIdx.x == 0 && found) atomicMin(&st->best, sbest);
            __syncthreads();
            if (found) break;   // later tiles of this CTA lie behind the hit
        }
        __threadfence();
        grid.sync();
        if (blockIdx.x == 0 && threadIdx.x == 0) { lattice_adopt(cloud, st); __threadfence(); }
        grid.sync();
        if (!*reinterpret_cast<volatile int32_t*>(&st->found)) break;
    }
}

// the lattice replay of one cloud; the final LatticeState is read by the caller
cudaError_t launch_lattice_replay(const uint8_t* cloud, int64_t n, LatticeState* st, cudaStream_t s) {
    static int max_blocks = 0;
    if (!max_blocks) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lattice_replay_kernel, 256, 0);
        max_blocks = sms * std::max(1, std::min(per_sm, 4));
    }
    const int64_t tiles = std::max<int64_t>(1, (n + 4095) / 4096);
    const unsigned grid = (unsigned)std::min<int64_t>(tiles, max_blocks);
    void* args[] = {(void*)&cloud, (void*)&n, (void*)&st};
    g_launches++;
    return cudaLaunchCooperativeKernel((const void*)lattice_replay_kernel, dim3(grid), dim3(256), args, 0, s);
}

// ---- sharded binning (strong scaling of one cloud) -------------------------------------------------
// Every rank holds the whole cloud.  The Morton keys are cut into `count` contiguous ranges of about equal point counts
// (splitters = quantiles of a sorted key sample, rounded down to blocks of 8 x 8 x 8 voxels; identical on every rank:
// integer work on identical input); a rank bins only the points within HALO = 3 voxels of its own range: a patch's
// claimed set depends on the frames of the leaves within two rings, and those on the points within three.
constexpr int SHARD_HALO = 3;
constexpr int SHARD_BLOCK_BITS = 9;   // 8^3 voxels: wider than the 7-voxel halo box, so its 8 corners meet every block it touches

__global__ void __launch_bounds__(256) shard_sample_kernel(const uint64_t* __restrict__ keys, int64_t n, int64_t stride, int64_t m,
                                                           uint64_t* __restrict__ sample, uint32_t* __restrict__ dummy) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    sample[i] = keys[i * stride];
    dummy[i] = 0u;
}

// Weight of every sorted sample for the range cut.  Measured on C5 (profiles/r2_strong_scaling.md): cutting at equal COST
// (weight growing with the points of the sample's voxel, estimated from the run of equal keys) moved a third of the points
// off the ranks that own the dense core and did not shorten their fit -- they are bound by the sequential recursion of
// their ~20 largest patches (2 000+ points, 40+ basis vectors), not by throughput -- while the halo of their neighbours
// grew.  So the ranges are cut at equal point counts: weight 1 per finite sample.
__global__ void __launch_bounds__(256) shard_weights_kernel(const uint64_t* __restrict__ sorted, int64_t m, int64_t stride, uint64_t invalid,
                                                            int64_t* __restrict__ weight) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    weight[i] = sorted[i] < invalid ? 1 : 0;
}
// range[0..1] = [klo, khi): the key range of `rank` in the visiting order (leaf_order 0 visits keys in descending order).
// prefix = exclusive scan of the weights (m + 1 entries): splitter j sits where the cumulative cost reaches j / count of the total.
__global__ void shard_splitters_kernel(const uint64_t* __restrict__ sorted, const int64_t* __restrict__ prefix, int64_t m, int depth3, int rank,
                                       int count, int leaf_order, uint64_t* __restrict__ range) {
    const uint64_t invalid = 1ull << depth3;
    const int64_t total = prefix[m];
    const int ja = leaf_order == 0 ? count - 1 - rank : rank;   // ascending index of the rank's range
    uint64_t k[2];
    for (int e = 0; e < 2; e++) {
        const int j = ja + e;
        if (j <= 0) k[e] = 0;
        else if (j >= count || total == 0) k[e] = invalid;
        else {
            const int64_t target = (int64_t)(((__int128)total * j) / count);
            int64_t lo = 0, hi = m;   // first sample whose prefix reaches the target
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (prefix[mid] < target) lo = mid + 1; else hi = mid; }
            const int64_t at = lo < m ? lo : m - 1;
            k[e] = sorted[at] >= invalid ? invalid : (sorted[at] >> SHARD_BLOCK_BITS) << SHARD_BLOCK_BITS;
        }
    }
    range[0] = k[0];
    range[1] = k[1] < k[0] ? k[0] : k[1];
}

// flags[i] = 1 if point i lies within SHARD_HALO voxels of a block of the rank's key range
__global__ void __launch_bounds__(256) shard_select_kernel(const uint64_t* __restrict__ keys, int64_t n, int depth,
                                                           const uint64_t* __restrict__ range, int64_t* __restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = keys[i], klo = range[0], khi = range[1];
    int sel = 0;
    if (!(k >> (3 * depth))) {
        const int64_t kmax = (1ll << depth) - 1;
        const int64_t vx = compact3(k >> 2), vy = compact3(k >> 1), vz = compact3(k);
#pragma unroll
        for (int c = 0; c < 8; c++) {
            int64_t x = vx + ((c & 4) ? SHARD_HALO : -SHARD_HALO), y = vy + ((c & 2) ? SHARD_HALO : -SHARD_HALO),
                    z = vz + ((c & 1) ? SHARD_HALO : -SHARD_HALO);
            x = x < 0 ? 0 : (x > kmax ? kmax : x);
            y = y < 0 ? 0 : (y > kmax ? kmax : y);
            z = z < 0 ? 0 : (z > kmax ? kmax : z);
            const uint64_t kc = morton((uint32_t)x, (uint32_t)y, (uint32_t)z);
            sel |= (kc >= klo && kc < khi) ? 1 : 0;
        }
    }
    flags[i] = sel;
}

__global__ void __launch_bounds__(256) shard_compact_kernel(const uint8_t* __restrict__ cloud, const int64_t* __restrict__ ex, int64_t n,
                                                            uint8_t* __restrict__ sel_cloud, int32_t* __restrict__ sel_idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t d = ex[i];
    if (ex[i + 1] == d) return;
    const float4* src = reinterpret_cast<const float4*>(cloud + i * GPC_POINT_BYTES);
    float4* dst = reinterpret_cast<float4*>(sel_cloud + d * GPC_POINT_BYTES);
    dst[0] = src[0];
    dst[1] = src[1];
    sel_idx[d] = (int32_t)i;
}

// Sharded binning, one pass over the keys of the whole cloud: halo test, chained scan (decoupled look-back over the tiles)
// and compaction of the selected PointXYZRGB records.  Replaces shard_select + a three-kernel int64 scan + shard_compact.
// n_sel_out[0] = number of selected points.
constexpr int SS_T = 256, SS_ITEMS = 8, SS_TILE = SS_T * SS_ITEMS;
__global__ void __launch_bounds__(SS_T) shard_select_compact_kernel(const uint64_t* __restrict__ keys, int64_t n, int depth,
                                                                    const uint64_t* __restrict__ range, const uint8_t* __restrict__ cloud,
                                                                    uint8_t* __restrict__ sel_cloud, int32_t* __restrict__ sel_idx,
                                                                    unsigned long long* __restrict__ n_sel_out,
                                                                    unsigned long long* __restrict__ status, unsigned int* __restrict__ counter) {
    __shared__ unsigned int tile_s;
    __shared__ unsigned int wsum[SS_T / 32];
    __shared__ unsigned long long prefix_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) tile_s = atomicAdd(counter, 1u);
    __syncthreads();
    const int64_t tile = tile_s;
    const uint64_t klo = range[0], khi = range[1];
    const int64_t kmax = (1ll << depth) - 1;
    // thread t owns the points tile * SS_TILE + r * SS_T + t (coalesced key loads); its rank inside the tile follows the
    // same order: round-major, so the compacted order is the input order
    unsigned int sel = 0;   // bit r: point of round r is selected
#pragma unroll
    for (int r = 0; r < SS_ITEMS; r++) {
        const int64_t i = tile * SS_TILE + (int64_t)r * SS_T + t;
        if (i < n) {
            const uint64_t k = keys[i];
            if (!(k >> (3 * depth))) {
                const int64_t vx = compact3(k >> 2), vy = compact3(k >> 1), vz = compact3(k);
                int hit = 0;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    int64_t x = vx + ((c & 4) ? SHARD_HALO : -SHARD_HALO), y = vy + ((c & 2) ? SHARD_HALO : -SHARD_HALO),
                            z = vz + ((c & 1) ? SHARD_HALO : -SHARD_HALO);
                    x = x < 0 ? 0 : (x > kmax ? kmax : x);
                    y = y < 0 ? 0 : (y > kmax ? kmax : y);
                    z = z < 0 ? 0 : (z > kmax ? kmax : z);
                    const uint64_t kc = morton((uint32_t)x, (uint32_t)y, (uint32_t)z);
                    hit |= (kc >= klo && kc < khi) ? 1 : 0;
                }
                sel |= (unsigned)hit << r;
            }
        }
    }
    // per round: ballot gives the warp's selected lanes; round-major exclusive offsets inside the tile
    unsigned int wcount[SS_ITEMS], below[SS_ITEMS];
    __shared__ unsigned int rcnt[SS_ITEMS][SS_T / 32];
#pragma unroll
    for (int r = 0; r < SS_ITEMS; r++) {
        const unsigned b = __ballot_sync(0xffffffffu, (sel >> r) & 1u);
        wcount[r] = __popc(b);
        below[r] = __popc(b & ((1u << lane) - 1u));
        if (lane == 0) rcnt[r][w] = wcount[r];
    }
    __syncthreads();
    // offsets of (round r, warp w) in round-major order; the tile total
    unsigned int base[SS_ITEMS], total = 0;
#pragma unroll
    for (int r = 0; r < SS_ITEMS; r++) {
        unsigned int acc = total;
#pragma unroll
        for (int ww = 0; ww < SS_T / 32; ww++) {
            if (ww == w) base[r] = acc;
            acc += rcnt[r][ww];
        }
        total = acc;
    }
    constexpr unsigned long long AGG = 1ull << 62, INCL = 2ull << 62, MASK = (1ull << 62) - 1;
    if (w == 0) {
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(&status[0]) = INCL | total;
        } else {
            if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(&status[tile]) = AGG | total;
            for (int64_t j = tile - 1;; j -= 32) {
                const int64_t idx = j - lane;
                unsigned long long sw = INCL;
                if (idx >= 0) {
                    do { sw = *reinterpret_cast<volatile const unsigned long long*>(&status[idx]); } while ((sw >> 62) == 0ull);
                }
                const unsigned incl = __ballot_sync(0xffffffffu, (sw & INCL) != 0ull);
                const int stop = incl ? (__ffs(incl) - 1) : 31;
                unsigned long long v = (lane <= stop) ? (sw & MASK) : 0ull;
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                excl += v;
                if (incl) break;
            }
            if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(&status[tile]) = INCL | (excl + total);
        }
        if (lane == 0) {
            prefix_s = excl;
            if ((tile + 1) * SS_TILE >= n) n_sel_out[0] = excl + total;
        }
    }
    __syncthreads();
    const int64_t pre = (int64_t)prefix_s;
#pragma unroll
    for (int r = 0; r < SS_ITEMS; r++) {
        if ((sel >> r) & 1u) {
            const int64_t i = tile * SS_TILE + (int64_t)r * SS_T + t;
            const int64_t d = pre + base[r] + below[r];
            const float4* src = reinterpret_cast<const float4*>(cloud + i * GPC_POINT_BYTES);
            float4* dst = reinterpret_cast<float4*>(sel_cloud + d * GPC_POINT_BYTES);
            dst[0] = src[0];
            dst[1] = src[1];
            sel_idx[d] = (int32_t)i;
        }
    }
}

// patches are in visiting order: out[0] = first patch inside the key range, out[1] = first patch behind it
__global__ void owned_range_kernel(const uint64_t* __restrict__ code, int64_t P, int leaf_order, const uint64_t* __restrict__ range,
                                   int64_t* __restrict__ out) {
    const uint64_t klo = range[0], khi = range[1];
    for (int e = 0; e < 2; e++) {
        int64_t lo = 0, hi = P;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            const uint64_t c = code[mid];
            // ascending: before the range while c < klo (e = 0) / c < khi (e = 1); descending: while c >= khi / c >= klo
            const bool before = leaf_order == 0 ? (e == 0 ? c >= khi : c >= klo) : (e == 0 ? c < klo : c < khi);
            if (before) lo = mid + 1; else hi = mid;
        }
        out[e] = lo;
    }
}

void launch_shard_sample(const uint64_t* keys, int64_t n, int64_t stride, int64_t m, uint64_t* sample, uint32_t* dummy, cudaStream_t s) {
    if (m <= 0) return;
    shard_sample_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(keys, n, stride, m, sample, dummy);
    g_launches++;
}
// weight / prefix: m + 1 int64 each; scan_tmp >= scan_tmp_bytes(m)
void launch_shard_splitters(const uint64_t* sorted, int64_t m, int64_t stride, int depth, int rank, int count, int leaf_order, int64_t* weight,
                            int64_t* prefix, void* scan_tmp, uint64_t* range2, cudaStream_t s) {
    shard_weights_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(sorted, m, stride, 1ull << (3 * depth), weight);
    launch_exclusive_scan_i64(weight, prefix, m, scan_tmp, s);
    shard_splitters_kernel<<<1, 1, 0, s>>>(sorted, prefix, m, 3 * depth, rank, count, leaf_order, range2);
    g_launches += 2;
}
void launch_shard_select(const uint64_t* keys, int64_t n, int depth, const uint64_t* range2, int64_t* flags, cudaStream_t s) {
    if (n <= 0) return;
    shard_select_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(keys, n, depth, range2, flags);
    g_launches++;
}
void launch_shard_compact(const uint8_t* cloud, const int64_t* ex, int64_t n, uint8_t* sel_cloud, int32_t* sel_idx, cudaStream_t s) {
    if (n <= 0) return;
    shard_compact_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cloud, ex, n, sel_cloud, sel_idx);
    g_launches++;
}
size_t shard_select_tmp_bytes(int64_t n) { return (size_t)((n + SS_TILE - 1) / SS_TILE + 4) * sizeof(unsigned long long); }
// n_sel_out: one device counter; tmp >= shard_select_tmp_bytes(n); sel_cloud / sel_idx sized for n points
void launch_shard_select_compact(const uint64_t* keys, int64_t n, int depth, const uint64_t* range2, const uint8_t* cloud, uint8_t* sel_cloud,
                                 int32_t* sel_idx, unsigned long long* n_sel_out, void* tmp, cudaStream_t s) {
    cudaMemsetAsync(n_sel_out, 0, sizeof(unsigned long long), s);
    if (n <= 0) return;
    cudaMemsetAsync(tmp, 0, shard_select_tmp_bytes(n), s);
    unsigned long long* status = reinterpret_cast<unsigned long long*>(tmp) + 2;
    unsigned int* counter = reinterpret_cast<unsigned int*>(tmp);
    shard_select_compact_kernel<<<(unsigned)((n + SS_TILE - 1) / SS_TILE), SS_T, 0, s>>>(keys, n, depth, range2, cloud, sel_cloud, sel_idx, n_sel_out,
                                                                                        status, counter);
    g_launches++;
}
void launch_owned_range(const uint64_t* code, int64_t P, int leaf_order, const uint64_t* range2, int64_t* out2, cudaStream_t s) {
    owned_range_kernel<<<1, 1, 0, s>>>(code, P, leaf_order, range2, out2);
    g_launches++;
}

void launch_point_keys(const uint8_t* cloud, int64_t n, const LatticeDev& lat, uint64_t* keys, uint32_t* vals, void* cpt, cudaStream_t s) {
    if (n <= 0) return;
    point_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cloud, n, lat, keys, vals, reinterpret_cast<float4*>(cpt));
    g_launches++;
}

size_t leaves_fused_tmp_bytes(int64_t n) { return (size_t)((n + LF_TILE - 1) / LF_TILE + 4) * sizeof(unsigned long long); }

// counts2: two device counters (n_valid, leaves), zeroed here; tmp >= leaves_fused_tmp_bytes(n)
void launch_leaves_fused(const uint64_t* keys, const uint32_t* vals, int64_t n, uint32_t depth, const void* cpt, int32_t* leaf_of,
                         int64_t* leaf_start, uint64_t* leaf_code, void* spt, unsigned long long* counts2, void* tmp, cudaStream_t s) {
    cudaMemsetAsync(counts2, 0, 2 * sizeof(unsigned long long), s);
    cudaMemsetAsync(leaf_start, 0, sizeof(int64_t), s);   // no valid point: leaf_start[0] = 0
    if (n <= 0) return;
    const int64_t tiles = (n + LF_TILE - 1) / LF_TILE;
    cudaMemsetAsync(tmp, 0, leaves_fused_tmp_bytes(n), s);
    unsigned long long* status = reinterpret_cast<unsigned long long*>(tmp) + 2;
    unsigned int* counter = reinterpret_cast<unsigned int*>(tmp);
    leaves_fused_kernel<<<(unsigned)tiles, LF_T, 0, s>>>(keys, vals, n, 1ull << (3 * depth), reinterpret_cast<const float4*>(cpt), leaf_of, leaf_start, leaf_code,
                                                         reinterpret_cast<float4*>(spt), counts2, status, counter);
    g_launches++;
}

void launch_leaf_neighbours(const uint64_t* leaf_code, int64_t P, const LatticeDev& lat, int32_t* nbr, int32_t* nnbr,
                            float* center, cudaStream_t s) {
    if (P <= 0) return;
    leaf_neighbours_kernel<<<(unsigned)((P * 32 + 255) / 256), 256, 0, s>>>(leaf_code, P, lat, nbr, nnbr, center);
    g_launches++;
}

void launch_leaf_rotation(const void* spt, const int64_t* leaf_start, const int32_t* nbr, const int32_t* nnbr,
                          const float* center, int64_t P, double r2, double* sums, double* Rm, int32_t* ncand, cudaStream_t s) {
    if (P <= 0) return;
    leaf_sums_kernel<<<(unsigned)((P * 32 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4*>(spt), leaf_start, nbr, nnbr, center,
                                                                    P, r2, sums);
    leaf_solve_kernel<<<(unsigned)((P + 127) / 128), 128, 0, s>>>(sums, center, P, Rm, ncand);
    g_launches += 2;
}

void launch_claim(const void* spt, const int32_t* leaf_of, const int32_t* nbr, const int32_t* nnbr, const float* center,
                  const double* Rm, const int32_t* ncand, int64_t n_valid, int64_t P, double r2, double half, int leaf_order,
                  uint64_t* okey, uint32_t* oval, double* pt0, double* pt1, double* pt2, cudaStream_t s) {
    if (n_valid <= 0) return;
    claim_kernel<<<(unsigned)((n_valid + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4*>(spt), leaf_of, nbr, nnbr, center,
                                                                  Rm, ncand, n_valid, P, r2, half, leaf_order, okey, oval, pt0, pt1,
                                                                  pt2);
    g_launches++;
}

void launch_patch_bounds(const uint64_t* okey, int64_t n_valid, int64_t P, int64_t* patch_off, cudaStream_t s) {
    patch_bounds_kernel<<<(unsigned)((P + 1 + 255) / 256), 256, 0, s>>>(okey, n_valid, P, patch_off);
    g_launches++;
}

void launch_group_gather(const uint64_t* okey, const uint32_t* oval, const uint32_t* sorted_idx, const void* spt,
                         const double* pt0, const double* pt1, const double* pt2, int64_t n_claimed, int32_t* st_idx, double* h,
                         double* x1, double* x2, uint32_t* rgb, int32_t* owner, cudaStream_t s) {
    if (n_claimed <= 0) return;
    group_gather_kernel<<<(unsigned)((n_claimed + 255) / 256), 256, 0, s>>>(okey, oval, sorted_idx, reinterpret_cast<const float4*>(spt),
                                                                          pt0, pt1, pt2, n_claimed, st_idx, h, x1, x2, rgb, owner);
    g_launches++;
}

void launch_patch_frames(const int64_t* patch_off, int64_t P, int leaf_order, const double* h, const uint32_t* rgb,
                         const uint64_t* leaf_code_a, const float* center_a, const double* Rm_a, const int32_t* ncand_a, double* y,
                         uint64_t* code, float* center, int32_t* ncand, double* Rm, double* quat, double* mean, double* rgbmean,
                         cudaStream_t s) {
    if (P <= 0) return;
    patch_frames_kernel<<<(unsigned)((P * 32 + 127) / 128), 128, 0, s>>>(patch_off, P, leaf_order, h, rgb, leaf_code_a, center_a, Rm_a,
                                                                       ncand_a, y, code, center, ncand, Rm, quat, mean, rgbmean);
    g_launches++;
}

}  // namespace gpc
