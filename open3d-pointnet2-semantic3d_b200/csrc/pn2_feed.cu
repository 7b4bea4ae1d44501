// pn2_feed.cu -- the step in front of the SA/FP path: z-column box sampling of a scene to fixed-size
// training samples, on the GPU.
//
// Reference (host numpy, one sample at a time): dataset/semantic_dataset.py
//   sample()                     :150-186  pick a centre point, crop the z-column, fix the size, centre
//   _extract_z_box()             :123-148  searchsorted over the x-sorted scene + 3-axis interval test
//   _get_fix_sized_sample_mask() : 90-107  random subset (order preserved) or tiling when short
//   _center_box()                :109-121  shift so that min z = 0 and the box is centred in x, y
//   sample_in_all_files()        :320-326  weights = label_weights[labels]
//   util/provider.py rotate_feature_point_cloud :72-102  random rotation about z (fp64), cast to fp32
// and train.py:225-244 feeds the result to the network every step (host -> device copy).
//
// Here the scene lives in HBM (fp64 coordinates / colours like Open3D's arrays, x-sorted like the
// reference keeps it) and ONE launch cuts B samples: CTA b owns sample b.  All arithmetic on the
// coordinates is fp64 like numpy's, the result is cast to fp32 once, so the output equals the numpy
// restatement (oracle/box_sample_ref.py) bit for bit without rotation and to the last fp32 bit or two
// with it (BLAS may fuse the 3-term dot product).
//
// Randomness: the centre index and the rotation angle of every sample are INPUTS (the caller's RNG
// stream, e.g. numpy's).  The fixed-size random subset cannot reuse numpy's Mersenne-Twister shuffle on
// the device; it is drawn with a counter-based generator instead: every in-box point i gets the key
// hash(seed, sample, i) and the num_point smallest keys are kept IN SCENE ORDER (a boolean-mask index
// like the reference's) -- a uniformly random subset, reproducible from (seed, sample).
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>

#include "pn2_common.cuh"

namespace pn2 {

constexpr int BS_THREADS = 1024;

__host__ __device__ __forceinline__ unsigned box_key(unsigned long long seed, int sample, long i) {
    unsigned long long z = seed + 0xD1B54A32D192ED03ull * (unsigned long long)(sample + 1) +
                           0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (unsigned)(z >> 32);
}

struct BoxParams {
    int B, num_point, feat, num_classes;
    long P;
    const double *points, *colors;
    const int *labels;
    const float *label_weights;
    const long *center_idx;
    const double *angles;
    double half_x, half_y, z_size;
    unsigned long long seed;
    float *out_data, *out_weights;
    int *out_labels, *out_index, *out_count;
};

__device__ __forceinline__ bool in_box(const double *__restrict__ p, const double *lo, const double *hi) {
    const double x = p[0], y = p[1], z = p[2];
    return x >= lo[0] && x <= hi[0] && y >= lo[1] && y <= hi[1] && z >= lo[2] && z <= hi[2];
}

__global__ void __launch_bounds__(BS_THREADS) box_sample_kernel(const BoxParams p) {
    pdl_enter();
    typedef cub::BlockScan<int, BS_THREADS> Scan;
    typedef cub::BlockReduce<int, BS_THREADS> ReduceI;
    typedef cub::BlockReduce<double, BS_THREADS> ReduceD;
    __shared__ union {
        typename Scan::TempStorage scan;
        typename ReduceI::TempStorage redi;
        typename ReduceD::TempStorage redd;
    } tmp;
    __shared__ double lo[3], hi[3], shift[3];
    __shared__ long range[2];
    __shared__ int hist[256];
    __shared__ int s_cnt, s_need, s_bin;
    __shared__ unsigned s_prefix;

    const int b = blockIdx.x, t = threadIdx.x;
    const int num = p.num_point;
    int *sel = p.out_index + (size_t)b * num;

    if (t == 0) {
        long ci = p.center_idx[b];
        ci = ci < 0 ? 0 : (ci >= p.P ? p.P - 1 : ci);  // np.random.randint(0, len(points)) never leaves [0, P)
        const double *c = p.points + 3 * ci;
        lo[0] = c[0] - p.half_x;  hi[0] = c[0] + p.half_x;   // semantic_dataset.py:134-143
        lo[1] = c[1] - p.half_y;  hi[1] = c[1] + p.half_y;
        lo[2] = c[2] - p.z_size;  hi[2] = c[2] + p.z_size;
        // np.searchsorted(points[:,0], v) (side='left'): first index with x >= v   (:145-146)
        for (int side = 0; side < 2; ++side) {
            const double v = side ? hi[0] : lo[0];
            long a = 0, e = p.P;
            while (a < e) {
                const long mid = (a + e) >> 1;
                if (p.points[3 * mid] < v) a = mid + 1;
                else e = mid;
            }
            range[side] = a;
        }
    }
    __syncthreads();
    const long i0 = range[0], i1 = range[1];

    // ---- pass A: how many points of the x-slice lie in the box (:147-155) ----------------------
    int mine = 0;
    for (long i = i0 + t; i < i1; i += BS_THREADS) mine += in_box(p.points + 3 * i, lo, hi) ? 1 : 0;
    const int cnt_r = ReduceI(tmp.redi).Sum(mine);
    if (t == 0) {
        s_cnt = cnt_r;
        p.out_count[b] = cnt_r;
    }
    __syncthreads();
    const int cnt = s_cnt;
    if (cnt == 0) {  // cannot happen when the centre is a scene point (the reference asserts, :157)
        for (int j = t; j < num; j += BS_THREADS) sel[j] = -1;
        return;
    }

    // ---- pass B: the num-th smallest key (radix select, 4 x 8 bits) when the box holds too many ---
    unsigned thr = 0xFFFFFFFFu;
    int take_eq = 0;  // how many of the points whose key == thr are kept (the first ones in scene order)
    const bool subsample = cnt > num;
    if (subsample) {
        if (t == 0) {
            s_need = num;
            s_prefix = 0u;
        }
        for (int pass = 0; pass < 4; ++pass) {
            const int shift_bits = 24 - 8 * pass;
            for (int k = t; k < 256; k += BS_THREADS) hist[k] = 0;
            __syncthreads();
            const unsigned prefix = s_prefix;
            const unsigned hi_mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift_bits + 8));
            for (long i = i0 + t; i < i1; i += BS_THREADS) {
                if (!in_box(p.points + 3 * i, lo, hi)) continue;
                const unsigned key = box_key(p.seed, b, i);
                if ((key & hi_mask) == prefix) atomicAdd(&hist[(key >> shift_bits) & 255u], 1);
            }
            __syncthreads();
            if (t == 0) {
                int need = s_need, bin = 0;
                while (bin < 255 && hist[bin] < need) {
                    need -= hist[bin];
                    ++bin;
                }
                s_need = need;
                s_bin = bin;
                s_prefix = prefix | ((unsigned)bin << shift_bits);
            }
            __syncthreads();
        }
        thr = s_prefix;
        take_eq = s_need;
    }

    // ---- pass C: ordered compaction of the kept points into sel[] -----------------------------
    int base_lt = 0, base_eq = 0;  // kept (key < thr) / seen (key == thr) so far, identical in all threads
    for (long c0 = i0; c0 < i1; c0 += BS_THREADS) {
        const long i = c0 + t;
        int lt = 0, eq = 0;
        if (i < i1 && in_box(p.points + 3 * i, lo, hi)) {
            if (!subsample) lt = 1;
            else {
                const unsigned key = box_key(p.seed, b, i);
                lt = key < thr ? 1 : 0;
                eq = key == thr ? 1 : 0;
            }
        }
        int lt_before, eq_before, lt_total, eq_total;
        Scan(tmp.scan).ExclusiveSum(lt, lt_before, lt_total);
        __syncthreads();
        Scan(tmp.scan).ExclusiveSum(eq, eq_before, eq_total);
        __syncthreads();
        const int eqb = base_eq + eq_before;
        const bool keep = lt || (eq && eqb < take_eq);
        if (keep) {
            const int pos = base_lt + lt_before + (eqb < take_eq ? eqb : take_eq);
            if (pos < num) sel[pos] = (int)i;  // P < 2^31 is checked on the host
        }
        base_lt += lt_total;
        base_eq += eq_total;
    }
    __syncthreads();
    const int kept = subsample ? num : cnt;
    // not enough points: arange(cnt) tiled to num_point (:100-106)
    for (int j = kept + t; j < num; j += BS_THREADS) sel[j] = sel[j % kept];
    __syncthreads();

    // ---- pass D: _center_box (:109-121): min over the SAMPLE's points -------------------------
    double mn[3] = {1e300, 1e300, 1e300};
    for (int j = t; j < kept; j += BS_THREADS) {
        const double *q = p.points + 3 * (long)sel[j];
        mn[0] = fmin(mn[0], q[0]);
        mn[1] = fmin(mn[1], q[1]);
        mn[2] = fmin(mn[2], q[2]);
    }
    for (int a = 0; a < 3; ++a) {
        const double r = ReduceD(tmp.redd).Reduce(mn[a], cub::Min());
        if (t == 0) shift[a] = a == 0 ? r + p.half_x : (a == 1 ? r + p.half_y : r);
        __syncthreads();
    }

    // ---- pass E: centre, rotate about z (provider.py:94-101), cast, gather colours/labels/weights ---
    double cs = 1.0, sn = 0.0;
    if (p.angles) {
        cs = cos(p.angles[b]);
        sn = sin(p.angles[b]);
    }
    const int w = 3 + p.feat;
    for (int j = t; j < num; j += BS_THREADS) {
        const long i = sel[j];
        const double *q = p.points + 3 * i;
        const double x = q[0] - shift[0], y = q[1] - shift[1], z = q[2] - shift[2];
        float *o = p.out_data + ((size_t)b * num + j) * w;
        if (p.angles) {
            // row vector times [[c, s, 0], [-s, c, 0], [0, 0, 1]]
            o[0] = (float)(x * cs + y * (-sn) + z * 0.0);
            o[1] = (float)(x * sn + y * cs + z * 0.0);
            o[2] = (float)(x * 0.0 + y * 0.0 + z * 1.0);
        } else {
            o[0] = (float)x;
            o[1] = (float)y;
            o[2] = (float)z;
        }
        for (int a = 0; a < p.feat; ++a) o[3 + a] = p.colors ? (float)p.colors[p.feat * i + a] : 0.f;
        const int lab = p.labels ? p.labels[i] : 0;
        if (p.out_labels) p.out_labels[(size_t)b * num + j] = lab;
        if (p.out_weights)
            p.out_weights[(size_t)b * num + j] =
                (p.label_weights && lab >= 0 && lab < p.num_classes) ? p.label_weights[lab] : 1.f;
    }
}

}  // namespace pn2

using namespace pn2;

PN2_API unsigned pn2_box_sample_key(unsigned long long seed, int sample, long i) { return box_key(seed, sample, i); }

PN2_API int pn2_box_sample(int B, long P, int num_point, int feat, const double *points, const double *colors,
                           const int *labels, const float *label_weights, int num_classes,
                           const long *center_idx, const double *angles, double box_size_x,
                           double box_size_y, double scene_z_size, unsigned long long seed, float *out_data,
                           int *out_labels, float *out_weights, int *out_index, int *out_count,
                           pn2_stream_t s) {
    PN2_REQUIRE(B >= 0 && P > 0 && P < (1L << 31) && num_point > 0 && feat >= 0 && feat <= 16);
    PN2_REQUIRE(box_size_x > 0.0 && box_size_y > 0.0 && scene_z_size >= 0.0);
    if (B == 0) return PN2_OK;
    PN2_REQUIRE_PTR(points);
    PN2_REQUIRE_PTR(center_idx);
    PN2_REQUIRE_PTR(out_data);
    PN2_REQUIRE_PTR(out_index);
    PN2_REQUIRE_PTR(out_count);
    if (feat > 0) PN2_REQUIRE_PTR(colors);
    BoxParams p;
    p.B = B;
    p.num_point = num_point;
    p.feat = feat;
    p.num_classes = num_classes;
    p.P = P;
    p.points = points;
    p.colors = colors;
    p.labels = labels;
    p.label_weights = label_weights;
    p.center_idx = center_idx;
    p.angles = angles;
    p.half_x = box_size_x / 2;
    p.half_y = box_size_y / 2;
    p.z_size = scene_z_size;
    p.seed = seed;
    p.out_data = out_data;
    p.out_weights = out_weights;
    p.out_labels = out_labels;
    p.out_index = out_index;
    p.out_count = out_count;
    launch_k(box_sample_kernel, B, BS_THREADS, 0, as_stream(s), p);
    return finish_launch();
}
