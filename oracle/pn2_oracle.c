/*
 * pn2_oracle.c -- CPU restatement of the reference tf_ops kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / the
 * timed CPU baseline.
 *
 * Each function restates, loop for loop, the semantics of one reference kernel
 * (paths relative to the reference repository root):
 *
 *   orc_fps                    tf_ops/tf_sampling.cu:111-176
 *   orc_gather_point           tf_ops/tf_sampling.cu:178-191
 *   orc_gather_point_grad      tf_ops/tf_sampling.cu:193-206 (+memset tf_sampling.cpp:236)
 *   orc_query_ball_point       tf_ops/tf_grouping.cu:3-43
 *   orc_group_point            tf_ops/tf_grouping.cu:47-66
 *   orc_group_point_grad       tf_ops/tf_grouping.cu:70-90 (+memset tf_grouping.cpp:271)
 *   orc_three_nn               tf_ops/tf_interpolate.cpp:213-243 (Open3D KDTreeFlann
 *                              SearchKNN(q,3) == exact 3 smallest fp64 squared distances)
 *   orc_three_interpolate      tf_ops/tf_interpolate.cpp:307-330
 *   orc_three_interpolate_grad tf_ops/tf_interpolate.cpp:397-421 (+memset :477)
 *   orc_selection_sort         tf_ops/tf_grouping.cu:95-136
 *   orc_knn_point              tf_ops/tf_grouping.py:64-89 (+ tf_grouping.cu:95-136)
 *   orc_interpolate_label_with_color  tf_ops/tf_interpolate.cpp:71-115 (kNN label vote)
 *   orc_cumsum                 tf_ops/tf_sampling.cu:7-92 (prefix sum, reference rounding order)
 *   orc_prob_sample            tf_ops/tf_sampling.cu:7-110 (cumsum + binary search)
 *
 * Pinning status:
 *   - three_nn is pinned by the reference's only golden vector
 *     (tf_ops/test_interpolate.py:30-35), see tests/test_oracle_golden.py.
 *   - FPS / ball query / gather / group are NOT pinned by any reference test;
 *     they are pinned (a) by the kernel source restated here and (b) on the GPU
 *     box by running the reference's own unmodified .cu kernels compiled into
 *     oracle/_ref (tests/test_ref_kernels_gpu.py), bit for bit.
 *
 * Floating point: compile with -ffp-contract=off.  Where nvcc contracts the
 * reference expression into FMAs (verified in sm_100a SASS:
 * FMUL dy,dy ; FFMA dx,dx,. ; FFMA dz,dz,.) the contraction is written out with
 * fmaf() so that the result is bit-identical to what the reference kernel
 * computes on the device.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* squared distance exactly as the nvcc-compiled reference evaluates
 * (x2-x1)*(x2-x1)+(y2-y1)*(y2-y1)+(z2-z1)*(z2-z1)   (tf_sampling.cu:150-151,
 * tf_grouping.cu:27-29) */
static inline float sqdist_ref(float x1, float y1, float z1, float x2, float y2,
                               float z2) {
    float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* ------------------------------------------------------------------------- */
/* Farthest point sampling.  tf_sampling.cu:111-176.
 * The reference runs 512 threads; thread t scans k = t, t+512, ... keeping its
 * first strict maximum (best=-1, besti=0), then a 9-level pairwise tree keeps
 * the LEFT entry unless left < right.  That is simulated literally here with
 * 512 "lanes" so that the tie order is exactly the reference's:
 * lowest (k mod 512) first, then lowest k.
 * temp has n floats per cloud (the reference uses 32*n scratch, one row per
 * resident block).  threads=0 -> serial over batch, else OpenMP over batch. */
ORC_API void orc_fps(int b, int n, int m, const float *inp, int *out,
                     int threads) {
    if (m <= 0) return;
    enum { BS = 512 };
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic) num_threads(threads > 0 ? threads : 1)
#endif
    for (int i = 0; i < b; ++i) {
        const float *ds = inp + (size_t)i * n * 3;
        float *temp = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        float dists[BS];
        int dists_i[BS];
        int old = 0;
        out[(size_t)i * m + 0] = old;
        for (int j = 0; j < n; ++j) temp[j] = 1e38f;
        for (int j = 1; j < m; ++j) {
            float x1 = ds[old * 3 + 0], y1 = ds[old * 3 + 1], z1 = ds[old * 3 + 2];
            for (int t = 0; t < BS; ++t) {
                int besti = 0;
                float best = -1.f;
                for (int k = t; k < n; k += BS) {
                    float td = temp[k];
                    float d = sqdist_ref(x1, y1, z1, ds[k * 3 + 0], ds[k * 3 + 1],
                                         ds[k * 3 + 2]);
                    float d2 = fminf(d, td);
                    if (d2 != td) temp[k] = d2;
                    if (d2 > best) {
                        best = d2;
                        besti = k;
                    }
                }
                dists[t] = best;
                dists_i[t] = besti;
            }
            for (int u = 0; (1 << u) < BS; ++u) {
                for (int t = 0; t < (BS >> (u + 1)); ++t) {
                    int i1 = (t * 2) << u, i2 = (t * 2 + 1) << u;
                    if (dists[i1] < dists[i2]) {
                        dists[i1] = dists[i2];
                        dists_i[i1] = dists_i[i2];
                    }
                }
            }
            old = dists_i[0];
            out[(size_t)i * m + j] = old;
        }
        free(temp);
    }
}

/* tf_sampling.cu:178-191 */
ORC_API void orc_gather_point(int b, int n, int m, const float *inp,
                              const int *idx, float *out) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            int a = idx[(size_t)i * m + j];
            for (int c = 0; c < 3; ++c)
                out[((size_t)i * m + j) * 3 + c] = inp[((size_t)i * n + a) * 3 + c];
        }
}

/* tf_sampling.cu:193-206; output zeroed first as tf_sampling.cpp:236 does.
 * The device order of the atomic adds is unspecified; this one is j-ascending. */
ORC_API void orc_gather_point_grad(int b, int n, int m, const float *out_g,
                                   const int *idx, float *inp_g) {
    memset(inp_g, 0, sizeof(float) * (size_t)b * n * 3);
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            int a = idx[(size_t)i * m + j];
            for (int c = 0; c < 3; ++c)
                inp_g[((size_t)i * n + a) * 3 + c] += out_g[((size_t)i * m + j) * 3 + c];
        }
}

/* tf_grouping.cu:3-43.  Rows with no hit are left untouched by the reference
 * (uninitialised TF output); the oracle writes zeros there and reports cnt=0. */
ORC_API void orc_query_ball_point(int b, int n, int m, float radius, int nsample,
                                  const float *xyz1, const float *xyz2, int *idx,
                                  int *pts_cnt, int threads) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic) num_threads(threads > 0 ? threads : 1)
#endif
    for (int i = 0; i < b; ++i) {
        const float *p1 = xyz1 + (size_t)i * n * 3;
        const float *p2 = xyz2 + (size_t)i * m * 3;
        int *oi = idx + (size_t)i * m * nsample;
        int *oc = pts_cnt + (size_t)i * m;
        for (int j = 0; j < m; ++j) {
            int cnt = 0;
            for (int l = 0; l < nsample; ++l) oi[j * nsample + l] = 0;
            float x2 = p2[j * 3 + 0], y2 = p2[j * 3 + 1], z2 = p2[j * 3 + 2];
            for (int k = 0; k < n; ++k) {
                if (cnt == nsample) break;
                float x1 = p1[k * 3 + 0], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
                float d = fmaxf(sqrtf(sqdist_ref(x1, y1, z1, x2, y2, z2)), 1e-20f);
                if (d < radius) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) oi[j * nsample + l] = k;
                    oi[j * nsample + cnt] = k;
                    cnt += 1;
                }
            }
            oc[j] = cnt;
        }
    }
}

/* tf_grouping.cu:47-66 */
ORC_API void orc_group_point(int b, int n, int c, int m, int nsample,
                             const float *points, const int *idx, float *out) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j)
            for (int k = 0; k < nsample; ++k) {
                int ii = idx[((size_t)i * m + j) * nsample + k];
                memcpy(out + (((size_t)i * m + j) * nsample + k) * c,
                       points + ((size_t)i * n + ii) * c, sizeof(float) * c);
            }
}

/* tf_grouping.cu:70-90; zeroed first as tf_grouping.cpp:271 does. */
ORC_API void orc_group_point_grad(int b, int n, int c, int m, int nsample,
                                  const float *grad_out, const int *idx,
                                  float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * n * c);
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j)
            for (int k = 0; k < nsample; ++k) {
                int ii = idx[((size_t)i * m + j) * nsample + k];
                const float *g = grad_out + (((size_t)i * m + j) * nsample + k) * c;
                float *d = grad_points + ((size_t)i * n + ii) * c;
                for (int l = 0; l < c; ++l) d[l] += g[l];
            }
}

/* tf_interpolate.cpp:213-243.  Open3D's KDTreeFlann::SearchKNN(q, 3) is an exact
 * search (eps 0) on double-precision copies of the fp32 coordinates
 * (buffer_to_eigen_vector, tf_interpolate.cpp:20-28) and returns squared L2
 * distances in ascending order; the result is cast to float/int.  Restated as a
 * brute-force scan: d = (dx*dx + dy*dy) + dz*dz in fp64 without contraction,
 * the 3 smallest kept ascending, equal distances resolved to the lowest index.
 * (FLANN's tie order is unspecified; continuous data has no ties.) */
ORC_API void orc_three_nn(int b, int n, int m, const float *xyz1,
                          const float *xyz2, float *dists, int *indices,
                          int threads) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic) num_threads(threads > 0 ? threads : 1)
#endif
    for (int i = 0; i < b; ++i) {
        const float *q = xyz1 + (size_t)i * n * 3;
        const float *r = xyz2 + (size_t)i * m * 3;
        for (int j = 0; j < n; ++j) {
            double qx = q[j * 3 + 0], qy = q[j * 3 + 1], qz = q[j * 3 + 2];
            double b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; ++k) {
                double dx = qx - (double)r[k * 3 + 0];
                double dy = qy - (double)r[k * 3 + 1];
                double dz = qz - (double)r[k * 3 + 2];
                double d = (dx * dx + dy * dy) + dz * dz;
                if (d < b1) {
                    b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k;
                } else if (d < b2) {
                    b3 = b2; i3 = i2; b2 = d; i2 = k;
                } else if (d < b3) {
                    b3 = d; i3 = k;
                }
            }
            size_t o = ((size_t)i * n + j) * 3;
            dists[o + 0] = (float)b1; dists[o + 1] = (float)b2; dists[o + 2] = (float)b3;
            indices[o + 0] = i1; indices[o + 1] = i2; indices[o + 2] = i3;
        }
    }
}

/* tf_interpolate.cpp:307-330: fp32, p1*w1 + p2*w2 + p3*w3 left to right. */
ORC_API void orc_three_interpolate(int b, int m, int c, int n,
                                   const float *points, const int *idx,
                                   const float *weight, float *out) {
    for (int i = 0; i < b; ++i) {
        const float *p = points + (size_t)i * m * c;
        for (int j = 0; j < n; ++j) {
            size_t o = ((size_t)i * n + j) * 3;
            float w1 = weight[o], w2 = weight[o + 1], w3 = weight[o + 2];
            int i1 = idx[o], i2 = idx[o + 1], i3 = idx[o + 2];
            float *dst = out + ((size_t)i * n + j) * c;
            for (int l = 0; l < c; ++l)
                dst[l] = p[(size_t)i1 * c + l] * w1 + p[(size_t)i2 * c + l] * w2 +
                         p[(size_t)i3 * c + l] * w3;
        }
    }
}

/* tf_interpolate.cpp:397-421; zeroed first as :477 does. */
ORC_API void orc_three_interpolate_grad(int b, int n, int c, int m,
                                        const float *grad_out, const int *idx,
                                        const float *weight, float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * m * c);
    for (int i = 0; i < b; ++i) {
        float *gp = grad_points + (size_t)i * m * c;
        for (int j = 0; j < n; ++j) {
            size_t o = ((size_t)i * n + j) * 3;
            float w1 = weight[o], w2 = weight[o + 1], w3 = weight[o + 2];
            int i1 = idx[o], i2 = idx[o + 1], i3 = idx[o + 2];
            const float *g = grad_out + ((size_t)i * n + j) * c;
            for (int l = 0; l < c; ++l) {
                gp[(size_t)i1 * c + l] += g[l] * w1;
                gp[(size_t)i2 * c + l] += g[l] * w2;
                gp[(size_t)i3 * c + l] += g[l] * w3;
            }
        }
    }
}

/* tf_grouping.cu:95-136: copy dist, then partial selection sort of the first k
 * entries of every (b,m) row; strict '<' keeps the earliest minimum. */
ORC_API void orc_selection_sort(int b, int n, int m, int k, const float *dist,
                                int *outi, float *out) {
    for (size_t r = 0; r < (size_t)b * m; ++r) {
        float *p = out + r * n;
        int *pi = outi + r * n;
        for (int s = 0; s < n; ++s) {
            p[s] = dist[r * n + s];
            pi[s] = s;
        }
        for (int s = 0; s < k && s < n; ++s) {
            int mn = s;
            for (int t = s + 1; t < n; ++t)
                if (p[t] < p[mn]) mn = t;
            if (mn != s) {
                float tf = p[mn]; p[mn] = p[s]; p[s] = tf;
                int ti = pi[mn]; pi[mn] = pi[s]; pi[s] = ti;
            }
        }
    }
}

/* tf_grouping.py:64-89 knn_point: dist = reduce_sum((xyz1 - xyz2)**2, -1) in fp32 (left to right over
 * the c coordinates, separate multiply and add), select_top_k(k, dist) = the selection sort above, then the
 * first k columns.  The (b,m,n) matrix is built one row at a time.  PARITY UNPINNED for the summation
 * order of the three squares (TensorFlow absent; test_tf_ops.py:9-36 asserts nothing). */
ORC_API void orc_knn_point(int b, int n, int c, int m, int k, const float *xyz1, const float *xyz2,
                           float *val, int *idx) {
    float *p = (float *)malloc(sizeof(float) * (size_t)n);
    int *pi = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            const float *q = xyz2 + ((size_t)i * m + j) * c;
            for (int t = 0; t < n; ++t) {
                const float *x = xyz1 + ((size_t)i * n + t) * c;
                float acc = 0.f;
                for (int a = 0; a < c; ++a) {
                    const float df = x[a] - q[a];
                    const float sq = df * df;
                    acc = a == 0 ? sq : acc + sq;
                }
                p[t] = acc;
                pi[t] = t;
            }
            for (int s = 0; s < k && s < n; ++s) {
                int mn = s;
                for (int t = s + 1; t < n; ++t)
                    if (p[t] < p[mn]) mn = t;
                if (mn != s) {
                    float tf = p[mn]; p[mn] = p[s]; p[s] = tf;
                    int ti = pi[mn]; pi[mn] = pi[s]; pi[s] = ti;
                }
                val[((size_t)i * m + j) * k + s] = p[s];
                idx[((size_t)i * m + j) * k + s] = pi[s];
            }
        }
    free(p);
    free(pi);
}

/* ------------------------------------------------------------------------- */
/* Inclusive prefix sum with the reference's rounding sequence.
 * tf_sampling.cu:7-92 (cumsumKernel).  fp32 addition is not associative, so
 * the ORDER of the additions is part of the result.  The reference fixes it
 * as follows (restated here as a recurrence, not as its shared-memory scan):
 *   - a row is cut into chunks of 8192 elements (:9 BlockSize*4, :15);
 *   - inside a chunk, elements form quads.  A full quad (a,b,c,d) gets the
 *     local prefixes a, b+a, c+(b+a), (d+c)+(b+a) (:21-33); the ragged last
 *     quad is summed left to right starting from 0 (:35-44);
 *   - over the quad totals q[0..nq) the up-sweep (:47-58) builds balanced
 *     tree sums  S(p) = sum of the block of lowbit(p+1) quads ending at p,
 *     S over 2^(u+1) quads = S(right half) + S(left half);
 *   - the down-sweep (:59-70) turns them into inclusive prefixes
 *     P(p) = S(p) + P(p - lowbit(p+1))      (P(p) = S(p) when p+1 is a power of two);
 *   - element value = local prefix (+ P(quad-1) unless quad 0, :72-80), then
 *     + the carry of the previous chunks (:82-84);
 *   - the carry over chunk totals is a compensated (Kahan-style) sum (:85-89).
 */
#define ORC_SCAN_CHUNK 8192
static void cumsum_row_ref(int n, const float *in, float *out) {
    float loc[ORC_SCAN_CHUNK + 4];
    float S[ORC_SCAN_CHUNK / 4], P[ORC_SCAN_CHUNK / 4];
    float run = 0.0f, comp = 0.0f;
    for (int j = 0; j < n; j += ORC_SCAN_CHUNK) {
        const int len = n - j < ORC_SCAN_CHUNK ? n - j : ORC_SCAN_CHUNK;
        const int nq = (len + 3) >> 2;
        for (int q = 0; q < nq; ++q) {
            const float *v = in + j + 4 * q;
            float *e = loc + 4 * q;
            if (4 * q + 3 < len) {
                float ba = v[1] + v[0];
                float dc = v[3] + v[2];
                e[0] = v[0];
                e[1] = ba;
                e[2] = v[2] + ba;
                e[3] = dc + ba;
                S[q] = e[3];
            } else {
                float acc = 0.0f;
                for (int k = 4 * q; k < len; ++k) {
                    acc += in[j + k];
                    loc[k] = acc;
                }
                S[q] = acc;
            }
        }
        /* balanced tree sums, level by level (blocks of 2, 4, 8, ... quads) */
        for (int half = 1; 2 * half <= nq; half <<= 1)
            for (int p = 2 * half - 1; p < nq; p += 2 * half) S[p] = S[p] + S[p - half];
        /* inclusive prefixes in ascending order: P(p - lowbit) is final by then */
        for (int p = 0; p < nq; ++p) {
            int low = (p + 1) & -(p + 1);
            P[p] = (p + 1 == low) ? S[p] : S[p] + P[p - low];
        }
        for (int k = 0; k < len; ++k) {
            int q = k >> 2;
            float v = loc[k];
            if (q > 0) v = v + P[q - 1];
            out[j + k] = v + run;
        }
        {
            float t = P[nq - 1] + comp;
            float r2 = run + t;
            comp = t - (r2 - run);
            run = r2;
        }
    }
}

ORC_API void orc_cumsum(int b, int n, const float *inp, float *out) {
    for (int i = 0; i < b; ++i) cumsum_row_ref(n, inp + (size_t)i * n, out + (size_t)i * n);
}

/* tf_sampling.cu:94-110 (binarysearchKernel) over the prefix sum above:
 * q = r * cdf[n-1]; descending power-of-two steps, r -= k while cdf[r-k] >= q. */
ORC_API void orc_prob_sample(int b, int n, int m, const float *inp_p,
                             const float *inp_r, int *out) {
    float *cdf = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    int base = 1;
    while (base < n) base <<= 1;
    for (int i = 0; i < b; ++i) {
        cumsum_row_ref(n, inp_p + (size_t)i * n, cdf);
        for (int j = 0; j < m; ++j) {
            float q = inp_r[(size_t)i * m + j] * cdf[n - 1];
            int r = n - 1;
            for (int k = base; k >= 1; k >>= 1)
                if (r >= k && cdf[r - k] >= q) r -= k;
            out[(size_t)i * m + j] = r;
        }
    }
    free(cdf);
}

/* ------------------------------------------------------------------------- */
/* InterpolateLabelWithColor.  tf_interpolate.cpp:71-115: per dense point the
 * knn nearest sparse points (Open3D KDTreeFlann::SearchKNN on fp64 copies of
 * the fp32 coordinates == the knn smallest fp64 squared distances, ascending),
 * then the label that first reaches the highest count while walking the
 * neighbours from nearest to farthest (:97-106), then the colour table (:46-48).
 * Brute force; equal distances resolve to the lowest index (FLANN's tie order
 * is unspecified).  Labels outside the 9-entry table are undefined behaviour in
 * the reference (vector index out of range); here they get colour (0,0,0). */
static const unsigned char orc_label_colors[9][3] = {
    {255, 255, 255}, {0, 0, 255}, {128, 0, 0}, {255, 0, 255}, {0, 128, 0},
    {255, 0, 0},     {128, 0, 128}, {0, 0, 128}, {128, 128, 0}};

ORC_API void orc_interpolate_label_with_color(int num_sparse, int num_dense,
                                              const float *sparse_points,
                                              const int *sparse_labels,
                                              const float *dense_points,
                                              int *dense_labels,
                                              unsigned char *dense_colors, int knn,
                                              int threads) {
    if (knn < 0) knn = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
#endif
    for (int j = 0; j < num_dense; ++j) {
        double *bd = (double *)malloc(sizeof(double) * (size_t)(knn + 1));
        int *bi = (int *)malloc(sizeof(int) * (size_t)(knn + 1));
        int found = 0;
        double qx = dense_points[(size_t)j * 3 + 0], qy = dense_points[(size_t)j * 3 + 1],
               qz = dense_points[(size_t)j * 3 + 2];
        for (int s = 0; s < num_sparse; ++s) {
            double dx = qx - (double)sparse_points[(size_t)s * 3 + 0];
            double dy = qy - (double)sparse_points[(size_t)s * 3 + 1];
            double dz = qz - (double)sparse_points[(size_t)s * 3 + 2];
            double d = (dx * dx + dy * dy) + dz * dz;
            if (found < knn) {
                bd[found] = d;
                bi[found] = s;
                ++found;
            } else if (knn > 0 && d < bd[knn - 1]) {
                bd[knn - 1] = d;
                bi[knn - 1] = s;
            } else {
                continue;
            }
            for (int t = (found < knn ? found : knn) - 1; t > 0 && bd[t] < bd[t - 1]; --t) {
                double td = bd[t]; bd[t] = bd[t - 1]; bd[t - 1] = td;
                int ti = bi[t]; bi[t] = bi[t - 1]; bi[t - 1] = ti;
            }
        }
        int best = -1, best_count = 0;
        for (int t = 0; t < found; ++t) {
            int lab = sparse_labels[bi[t]], cnt = 0;
            for (int u = 0; u <= t; ++u) cnt += sparse_labels[bi[u]] == lab;
            if (cnt > best_count) {
                best = lab;
                best_count = cnt;
            }
        }
        dense_labels[j] = best;
        for (int c = 0; c < 3; ++c)
            dense_colors[(size_t)j * 3 + c] = (best >= 0 && best < 9) ? orc_label_colors[best][c] : 0;
        free(bd);
        free(bi);
    }
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
