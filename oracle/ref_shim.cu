/*
 * ref_shim.cu -- extern "C" doorway onto the reference's OWN launchers.
 *
 * TEST INFRASTRUCTURE ONLY.  The Makefile compiles the reference's unmodified
 * tf_ops/tf_sampling.cu and tf_ops/tf_grouping.cu where they lie under
 * /root/reference (never copied into this repo) and links them with this shim
 * into oracle/_ref/libref_tfops.so.  The shim only forwards to the launchers the
 * reference defines (tf_sampling.cu:208-229, tf_grouping.cu:138-162, incl. selectionSortLauncher :145-149) and adds the
 * memsets the reference's TF glue performs (tf_sampling.cpp:236,
 * tf_grouping.cpp:271) plus a device synchronise + error code, because the
 * reference launchers use the legacy default stream and never check errors.
 * All pointers are DEVICE pointers.
 */
#include <cuda_runtime.h>

void cumsumLauncher(int b, int n, const float* inp, float* out);
void probsampleLauncher(int b, int n, int m, const float* inp_p, const float* inp_r, float* temp,
                        int* out);
void farthestpointsamplingLauncher(int b, int n, int m, const float* inp, float* temp, int* out);
void gatherpointLauncher(int b, int n, int m, const float* inp, const int* idx, float* out);
void scatteraddpointLauncher(int b, int n, int m, const float* out_g, const int* idx, float* inp_g);
void queryBallPointLauncher(int b, int n, int m, float radius, int nsample, const float* xyz1,
                            const float* xyz2, int* idx, int* pts_cnt);
void selectionSortLauncher(int b, int n, int m, int k, const float* dist, int* outi, float* out);
void groupPointLauncher(int b, int n, int c, int m, int nsample, const float* points,
                        const int* idx, float* out);
void groupPointGradLauncher(int b, int n, int c, int m, int nsample, const float* grad_out,
                            const int* idx, float* grad_points);

static int finish(int sync) {
    cudaError_t e = cudaPeekAtLastError();
    if (e == cudaSuccess && sync) e = cudaDeviceSynchronize();
    return e == cudaSuccess ? 0 : -(int)e;
}

extern "C" {
int ref_cumsum(int b, int n, const float* inp, float* out, int sync) {
    cumsumLauncher(b, n, inp, out);
    return finish(sync);
}
/* temp: b*n floats of device scratch (tf_sampling.cpp:104-108) */
int ref_prob_sample(int b, int n, int m, const float* inp_p, const float* inp_r, float* temp,
                    int* out, int sync) {
    probsampleLauncher(b, n, m, inp_p, inp_r, temp, out);
    return finish(sync);
}
/* temp: 32*n floats of device scratch (tf_sampling.cpp:143-146) */
int ref_fps(int b, int n, int m, const float* inp, float* temp, int* out, int sync) {
    farthestpointsamplingLauncher(b, n, m, inp, temp, out);
    return finish(sync);
}
int ref_gather_point(int b, int n, int m, const float* inp, const int* idx, float* out, int sync) {
    gatherpointLauncher(b, n, m, inp, idx, out);
    return finish(sync);
}
int ref_gather_point_grad(int b, int n, int m, const float* out_g, const int* idx, float* inp_g,
                          int sync) {
    cudaMemset(inp_g, 0, sizeof(float) * (size_t)b * n * 3);
    scatteraddpointLauncher(b, n, m, out_g, idx, inp_g);
    return finish(sync);
}
/* idx is zero-filled first so that rows without any hit are defined (the reference leaves them
 * uninitialised, tf_grouping.cpp:108-114) */
int ref_query_ball_point(int b, int n, int m, float radius, int nsample, const float* xyz1,
                         const float* xyz2, int* idx, int* pts_cnt, int sync) {
    cudaMemset(idx, 0, sizeof(int) * (size_t)b * m * nsample);
    queryBallPointLauncher(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt);
    return finish(sync);
}
/* tf_grouping.cu:145-149: partial selection sort of every (b,m) row of the distance matrix */
int ref_selection_sort(int b, int n, int m, int k, const float* dist, int* outi, float* out, int sync) {
    selectionSortLauncher(b, n, m, k, dist, outi, out);
    return finish(sync);
}
int ref_group_point(int b, int n, int c, int m, int nsample, const float* points, const int* idx,
                    float* out, int sync) {
    groupPointLauncher(b, n, c, m, nsample, points, idx, out);
    return finish(sync);
}
int ref_group_point_grad(int b, int n, int c, int m, int nsample, const float* grad_out,
                         const int* idx, float* grad_points, int sync) {
    cudaMemset(grad_points, 0, sizeof(float) * (size_t)b * n * c);
    groupPointGradLauncher(b, n, c, m, nsample, grad_out, idx, grad_points);
    return finish(sync);
}
}
