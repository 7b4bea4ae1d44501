// pn2_layers.cu -- whole PointNet++ layers behind ONE C-ABI call each:
//
//   pn2_sa_forward / pn2_sa_backward   pointnet_sa_module, ball-query grouping + max pooling
//                                      (util/pointnet_util.py:98-216, the configuration model.py uses)
//   pn2_fp_forward / pn2_fp_backward   pointnet_fp_module (util/pointnet_util.py:285-326)
//
// No reference analogue: the reference strings ~10 TF ops per layer together in Python.  Here the host
// side of a layer is native: the entry point walks the op sequence (FPS -> gather -> ball query -> fused
// group+centre+concat -> [GEMM with the previous layer's BN+ReLU in its prologue and the batch statistics
// in its epilogue -> BN finalize] x L -> fused BN+ReLU+max-pool) and its exact reverse, launching the same
// kernels the Python layers use, on the caller's stream, with NO allocation: every intermediate lives in the
// caller-owned workspace whose layout `plan()` fixes (the forward leaves what the backward needs there).
// Variables and gradient accumulators are caller-owned device arrays handed over in pn2_conv_layer records.
#include <string.h>

#include "pn2_common.cuh"

PN2_API int pn2_fill_f32(long n, float value, float *dst, pn2_stream_t s);

namespace pn2 {

static inline size_t al(size_t bytes) { return (bytes + 255) / 256 * 256; }
static inline int pad4(int w) { return (w + 3) / 4 * 4; }

struct ChainPlan {
    size_t Y[PN2_MAX_LAYERS], stats[PN2_MAX_LAYERS], red[PN2_MAX_LAYERS], scale[PN2_MAX_LAYERS],
        shift[PN2_MAX_LAYERS], saved[PN2_MAX_LAYERS];
    size_t image, tmp0, tmp1, end;
};

// activations, BN scratch and backward temporaries of an L-layer chain over M rows, starting at `off`
static bool plan_chain(long M, int K0, int nlayers, const pn2_conv_layer *layers, size_t off, ChainPlan *p) {
    if (nlayers < 1 || nlayers > PN2_MAX_LAYERS) return false;
    int k = K0, maxw = K0;
    size_t img = 0;
    for (int i = 0; i < nlayers; ++i) {
        const pn2_conv_layer &L = layers[i];
        if (L.K != k || L.N < 1) return false;
        p->Y[i] = off;      off += al((size_t)M * L.N * 4);
        p->stats[i] = off;  off += al((size_t)2 * L.N * 8);
        p->red[i] = off;    off += al((size_t)2 * L.N * 8);
        p->scale[i] = off;  off += al((size_t)L.N * 4);
        p->shift[i] = off;  off += al((size_t)L.N * 4);
        p->saved[i] = off;  off += al((size_t)2 * L.N * 4);
        const size_t w = (size_t)pn2_linear_workspace_bytes(L.K, L.N);
        img = w > img ? w : img;
        k = L.N;
        maxw = L.N > maxw ? L.N : maxw;
    }
    p->image = off;  off += al(img + 16);
    p->tmp0 = off;   off += al((size_t)M * maxw * 4);   // backward: dY of the current layer
    p->tmp1 = off;   off += al((size_t)M * maxw * 4);   // backward: upstream gradient (dX)
    p->end = off;
    return true;
}

#define PN2_TRY(expr)            \
    do {                         \
        int rc__ = (expr);       \
        if (rc__) return rc__;   \
    } while (0)

template <typename T>
static inline T *at(void *ws, size_t off) { return reinterpret_cast<T *>(static_cast<char *>(ws) + off); }

// forward of the chain: x (M, K0) with row pitch ldx -> pre-activations Y_i, scale/shift/saved per layer
static int chain_forward(long M, int nlayers, const pn2_conv_layer *layers, const float *x, int ldx,
                         int is_training, float bn_eps, float bn_decay, void *ws, const ChainPlan &p,
                         pn2_stream_t s) {
    const float *a = x, *a_sc = nullptr, *a_sh = nullptr;
    int lda = ldx, a_relu = 0;
    cudaStream_t st = as_stream(s);
    for (int i = 0; i < nlayers; ++i) {
        const pn2_conv_layer &L = layers[i];
        float *Y = at<float>(ws, p.Y[i]);
        double *stats = nullptr;
        if (L.bn && is_training) {
            stats = at<double>(ws, p.stats[i]);
            PN2_TRY(cuda_status(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * L.N, st)));
        }
        const long img = pn2_linear_workspace_bytes(L.K, L.N);
        PN2_TRY(pn2_linear_fwd(M, L.K, L.N, a, lda, a_sc, a_sh, a_relu, L.W, L.bias, Y, stats,
                               at<float>(ws, p.image), img, -1, s));
        float *sc = at<float>(ws, p.scale[i]), *sh = at<float>(ws, p.shift[i]);
        if (L.bn) {
            if (is_training)
                PN2_TRY(pn2_bn_train_finalize(L.N, M, stats, L.gamma, L.beta, bn_eps, bn_decay, L.rank4 ? 1 : 0,
                                              L.moving_mean, L.moving_var, sc, sh, at<float>(ws, p.saved[i]), s));
            else
                PN2_TRY(pn2_bn_eval_affine(L.N, L.gamma, L.beta, L.moving_mean, L.moving_var, bn_eps, sc, sh, s));
            a_sc = sc;
            a_sh = sh;
        } else {
            a_sc = a_sh = nullptr;  // identity affine: the prologue applies only the ReLU
            if (L.relu) {           // the GEMM prologue wants scale/shift whenever it applies ReLU
                PN2_TRY(pn2_fill_f32(L.N, 1.0f, sc, s));
                PN2_TRY(pn2_fill_f32(L.N, 0.0f, sh, s));
                a_sc = sc;
                a_sh = sh;
            }
        }
        a = Y;
        lda = L.N;
        a_relu = L.relu ? 1 : 0;
    }
    return PN2_OK;
}

// backward of the chain.  `pooled`: the upstream gradient is dOut (G, N_last) routed through arg (max-pool);
// otherwise dense (M, N_last).  dX0 (M, K0) with pitch ldx0 may be NULL (no input gradient needed).
static int chain_backward(long M, int nlayers, const pn2_conv_layer *layers, const float *x, int ldx,
                          const float *d_out, const int *arg, long G, int ns, float *dX0, int ldx0, void *ws,
                          const ChainPlan &p, pn2_stream_t s) {
    cudaStream_t st = as_stream(s);
    const float *up = d_out;
    // two scratch matrices: the BN/ReLU backward writes dY into bufA, the dgrad writes dX into bufB; the
    // next (lower) layer reads its upstream gradient from bufB and reuses bufA
    float *bufA = at<float>(ws, p.tmp0), *bufB = at<float>(ws, p.tmp1);
    for (int i = nlayers - 1; i >= 0; --i) {
        const pn2_conv_layer &L = layers[i];
        const float *Y = at<float>(ws, p.Y[i]);
        const float *sc = at<float>(ws, p.scale[i]), *sh = at<float>(ws, p.shift[i]);
        const float *saved = at<float>(ws, p.saved[i]);
        double *red = at<double>(ws, p.red[i]);
        const bool pooled = arg != nullptr && i == nlayers - 1;
        const float *dY = up;
        if (L.bn) PN2_TRY(cuda_status(cudaMemsetAsync(red, 0, sizeof(double) * 2 * L.N, st)));
        if (pooled) {
            const bool aff = L.bn || L.relu;
            if (L.bn) PN2_TRY(pn2_bn_bwd_reduce_pool(G, ns, L.N, up, arg, Y, sc, sh, saved, L.relu, red, s));
            PN2_TRY(pn2_bn_bwd_apply_pool(G, ns, L.N, up, arg, Y, aff ? sc : nullptr, aff ? sh : nullptr,
                                          L.bn ? saved : nullptr, L.bn ? L.gamma : nullptr, L.relu, L.bn,
                                          L.bn ? red : nullptr, bufA, L.bn ? L.dgamma : nullptr,
                                          L.bn ? L.dbeta : nullptr, s));
            dY = bufA;
        } else if (L.bn || L.relu) {
            if (L.bn) PN2_TRY(pn2_bn_bwd_reduce(M, L.N, up, L.N, Y, sc, sh, saved, L.relu, red, s));
            PN2_TRY(pn2_bn_bwd_apply(M, L.N, up, L.N, Y, sc, sh, L.bn ? saved : nullptr, L.bn ? L.gamma : nullptr,
                                     L.relu, L.bn, L.bn ? red : nullptr, bufA, L.bn ? L.dgamma : nullptr,
                                     L.bn ? L.dbeta : nullptr, s));
            dY = bufA;
        }
        const float *a, *a_sc = nullptr, *a_sh = nullptr;
        int lda, a_relu = 0;
        if (i == 0) {
            a = x;
            lda = ldx;
        } else {
            const pn2_conv_layer &P = layers[i - 1];
            a = at<float>(ws, p.Y[i - 1]);
            lda = P.N;
            if (P.bn || P.relu) {
                a_sc = at<float>(ws, p.scale[i - 1]);
                a_sh = at<float>(ws, p.shift[i - 1]);
            }
            a_relu = P.relu ? 1 : 0;
        }
        // a bias in front of a train-mode BatchNorm has an exactly zero gradient: not accumulated
        PN2_TRY(pn2_linear_wgrad(M, L.K, L.N, a, lda, a_sc, a_sh, a_relu, dY, L.dW, L.bn ? nullptr : L.dbias, -1, s));
        if (i > 0 || dX0) {
            float *dX = i == 0 ? dX0 : (dY == bufB ? bufA : bufB);
            const int ld = i == 0 ? ldx0 : L.K;
            const long img = pn2_linear_workspace_bytes(L.K, L.N);
            PN2_TRY(pn2_linear_dgrad(M, L.K, L.N, dY, L.W, dX, ld, at<float>(ws, p.image), img, -1, s));
            up = dX;
            if (dX == bufA) {  // keep the invariant "upstream in bufB, bufA free"
                bufA = bufB;
                bufB = dX;
            }
        }
    }
    return PN2_OK;
}

// ---- set abstraction -----------------------------------------------------------------------------------
struct SaPlan {
    size_t fps, cnt, x0, arg, dx0;
    ChainPlan chain;
    int ld0;
    long M;
};

static bool plan_sa(const pn2_sa_config *c, const pn2_conv_layer *layers, SaPlan *p) {
    if (c->b < 1 || c->n < 1 || c->c < 0 || c->npoint < 1 || c->nsample < 1 || !(c->radius > 0.f)) return false;
    p->M = (long)c->b * c->npoint * c->nsample;
    p->ld0 = pad4(3 + c->c);
    size_t off = 0;
    p->fps = off;  off += al((size_t)c->b * c->npoint * 4);
    p->cnt = off;  off += al((size_t)c->b * c->npoint * 4);
    p->x0 = off;   off += al((size_t)p->M * p->ld0 * 4);
    p->dx0 = off;  off += al((size_t)p->M * p->ld0 * 4);
    const int nl = c->nlayers;
    if (nl < 1 || nl > PN2_MAX_LAYERS) return false;
    p->arg = off;  off += al((size_t)c->b * c->npoint * layers[nl - 1].N * 4);
    return plan_chain(p->M, 3 + c->c, nl, layers, off, &p->chain);
}

}  // namespace pn2

using namespace pn2;

namespace pn2 {
__global__ void fill_f32_kernel(long n, float value, float *__restrict__ dst) {
    pdl_enter();
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) dst[i] = value;
}
}  // namespace pn2

PN2_API long pn2_sa_workspace_bytes(const pn2_sa_config *cfg, const pn2_conv_layer *layers) {
    SaPlan p;
    if (!cfg || !layers || !plan_sa(cfg, layers, &p)) return -1;
    return (long)p.chain.end;
}

PN2_API int pn2_sa_forward(const pn2_sa_config *cfg, const pn2_conv_layer *layers, const float *xyz,
                           const float *points, float *new_xyz, float *new_points, int *idx, void *workspace,
                           long workspace_bytes, pn2_stream_t s) {
    PN2_REQUIRE_PTR(cfg);
    PN2_REQUIRE_PTR(layers);
    SaPlan p;
    PN2_REQUIRE(plan_sa(cfg, layers, &p));
    PN2_REQUIRE_PTR(xyz);
    PN2_REQUIRE_PTR(new_xyz);
    PN2_REQUIRE_PTR(new_points);
    PN2_REQUIRE_PTR(idx);
    PN2_REQUIRE_PTR(workspace);
    PN2_REQUIRE(workspace_bytes >= (long)p.chain.end && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
    if (cfg->c > 0) PN2_REQUIRE_PTR(points);
    const int b = cfg->b, n = cfg->n, m = cfg->npoint, ns = cfg->nsample, c = cfg->c;
    int *fps = at<int>(workspace, p.fps);
    // pointnet_util.py:36-54
    PN2_TRY(pn2_fps(b, n, m, xyz, nullptr, fps, s));
    PN2_TRY(pn2_gather_point(b, n, m, xyz, fps, new_xyz, s));
    PN2_TRY(pn2_query_ball_point(b, n, m, cfg->radius, ns, xyz, new_xyz, idx, at<int>(workspace, p.cnt), s));
    float *x0 = at<float>(workspace, p.x0);
    PN2_TRY(pn2_group_concat_ld(b, n, m, ns, c, xyz, new_xyz, points, idx, 1, 1, x0, p.ld0, s));
    // :150-170
    PN2_TRY(chain_forward(p.M, cfg->nlayers, layers, x0, p.ld0, cfg->is_training, cfg->bn_eps, cfg->bn_decay,
                          workspace, p.chain, s));
    const pn2_conv_layer &L = layers[cfg->nlayers - 1];
    const bool aff = L.bn || L.relu;
    return pn2_affine_act_maxpool((long)b * m, ns, L.N, at<float>(workspace, p.chain.Y[cfg->nlayers - 1]),
                                  aff ? at<float>(workspace, p.chain.scale[cfg->nlayers - 1]) : nullptr,
                                  aff ? at<float>(workspace, p.chain.shift[cfg->nlayers - 1]) : nullptr,
                                  L.relu, new_points, at<int>(workspace, p.arg), s);
}

PN2_API int pn2_sa_backward(const pn2_sa_config *cfg, const pn2_conv_layer *layers, const float *d_new_points,
                            const int *idx, float *d_points, void *workspace, long workspace_bytes,
                            pn2_stream_t s) {
    PN2_REQUIRE_PTR(cfg);
    PN2_REQUIRE_PTR(layers);
    SaPlan p;
    PN2_REQUIRE(plan_sa(cfg, layers, &p));
    PN2_REQUIRE(cfg->is_training);
    PN2_REQUIRE_PTR(d_new_points);
    PN2_REQUIRE_PTR(idx);
    PN2_REQUIRE_PTR(workspace);
    PN2_REQUIRE(workspace_bytes >= (long)p.chain.end);
    const int b = cfg->b, n = cfg->n, m = cfg->npoint, ns = cfg->nsample, c = cfg->c;
    const bool need_in = d_points != nullptr && c > 0;
    float *dx0 = need_in ? at<float>(workspace, p.dx0) : nullptr;
    PN2_TRY(chain_backward(p.M, cfg->nlayers, layers, at<float>(workspace, p.x0), p.ld0, d_new_points,
                           at<int>(workspace, p.arg), (long)b * m, ns, dx0, 3 + c, workspace, p.chain, s));
    if (need_in)  // scatter the feature part of the grouped gradient back (tf_grouping.py:57-61)
        PN2_TRY(pn2_group_concat_grad(b, n, m, ns, c, dx0, idx, 1, 1, d_points, nullptr, nullptr, s));
    return PN2_OK;
}

// ---- feature propagation ---------------------------------------------------------------------------------
namespace pn2 {
struct FpPlan {
    size_t dist, nn, w, x0, dx0;
    ChainPlan chain;
    int ld0;
    long M;
};
static bool plan_fp(const pn2_fp_config *c, const pn2_conv_layer *layers, FpPlan *p) {
    if (c->b < 1 || c->n1 < 1 || c->n2 < 3 || c->c1 < 0 || c->c2 < 1) return false;
    p->M = (long)c->b * c->n1;
    p->ld0 = pad4(c->c2 + c->c1);
    size_t off = 0;
    p->dist = off;  off += al((size_t)p->M * 3 * 4);
    p->nn = off;    off += al((size_t)p->M * 3 * 4);
    p->w = off;     off += al((size_t)p->M * 3 * 4);
    p->x0 = off;    off += al((size_t)p->M * p->ld0 * 4);
    p->dx0 = off;   off += al((size_t)p->M * p->ld0 * 4);
    return plan_chain(p->M, c->c2 + c->c1, c->nlayers, layers, off, &p->chain);
}
}  // namespace pn2

PN2_API long pn2_fp_workspace_bytes(const pn2_fp_config *cfg, const pn2_conv_layer *layers) {
    FpPlan p;
    if (!cfg || !layers || !plan_fp(cfg, layers, &p)) return -1;
    return (long)p.chain.end;
}

PN2_API int pn2_fp_forward(const pn2_fp_config *cfg, const pn2_conv_layer *layers, const float *xyz1,
                           const float *xyz2, const float *points1, const float *points2, float *out,
                           void *workspace, long workspace_bytes, pn2_stream_t s) {
    PN2_REQUIRE_PTR(cfg);
    PN2_REQUIRE_PTR(layers);
    FpPlan p;
    PN2_REQUIRE(plan_fp(cfg, layers, &p));
    PN2_REQUIRE_PTR(xyz1);
    PN2_REQUIRE_PTR(xyz2);
    PN2_REQUIRE_PTR(points2);
    PN2_REQUIRE_PTR(out);
    PN2_REQUIRE_PTR(workspace);
    PN2_REQUIRE(workspace_bytes >= (long)p.chain.end && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
    if (cfg->c1 > 0) PN2_REQUIRE_PTR(points1);
    const int b = cfg->b, n1 = cfg->n1, n2 = cfg->n2, c1 = cfg->c1, c2 = cfg->c2;
    float *dist = at<float>(workspace, p.dist), *w = at<float>(workspace, p.w), *x0 = at<float>(workspace, p.x0);
    int *nn = at<int>(workspace, p.nn);
    // pointnet_util.py:299-309
    PN2_TRY(pn2_three_nn(b, n1, n2, xyz1, xyz2, dist, nn, s));
    PN2_TRY(pn2_fp_weights((int)p.M, dist, w, s));
    PN2_TRY(pn2_three_interpolate_ld(b, n2, c2, n1, points2, nn, w, x0, p.ld0, s));
    if (c1 > 0) PN2_TRY(pn2_copy_cols(p.M, c1, points1, c1, x0 + c2, p.ld0, 0, s));
    // :313-325
    PN2_TRY(chain_forward(p.M, cfg->nlayers, layers, x0, p.ld0, cfg->is_training, cfg->bn_eps, cfg->bn_decay,
                          workspace, p.chain, s));
    const int last = cfg->nlayers - 1;
    const pn2_conv_layer &L = layers[last];
    const bool aff = L.bn || L.relu;
    return pn2_affine_act(p.M, L.N, at<float>(workspace, p.chain.Y[last]),
                          aff ? at<float>(workspace, p.chain.scale[last]) : nullptr,
                          aff ? at<float>(workspace, p.chain.shift[last]) : nullptr, L.relu, out, L.N, s);
}

PN2_API int pn2_fp_backward(const pn2_fp_config *cfg, const pn2_conv_layer *layers, const float *d_out,
                            float *d_points1, float *d_points2, void *workspace, long workspace_bytes,
                            pn2_stream_t s) {
    PN2_REQUIRE_PTR(cfg);
    PN2_REQUIRE_PTR(layers);
    FpPlan p;
    PN2_REQUIRE(plan_fp(cfg, layers, &p));
    PN2_REQUIRE(cfg->is_training);
    PN2_REQUIRE_PTR(d_out);
    PN2_REQUIRE_PTR(workspace);
    PN2_REQUIRE(workspace_bytes >= (long)p.chain.end);
    const int b = cfg->b, n1 = cfg->n1, n2 = cfg->n2, c1 = cfg->c1, c2 = cfg->c2;
    const bool need_in = d_points2 != nullptr || (d_points1 != nullptr && c1 > 0);
    float *dx0 = need_in ? at<float>(workspace, p.dx0) : nullptr;
    PN2_TRY(chain_backward(p.M, cfg->nlayers, layers, at<float>(workspace, p.x0), p.ld0, d_out, nullptr, 0, 0,
                           dx0, c2 + c1, workspace, p.chain, s));
    if (d_points2)  // tf_interpolate.py:62-71
        PN2_TRY(pn2_three_interpolate_grad_ld(b, n1, c2, n2, dx0, c2 + c1, at<int>(workspace, p.nn),
                                              at<float>(workspace, p.w), d_points2, s));
    if (d_points1 && c1 > 0) PN2_TRY(pn2_copy_cols(p.M, c1, dx0 + c2, c2 + c1, d_points1, c1, 0, s));
    return PN2_OK;
}

PN2_API int pn2_fill_f32(long n, float value, float *dst, pn2_stream_t s) {
    PN2_REQUIRE(n >= 0);
    if (n == 0) return PN2_OK;
    PN2_REQUIRE_PTR(dst);
    long blocks = (n + 255) / 256;
    if (blocks > 148L * 8) blocks = 148L * 8;
    launch_k(fill_f32_kernel, (int)blocks, 256, 0, as_stream(s), n, value, dst);
    return finish_launch();
}
