"""One small call of every kernel family, for `compute-sanitizer --tool memcheck|racecheck|synccheck`
(SURVEY.md section 5).  Sizes are small (the sanitizer slows kernels 10-100x) but cover every dispatch branch:
  FPS reg / pruned / cluster(handshake), ball query resident / stream / grid, gathers + grads, 3-NN, interpolation,
  fused kNN + selection sort, tcgen05 forward / dgrad / wgrad (wide, narrow alt-epilogue, K > 512), BN kernels,
  poolings, loss, dropout, Adam, box sampling, label vote, prob_sample, the Trainer with its side streams."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pn2_b200
from pn2_b200 import model
from pn2_b200._ffi import F32, F64, I32, call, lib, ptr
from pn2_b200.tf_ops import tf_grouping as tg, tf_interpolate as ti, tf_sampling as ts
from pn2_b200.util import pointnet_util as pu, tf_util
from pn2_b200.dataset.semantic_dataset import SemanticFileData

dev = "cuda"
rs = np.random.RandomState(0)
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
only = set(sys.argv[1:])
def section(name):
    ok = not only or name in only
    if ok:
        torch.cuda.synchronize(); print("==", name, flush=True)
    return ok

if section("fps"):
    for b, n, m in ((2, 100, 20), (2, 1024, 64), (2, 4096, 96), (1, 8192, 80), (1, 20000, 40)):
        ts.farthest_point_sample(m, cu(rs.random_sample((b, n, 3)).astype(np.float32)))
if section("ball"):
    x = cu(rs.random_sample((2, 1024, 3)).astype(np.float32)); q = x[:, :64].contiguous()
    tg.query_ball_point(0.2, 16, x, q)                                   # resident
    xb = cu(rs.random_sample((40, 4096, 3)).astype(np.float32))
    os.environ.get("PN2_BALL_GRID")
    tg.query_ball_point(0.1, 16, xb, xb[:, :4096].contiguous())           # grid (n >= 4096)
    xs = cu(rs.random_sample((64, 2048, 3)).astype(np.float32))
    tg.query_ball_point(0.1, 16, xs, xs.contiguous())                     # stream (many queries, n < 4096)
if section("gather"):
    x = cu(rs.random_sample((2, 512, 3)).astype(np.float32)).requires_grad_(True)
    idx = ts.farthest_point_sample(64, x)
    ts.gather_point(x, idx).sum().backward()
    p = cu(rs.random_sample((2, 512, 7)).astype(np.float32)).requires_grad_(True)
    gi, _ = tg.query_ball_point(0.3, 8, x.detach(), ts.gather_point(x.detach(), idx))
    tg.group_point(p, gi).sum().backward()
if section("knn"):
    x = cu(rs.randint(0, 4, (2, 300, 3)).astype(np.float32))
    tg.knn_point(16, x, x[:, :40].contiguous())
    tg.select_top_k(8, cu(rs.random_sample((2, 10, 200)).astype(np.float32)))
if section("interp"):
    x1 = cu(rs.random_sample((2, 600, 3)).astype(np.float32)); x2 = x1[:, :100].contiguous()
    d, i = ti.three_nn(x1, x2)
    p2 = cu(rs.random_sample((2, 100, 12)).astype(np.float32)).requires_grad_(True)
    ti.three_interpolate(p2, i, pu.fp_weights(d)).sum().backward()
    ti.interpolate_label_with_color(x2[0].contiguous(), cu(rs.randint(0, 9, 100).astype(np.int32)), x1[0].contiguous(), 5)
    ts.prob_sample(cu(rs.random_sample((1, 9000)).astype(np.float32)), cu(rs.random_sample((1, 500)).astype(np.float32)))
if section("gemm"):
    for (M, K, N) in ((1024, 131, 128), (2048, 32, 32), (1024, 6, 32), (512, 768, 256), (1024, 128, 9), (640, 67, 64)):
        lda = (K + 3) // 4 * 4
        A = torch.randn(M, lda, device=dev); W = torch.randn(K, N, device=dev) * .1; b = torch.randn(N, device=dev)
        sc = torch.rand(K, device=dev) + .5; sh = torch.rand(K, device=dev) - .5
        Y = torch.empty(M, N, device=dev); st = torch.zeros(2 * N, dtype=torch.float64, device=dev)
        ws = torch.empty(int(lib().pn2_linear_workspace_bytes(K, N)) // 4 + 4, device=dev)
        call("pn2_linear_fwd", M, K, N, ptr(A), lda, ptr(sc), ptr(sh), 1, ptr(W), ptr(b), ptr(Y), ptr(st), ptr(ws), ws.numel() * 4, 1)
        dX = torch.empty(M, K, device=dev)
        call("pn2_linear_dgrad", M, K, N, ptr(Y), ptr(W), ptr(dX), K, ptr(ws), ws.numel() * 4, 1)
        dW = torch.zeros(K, N, device=dev); db = torch.zeros(N, device=dev)
        call("pn2_linear_wgrad", M, K, N, ptr(A), lda, ptr(sc), ptr(sh), 1, ptr(Y), ptr(dW), ptr(db), -1)
if section("layers"):
    store = tf_util.set_default_store(tf_util.VariableStore(device=dev, seed=0))
    xyz = cu(rs.random_sample((2, 512, 3)).astype(np.float32)); pts = cu(rs.random_sample((2, 512, 3)).astype(np.float32)).requires_grad_(True)
    for pooling in ("max", "avg", "weighted_avg", "max_and_avg"):
        nx, f, _ = pu.pointnet_sa_module(xyz, pts, 64, 0.3, 16, [16, 32], None, False, True, 0.5, "sa_" + pooling, pooling=pooling)
        out = pu.pointnet_fp_module(xyz, nx, pts, f, [32, 16], True, 0.5, "fp_" + pooling)
        pred = tf_util.conv1d(tf_util.dropout(out, True, "dp"), 9, 1, scope="fc_" + pooling, activation_fn=None)
        model.get_loss(pred, cu(rs.randint(0, 9, (2, 512)).astype(np.int32)), cu(np.ones((2, 512), np.float32))).backward()
    # channel counts that are multiples of 4: the float4 paths of group_concat (forward and gradient), copy_cols and
    # three_interpolate_grad
    nx1, f1, _ = pu.pointnet_sa_module(xyz, pts, 64, 0.3, 16, [16, 32], None, False, True, 0.5, "v4_sa1")
    nx2, f2, _ = pu.pointnet_sa_module(nx1, f1, 16, 0.6, 8, [32, 64], None, False, True, 0.5, "v4_sa2")
    up = pu.pointnet_fp_module(nx1, nx2, f1, f2, [32, 32], True, 0.5, "v4_fp1")
    tf_util.dropout(up, True, "v4_dp").sum().backward()
    n = 1000
    p_, g_, m_, v_ = (torch.randn(n, device=dev) for _ in range(4)); v_.abs_()
    call("pn2_adam_step", n, ptr(p_), ptr(g_), ptr(m_), ptr(v_), 1e-3, 0.9, 0.999, 1e-8, 3, 1.0)
if section("feed"):
    pts = rs.random_sample((30000, 3)) * [30, 30, 5]
    fd = SemanticFileData(pts, rs.randint(0, 9, 30000), rs.random_sample((30000, 3)), 10.0, 10.0)
    fd.sample_batch(3, 2048, rng=np.random.RandomState(1), augment=True, seed=5)
    fd.sample_batch(2, 8192, rng=np.random.RandomState(2), seed=6)
if section("trainer"):
    # the training step with its side streams: weight gradients on a second stream, geometry one batch ahead inside the
    # captured graph, programmatic dependent launch on every kernel
    from pn2_b200.train_step import Trainer
    hp = {"use_color": 1, "batch_size": 2, "learning_rate": 0.001, "decay_step": 200000, "learning_rate_decay_rate": 0.7,
          "bn_init_decay": 0.5, "bn_decay_decay_rate": 0.5, "bn_decay_clip": 0.99,
          "l1_npoint": 128, "l1_radius": 0.2, "l1_nsample": 16, "l2_npoint": 64, "l2_radius": 0.4, "l2_nsample": 16,
          "l3_npoint": 16, "l3_radius": 0.8, "l3_nsample": 16, "l4_npoint": 8, "l4_radius": 1.2, "l4_nsample": 8}
    bt = [(cu(rs.random_sample((2, 512, 6)).astype(np.float32)), cu(rs.randint(0, 9, (2, 512)).astype(np.int32)),
           cu(np.ones((2, 512), np.float32))) for _ in range(2)]
    tr = Trainer(hp, 9, device=dev, seed=0, geometry_ahead=True, wgrad_sms=48)
    tr.step(*bt[0])
    tr.prime(*bt[0])
    assert tr.capture(*bt[0]), tr._capture_error
    tr.step_graph(*bt[1])
    tr.step_graph(*bt[0])
torch.cuda.synchronize()
print("SANITIZE_RUN_OK")
