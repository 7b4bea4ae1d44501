#!/usr/bin/env python
"""Op-level microbenchmarks on ONE B200: BASELINE.json configs[4] (FPS + ball-query sweep,
N in {4096, 16384, 65536, 262144}, npoint = N/4, nsample = 64) plus the standalone index/gather
ops at their config-2 sizes, each against the roofline SURVEY.md 8(d) assigns to it.

    python profiles/op_sweep.py [--budget SECONDS] [--only cfg2|msg|fps_cluster|ball_grid] [--no-ref] [--out FILE]

Inputs are resident in HBM, every op is called through the public Python op surface (ctypes ->
C ABI -> sm_100a kernel), timed with CUDA events on the launching stream after one warm-up, a
256 MB L2 flush before every timed call.  Where the reference has a CUDA kernel of its own
(a1-a4) and oracle/_ref was built, that kernel -- compiled unmodified for sm_100a -- is timed on
the same GPU as the second comparator.  One JSON document goes to --out (default
gpurun_out/op_sweep.json); profiles/summarize_sweep.py turns it into the tracked table.

`--only cfg2` runs each config-2 op once without timing loops: the target of
`ncu --set full -k regex:...` captures.  `--only msg` times BASELINE.json configs[2]: one MSG
set-abstraction layer (B=16, N=8192, npoint=1024, radii 0.1/0.2/0.4, nsample 16/32/128, widths of
SURVEY.md 8d), forward + backward, in its own process.
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--budget", type=float, default=90.0, help="stop starting new cases after this many seconds")
    ap.add_argument("--only", default="", help="'cfg2': one untimed call per config-2 op (ncu target); "
                                               "'msg': the config-3 MSG layer, forward + backward")
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--fps-entry", default="pn2_fps_cluster",
                    help="entry point timed by --only fps_cluster (pn2_fps_cluster | pn2_fps_cluster_mb)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "op_sweep.json"))
    args = ap.parse_args()

    import torch
    import pn2_b200  # noqa: F401
    from pn2_b200.tf_ops import tf_grouping, tf_interpolate, tf_sampling

    assert torch.cuda.is_available(), "op_sweep needs a GPU"
    torch.cuda.set_device(0)
    t_start = time.time()
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm = peaks.get("hbm_gbs", 6650.0)
    ref = None
    if not args.no_ref:
        try:
            from _util import RefKernels
            ref = RefKernels()
        except Exception as e:  # oracle/_ref not built: second comparator unavailable
            print("reference kernels unavailable:", e, file=sys.stderr)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")

    def timed(fn, max_reps=5, min_total_ms=30.0):
        """Median of up to max_reps event-timed calls (1 warm-up); long calls are timed once."""
        fn()
        torch.cuda.synchronize()
        times = []
        for _ in range(max_reps):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            if times[-1] > 500.0 or (sum(times) > min_total_ms and len(times) >= 3):
                break
        return float(np.median(times))

    def cloud(seed, b, n, scale=(1.0, 1.0, 1.0), shift=(0.0, 0.0, 0.0)):
        rs = np.random.RandomState(seed)
        x = rs.random_sample((b, n, 3)).astype(np.float32)
        x = x * np.asarray(scale, np.float32) + np.asarray(shift, np.float32)
        return torch.as_tensor(x.astype(np.float32)).cuda()

    results = []

    def dump():
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump({"peaks": peaks, "hbm_gbs": hbm, "elapsed_s": time.time() - t_start,
                   "results": results}, open(args.out, "w"), indent=1)

    def record(**kw):
        results.append(kw)
        print(json.dumps(kw), flush=True)
        dump()  # after every case: a timeout must not lose what was measured

    def out_of_time():
        return time.time() - t_start > args.budget

    def fps_case(tag, b, n, m, x, with_ref):
        ms = timed(lambda: tf_sampling.farthest_point_sample(m, x))
        stream_bytes = b * (m - 1) * n * 20.0          # SURVEY 8d streaming model
        comp_bytes = b * (12.0 * n + 4.0 * m)          # compulsory
        rec = dict(op="farthest_point_sample", case=tag, B=b, N=n, npoint=m, ms=ms,
                   us_per_round=ms * 1e3 / max(m - 1, 1),
                   streaming_model_GBps=stream_bytes / ms / 1e6,
                   streaming_model_frac_of_hbm=stream_bytes / ms / 1e6 / hbm,
                   compulsory_GBps=comp_bytes / ms / 1e6, bound="serial latency (on-chip cloud)")
        if with_ref and ref is not None:
            rms = timed(lambda: ref.fps(x, m), max_reps=3)
            rec.update(ref_kernel_ms=rms, speedup_vs_ref_kernel=rms / ms)
        record(**rec)

    def ball_case(tag, b, n, m, radius, ns, x, q, with_ref):
        ms = timed(lambda: tf_grouping.query_ball_point(radius, ns, x, q))
        pairs = float(b) * m * n
        comp_bytes = b * (12.0 * n + 12.0 * m + 4.0 * m * ns + 4.0 * m)
        rec = dict(op="query_ball_point", case=tag, B=b, n=n, m=m, nsample=ns, radius=radius, ms=ms,
                   pair_tests_per_s=pairs / ms * 1e3, compulsory_GBps=comp_bytes / ms / 1e6,
                   compulsory_frac_of_hbm=comp_bytes / ms / 1e6 / hbm,
                   # 7 FP32 instructions per pair test (3 sub, mul, 2 fma, compare) on 148 SMs x 128 lanes
                   frac_of_fp32_issue=(pairs * 7 / ms * 1e3) / (148 * 128 * peaks.get("sm_max_mhz", 1965.0) * 1e6),
                   bound="FP32 issue (brute-force radius search)")
        if with_ref and ref is not None:
            rms = timed(lambda: ref.query_ball_point(radius, ns, x, q), max_reps=3)
            rec.update(ref_kernel_ms=rms, speedup_vs_ref_kernel=rms / ms)
        record(**rec)

    def nn_case(tag, b, n, m, q, known):
        ms = timed(lambda: tf_interpolate.three_nn(q, known))
        pairs = float(b) * n * m
        comp_bytes = b * (12.0 * n + 12.0 * m + 24.0 * n)
        record(op="three_nn", case=tag, B=b, n=n, m=m, ms=ms, pair_tests_per_s=pairs / ms * 1e3,
               fp64_ops_per_s=pairs * 8 / ms * 1e3, compulsory_GBps=comp_bytes / ms / 1e6,
               compulsory_frac_of_hbm=comp_bytes / ms / 1e6 / hbm, bound="FP64 ALU (bit-exact fp64 distances)")

    # ---------------------------------------------------------------- config 3: one MSG layer
    if args.only == "msg":
        from pn2_b200.util import pointnet_util, tf_util
        tf_util.set_default_store(tf_util.VariableStore(device="cuda", seed=0))
        B, N, npoint = 16, 8192, 1024
        radii, nss = [0.1, 0.2, 0.4], [16, 32, 128]
        mlps = [[32, 32, 64], [64, 64, 128], [64, 96, 128]]
        xyz = cloud(100, B, N)  # unit cube
        pts = torch.as_tensor(np.random.RandomState(7).random_sample((B, N, 6)).astype(np.float32)).cuda()
        pts.requires_grad_(True)
        g = torch.ones((B, npoint, sum(m[-1] for m in mlps)), dtype=torch.float32, device="cuda")
        state = {}

        def step():
            new_xyz, out = pointnet_util.pointnet_sa_module_msg(xyz, pts, npoint, radii, nss, mlps,
                                                                True, 0.9, "msg")
            out.backward(g)
            state["out"] = out

        ms = timed(step)
        out = state["out"].detach()
        rows = [B * npoint * k for k in nss]
        flops = sum(2.0 * r * sum(a * b for a, b in zip([9] + m[:-1], m)) for r, m in zip(rows, mlps))
        record(op="pointnet_sa_module_msg fwd+bwd", case="cfg3 B=16 N=8192 npoint=1024 C=6", ms=ms,
               points_per_s=B * N / ms * 1e3, out_shape=list(out.shape),
               finite=bool(torch.isfinite(out).all()), grouped_rows=rows,
               gemm_gflop_fwd=flops / 1e9, gemm_tflops_fwd_plus_bwd=3 * flops / ms / 1e9,
               note="FPS + 3x(ball query, group+concat, 3-layer shared MLP with train-mode BN, max pool) and "
                    "the backward pass; 3xTF32 tensor-core GEMMs")
        print("wrote", args.out)
        return

    # ------------------------------------------- config 5 FPS: cluster kernel vs one CTA per cloud
    if args.only == "fps_cluster":
        from pn2_b200._ffi import F32, I32, call, ptr

        def cluster_fps(x, m):
            out = torch.empty((x.shape[0], m), dtype=I32, device=x.device)
            call(args.fps_entry, x.shape[0], x.shape[1], m, ptr(x, F32), ptr(out, I32))
            return out

        for n in (16384, 65536, 262144):
            for b in (16, 1):
                m = n // 4
                x = cloud(100 + n, b, n)
                ms = timed(lambda: cluster_fps(x, m))
                same = None
                if not out_of_time():  # full-size parity: the single-CTA kernels are bit-exact vs the oracle
                    same = bool((cluster_fps(x, m) == tf_sampling.farthest_point_sample(m, x)).all())
                stream_bytes = b * (m - 1) * n * 20.0
                record(op="farthest_point_sample (%s)" % args.fps_entry, case="sweep N=%d B=%d" % (n, b), B=b, N=n,
                       npoint=m, ms=ms, us_per_round=ms * 1e3 / (m - 1),
                       streaming_model_GBps=stream_bytes / ms / 1e6,
                       streaming_model_frac_of_hbm=stream_bytes / ms / 1e6 / hbm,
                       identical_to_single_cta_kernel=same)
                del x
        print("wrote", args.out)
        return

    # ------------------------------------------- ball query: hashed grid (experimental) vs default kernels
    if args.only == "ball_grid":
        from pn2_b200._ffi import F32, I32, call, lib, ptr

        def grid_ball(radius, ns, x1, x2, ws, nbytes):
            idx = torch.empty((x1.shape[0], x2.shape[1], ns), dtype=I32, device=x1.device)
            cnt = torch.empty((x1.shape[0], x2.shape[1]), dtype=I32, device=x1.device)
            call("pn2_query_ball_point_grid", x1.shape[0], x1.shape[1], x2.shape[1], float(radius), ns,
                 ptr(x1, F32), ptr(x2, F32), ptr(idx, I32), ptr(cnt, I32), ptr(ws, F32), nbytes)
            return idx, cnt

        cases = [("cfg2 SA1", 16, 8192, 1024, 0.5, 32, (10, 10, 5), (-5, -5, 0))]
        for n in (4096, 16384, 65536, 262144):
            for b in (16, 1):
                cases.append(("sweep N=%d B=%d" % (n, b), b, n, n // 4,
                              float((3.0 * 2 * 64 / (4.0 * math.pi * n)) ** (1.0 / 3.0)), 64, (1, 1, 1), (0, 0, 0)))
        for tag, b, n, m, radius, ns, sc, sh in cases:
            if out_of_time():
                break
            x = cloud(100 + n, b, n, sc, sh)
            q = tf_sampling.gather_point(x, tf_sampling.farthest_point_sample(m, x))
            nbytes = int(lib().pn2_ball_grid_workspace_bytes(b, n))
            ws = torch.empty((nbytes + 15) // 16 * 4, dtype=torch.float32, device="cuda")
            ms_g = timed(lambda: grid_ball(radius, ns, x, q, ws, nbytes))
            ms_d = timed(lambda: tf_grouping.query_ball_point(radius, ns, x, q))
            gi, gc = grid_ball(radius, ns, x, q, ws, nbytes)
            di, dc = tf_grouping.query_ball_point(radius, ns, x, q)
            record(op="query_ball_point grid vs default", case=tag, B=b, n=n, m=m, nsample=ns, radius=radius,
                   grid_ms=ms_g, default_ms=ms_d, speedup=ms_d / ms_g,
                   identical=bool((gi == di).all()) and bool((gc == dc).all()))
            del x, q, ws
        print("wrote", args.out)
        return

    # ---------------------------------------------------------------- config 2 op sizes
    B, N = 16, 8192
    x = cloud(100, B, N, (10, 10, 5), (-5, -5, 0))
    feat = torch.as_tensor(np.random.RandomState(1).random_sample((B, 1024, 128)).astype(np.float32)).cuda()
    if args.only == "cfg2":
        idx = tf_sampling.farthest_point_sample(1024, x)
        new_xyz = tf_sampling.gather_point(x, idx)
        gidx, _ = tf_grouping.query_ball_point(0.5, 32, x, new_xyz)
        tf_grouping.group_point(x, gidx)
        dist, i3 = tf_interpolate.three_nn(x, new_xyz)
        w = torch.full_like(dist, 1.0 / 3)
        tf_interpolate.three_interpolate(feat, i3, w)
        torch.cuda.synchronize()
        print("cfg2 ops done")
        return

    fps_case("cfg2 SA1", B, N, 1024, x, True)
    idx = tf_sampling.farthest_point_sample(1024, x)
    new_xyz = tf_sampling.gather_point(x, idx)
    ball_case("cfg2 SA1", B, N, 1024, 0.5, 32, x, new_xyz, True)
    nn_case("cfg2 FP4", B, N, 1024, x, new_xyz)
    gidx, _ = tf_grouping.query_ball_point(0.5, 32, x, new_xyz)
    for c, src in ((3, x), (128, torch.as_tensor(
            np.random.RandomState(2).random_sample((B, N, 128)).astype(np.float32)).cuda())):
        ms = timed(lambda: tf_grouping.group_point(src, gidx))
        by = B * 1024 * 32 * (4.0 + 8.0 * c)
        rec = dict(op="group_point", case="cfg2 SA1 idx, C=%d" % c, B=B, n=N, m=1024, nsample=32, C=c,
                   ms=ms, GBps=by / ms / 1e6, frac_of_hbm=by / ms / 1e6 / hbm, bound="HBM gather")
        if ref is not None:
            rms = timed(lambda: ref.group_point(src, gidx), max_reps=3)
            rec.update(ref_kernel_ms=rms, speedup_vs_ref_kernel=rms / ms)
        record(**rec)
    dist, i3 = tf_interpolate.three_nn(x, new_xyz)
    w = torch.full_like(dist, 1.0 / 3)
    ms = timed(lambda: tf_interpolate.three_interpolate(feat, i3, w))
    by = B * N * (24.0 + 16.0 * 128)
    record(op="three_interpolate", case="cfg2 FP4 c=128", B=B, n=N, m=1024, C=128, ms=ms,
           GBps=by / ms / 1e6, frac_of_hbm=by / ms / 1e6 / hbm, bound="HBM / L2 gather")
    pw = torch.as_tensor(np.random.RandomState(3).random_sample((B, 4096)).astype(np.float32)).cuda()
    pr = torch.as_tensor(np.random.RandomState(4).random_sample((B, N)).astype(np.float32)).cuda()
    ms = timed(lambda: tf_sampling.prob_sample(pw, pr))
    record(op="prob_sample", case="B=16, 4096 categories, 8192 draws", ms=ms, bound="latency (tiny)")
    del feat, gidx, dist, i3, w

    # ---------------------------------------------------------------- config 5 sweep
    for n in (4096, 16384, 65536, 262144):
        for b in (16, 1):
            if out_of_time():
                record(op="sweep", case="N=%d B=%d" % (n, b), skipped="time budget of %.0f s used up" % args.budget)
                continue
            m, ns = n // 4, 64
            radius = float((3.0 * 2 * ns / (4.0 * math.pi * n)) ** (1.0 / 3.0))
            x = cloud(100 + n, b, n)
            tag = "sweep N=%d B=%d" % (n, b)
            try:
                # the reference FPS kernel needs tens of seconds beyond 64k points: comparator up to 16k
                fps_case(tag, b, n, m, x, with_ref=(n <= 16384))
                idx = tf_sampling.farthest_point_sample(m, x)
                q = tf_sampling.gather_point(x, idx)
                ball_case(tag, b, n, m, radius, ns, x, q, with_ref=(n <= 16384))
                nn_case(tag, b, n, m, x, q)
            except Exception as e:  # keep what has been measured
                record(op="sweep", case=tag, error=repr(e)[:300])
            del x

    dump()
    print("wrote", args.out)


if __name__ == "__main__":
    main()
