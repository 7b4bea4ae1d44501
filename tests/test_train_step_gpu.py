"""GPU: the optimizer step and the training loop around the path (train.py:387-388, 225-244)."""
import numpy as np
import pytest

from _util import to_cuda

pytestmark = pytest.mark.gpu


def adam_reference(p, g, m, v, lr, b1, b2, eps, t, gscale):
    """tf.train.AdamOptimizer in fp64: lr_t = lr*sqrt(1-b2^t)/(1-b1^t), "epsilon hat" added to sqrt(v)."""
    g = g.astype(np.float64) * gscale
    m = m + (g - m) * (1 - b1)
    v = v + (g * g - v) * (1 - b2)
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return p - lr_t * m / (np.sqrt(v) + eps), m, v


@pytest.mark.parametrize("t,gscale", [(1, 1.0), (7, 0.5), (1000, 0.125)])
def test_adam_step_matches_fp64_formula(cuda, t, gscale):
    import pn2_b200  # noqa: F401
    from pn2_b200._ffi import F32, call, ptr
    rs = np.random.RandomState(t)
    n = 100003
    p = rs.normal(size=n).astype(np.float32)
    g = (rs.normal(size=n) * rs.choice([1e-6, 1e-2, 10.0], size=n)).astype(np.float32)
    g[:100] = 0.0  # e.g. the biases in front of a train-mode BatchNorm
    m = (rs.normal(size=n) * 1e-2).astype(np.float32) if t > 1 else np.zeros(n, np.float32)
    v = (rs.random_sample(n) * 1e-3).astype(np.float32) if t > 1 else np.zeros(n, np.float32)
    # the hyper-parameters cross the C ABI as floats: the reference uses the same rounded values
    # (1 - float(0.999) differs from 0.001 by 1.3e-5 relative, far above the comparison tolerance)
    lr, b1, b2, eps = (float(np.float32(x)) for x in (1e-3, 0.9, 0.999, 1e-8))
    ep, em, ev = adam_reference(p.astype(np.float64), g, m.astype(np.float64), v.astype(np.float64),
                                lr, b1, b2, eps, t, gscale)
    dp, dg, dm, dv = to_cuda(p), to_cuda(g), to_cuda(m), to_cuda(v)
    call("pn2_adam_step", n, ptr(dp, F32), ptr(dg, F32), ptr(dm, F32), ptr(dv, F32), 1e-3, 0.9, 0.999, 1e-8,
         t, gscale)
    # fp32 evaluation: relative 2e-6, plus an absolute floor where m + (g-m)*(1-b1) cancels (|m| ~ 1e-2)
    np.testing.assert_allclose(dm.cpu().numpy(), em, rtol=2e-6, atol=1e-8)
    np.testing.assert_allclose(dv.cpu().numpy(), ev, rtol=2e-6, atol=1e-9)
    # the update is at most ~lr in magnitude; fp32 evaluation of m/(sqrt(v)+eps) is good to ~1e-6 relative
    np.testing.assert_allclose(dp.cpu().numpy(), ep, rtol=0, atol=5e-7)  # half an ulp of |p| < 8, plus the update
    assert np.array_equal(dp.cpu().numpy()[:100], p[:100]) or t > 1  # zero gradient, zero moments: no move


@pytest.mark.experimental
def test_two_training_steps_match_oracle(cuda):
    """Trainer.step twice (forward, loss, backward, Adam, BatchNorm moving statistics) against the fp64
    oracle driven by the same dropout masks and a numpy Adam.  EXPERIMENTAL marker: written without GPU
    time left; the tolerances on the second loss are a first guess."""
    import torch
    import pn2_b200  # noqa: F401
    from pn2_b200.train_step import Trainer
    from pn2_b200.util import tf_util
    from oracle import layers_ref as lr
    hp = {"use_color": 1, "batch_size": 2, "learning_rate": 0.001, "decay_step": 200000,
          "learning_rate_decay_rate": 0.7, "bn_init_decay": 0.5, "bn_decay_decay_rate": 0.5,
          "bn_decay_clip": 0.99, "l1_npoint": 256, "l1_radius": 0.1, "l1_nsample": 32, "l2_npoint": 64,
          "l2_radius": 0.2, "l2_nsample": 32, "l3_npoint": 16, "l3_radius": 0.4, "l3_nsample": 32,
          "l4_npoint": 8, "l4_radius": 0.8, "l4_nsample": 32}
    rs = np.random.RandomState(100)
    b, n = 2, 1024
    pc = np.concatenate([rs.random_sample((b, n, 3)), rs.random_sample((b, n, 3))], -1).astype(np.float32)
    labels = rs.randint(0, 9, (b, n)).astype(np.int32)
    smpw = rs.uniform(0.5, 2.0, (b, n)).astype(np.float32)
    params = lr.init_model_params(hp, 9, seed=1)
    tr = Trainer(hp, 9, device="cuda", seed=0, world_size=1)
    sd = {}
    for k, v in params.items():
        if k.endswith("/weights"):
            v = v.reshape((1,) + v.shape) if k.split("/")[0] in ("fc1", "fc2") else v.reshape((1, 1) + v.shape)
        sd[k] = v
    tr.store.load_state_dict(sd)
    lcg = lambda s: (s * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF  # noqa: E731
    seed0 = 1234
    tf_util.set_dropout_seed(seed0)
    d_pc, d_lab, d_w = to_cuda(pc), to_cuda(labels), to_cuda(smpw)
    # step 1 runs the forward twice (the first call creates + flattens the variables): its mask is seed 1
    seeds = [lcg(seed0), lcg(lcg(seed0))]
    m_adam = {k: np.zeros_like(v, np.float64) for k, v in params.items()}
    v_adam = {k: np.zeros_like(v, np.float64) for k, v in params.items()}
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    for step in (1, 2):
        loss = float(tr.step(d_pc, d_lab, d_w).item())
        mask = tf_util.dropout_mask(b * n * 128, 0.5, seeds[step - 1]).cpu().numpy().reshape(b, n, 128)
        ctx = lr.Ctx(p64, is_training=True, bn_decay=0.5, dropout_masks={"dp1": mask.astype(np.float64)})
        e_loss = lr.get_loss(lr.get_model(ctx, pc, 9, hp), labels, smpw)
        e_loss.backward()
        assert abs(loss - e_loss.item()) < (5e-5 if step == 1 else 2e-3), (step, loss, e_loss.item())
        for k, g in ctx.grads().items():
            p64[k], m_adam[k], v_adam[k] = adam_reference(p64[k], g, m_adam[k], v_adam[k], 1e-3, 0.9, 0.999,
                                                          1e-8, step, 1.0)
        for k, mv in ctx.new_moving.items():
            p64[k] = mv.astype(np.float64)
    torch.cuda.synchronize()
    got = tr.store.state_dict()
    for k in ("fc2/weights", "layer1/conv0/weights", "fa_layer4/conv_2/bn/gamma"):
        # two Adam steps move every weight by at most ~2*lr; the direction must agree with the oracle
        d_got = got[k].reshape(p64[k].shape).astype(np.float64) - params[k]
        d_exp = p64[k] - params[k]
        big = np.abs(d_exp) > 1.5e-3
        assert big.any() and np.mean(np.sign(d_got[big]) == np.sign(d_exp[big])) > 0.99, k
