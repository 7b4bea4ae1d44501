"""Dense layer helpers: the reference's util/tf_util.py surface on the sm_100a engine.

Kept names / signatures (reference file:line):
  conv2d(inputs, num_output_channels, kernel_size, scope, ...)   tf_util.py:128-204
  conv1d(inputs, num_output_channels, kernel_size, scope, ...)   tf_util.py:54-125
  batch_norm_for_conv2d / _conv1d / batch_norm_template          tf_util.py:555-629
  dropout(inputs, is_training, scope, keep_prob, noise_shape)    tf_util.py:646-665
Only what model.py reaches is in scope (1x1 kernels, VALID/SAME are identical for 1x1);
conv2d_transpose / conv3d / fully_connected / pools are not.

TensorFlow's variable scopes become a ``VariableStore``: ``variable_scope("layer1")`` +
``get_variable("weights", ...)`` yields the same names the reference checkpoints use
(``layer1/conv0/weights``, ``.../bn/gamma`` ...).  Variables are plain CUDA tensors (views into
one flat parameter buffer after ``flatten()``), NOT autograd leaves: the backward kernels
accumulate parameter gradients straight into the flat gradient buffer, which is what the
data-parallel all-reduce and the Adam kernel consume.

All arithmetic runs in libpn2_b200.so: Y = f(A) W + b as a GEMM with the previous layer's
BatchNorm+ReLU applied while loading A, batch statistics accumulated in the GEMM epilogue,
BN+ReLU(+max-pool over nsample) applied by one pass over the pre-activation tensor.
"""
import contextlib
import ctypes
import math

import torch

from .._ffi import F32, F64, I32, call, ptr, ptr_rows

BN_EPS = 1e-3  # tf.contrib.layers.batch_norm default epsilon (not in the reference source)
relu = "relu"  # stands in for tf.nn.relu as activation_fn


# ------------------------------------------------------------------------------------------
# variables
# ------------------------------------------------------------------------------------------
class Variable:
    __slots__ = ("name", "data", "grad", "trainable")

    def __init__(self, name, data, trainable=True):
        self.name, self.data, self.trainable = name, data, trainable
        self.grad = None

    def ensure_grad(self):
        if self.grad is None:
            self.grad = torch.zeros_like(self.data)
        return self.grad


class VariableStore:
    """name -> Variable, in creation order (== the reference graph's creation order)."""

    def __init__(self, device="cuda", seed=0):
        self.device = torch.device(device)
        self.vars = {}
        self.scope = []
        self.gen = torch.Generator(device="cpu")
        self.gen.manual_seed(seed)
        self.flat_params = self.flat_grads = None
        self.strict = False  # True: get_variable raises instead of creating (inference from a checkpoint)
        # tensor-core weight images built once per step (allocate_images / prepare_images)
        self.images, self.image_table, self.image_buf, self.images_fresh = {}, None, None, False
        # a leaf that requires grad, threaded through every layer Function so that autograd
        # visits the Function even when its data input does not require grad
        self.anchor = torch.zeros(1, device=self.device, requires_grad=True)

    # -- creation -----------------------------------------------------------------------
    def full_name(self, name):
        return "/".join(self.scope + [name])

    def get_variable(self, name, shape, init, trainable=True):
        full = self.full_name(name)
        v = self.vars.get(full)
        if v is None:
            if self.strict:
                raise KeyError("variable %s is not in the loaded checkpoint (strict store: nothing is "
                               "initialised silently)" % full)
            if callable(init):
                data = init(shape)
            else:
                data = torch.full(tuple(shape), float(init), dtype=F32)
            v = Variable(full, data.to(self.device, F32).contiguous(), trainable)
            self.vars[full] = v
            self.flat_params = self.flat_grads = None  # layout changed
        elif tuple(v.data.shape) != tuple(shape):
            # a checkpoint may hold a conv kernel as [k,n]: adopt the layer's rank, never another size
            if full.endswith("/weights") and v.data.numel() == math.prod(shape) and \
                    tuple(v.data.shape[-2:]) == tuple(shape[-2:]):
                v.data = v.data.view(*shape)
            else:
                raise ValueError("variable %s exists with shape %s, requested %s"
                                 % (full, tuple(v.data.shape), tuple(shape)))
        return v

    def xavier(self, shape):
        """tf.contrib.layers.xavier_initializer (uniform) for a [.., fan_in, fan_out] kernel."""
        fan_in, fan_out = shape[-2], shape[-1]
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(tuple(shape), generator=self.gen, dtype=F32) * 2 - 1) * lim

    def truncated_normal(self, stddev):
        def init(shape):
            t = torch.empty(tuple(shape), dtype=F32)
            torch.nn.init.trunc_normal_(t, 0.0, stddev, -2 * stddev, 2 * stddev, generator=self.gen)
            return t
        return init

    # -- flat buffers ---------------------------------------------------------------------
    def trainable(self):
        return [v for v in self.vars.values() if v.trainable]

    def flatten(self):
        """Move every trainable variable into one flat fp32 buffer (+ a flat gradient buffer).
        The flat gradient buffer is the single all-reduce message of the DP step."""
        tv = self.trainable()
        total = sum(v.data.numel() for v in tv)
        flat = torch.empty(total, dtype=F32, device=self.device)
        grads = torch.zeros(total, dtype=F32, device=self.device)
        off = 0
        for v in tv:
            n = v.data.numel()
            flat[off:off + n].copy_(v.data.reshape(-1))
            v.data = flat[off:off + n].view(v.data.shape)
            v.grad = grads[off:off + n].view(v.data.shape)
            off += n
        self.flat_params, self.flat_grads = flat, grads
        return flat, grads

    # -- tensor-core weight images: ONE preparation launch per step --------------------------------
    def allocate_images(self):
        """Give every conv kernel a persistent pair of 3xTF32 weight images (forward and dgrad
        orientation) plus the device table pn2_linear_prepare walks.  Call after flatten()."""
        import ctypes
        import numpy as np
        from .._ffi import lib
        L = lib()
        plan, off = [], 0
        for name, v in self.vars.items():
            if not name.endswith("/weights"):
                continue
            k, n = int(v.data.shape[-2]), int(v.data.shape[-1])
            for dg in (0, 1):
                nb = int(L.pn2_linear_image_bytes(k, n, dg))
                if nb <= 0:
                    continue
                plan.append((name, k, n, dg, off, nb))
                off += (nb + 255) // 256 * 256
        self.images, self.image_table, self.images_fresh = {}, None, False
        if not plan:
            return
        self.image_buf = torch.empty(off // 4, dtype=F32, device=self.device)
        table = np.zeros((len(plan), 64), np.uint8)
        base = self.image_buf.data_ptr()
        for i, (name, k, n, dg, o, nb) in enumerate(plan):
            img = self.image_buf[o // 4:(o + nb) // 4]
            rc = L.pn2_linear_image_describe(k, n, dg, ctypes.c_void_p(self.vars[name].data.data_ptr()),
                                             ctypes.c_void_p(base + o),
                                             ctypes.c_void_p(table[i].ctypes.data))
            if rc != 0:
                raise RuntimeError("pn2_linear_image_describe(%s) failed: %d" % (name, rc))
            self.images.setdefault(name, [None, None])[dg] = img
        self.image_table = torch.as_tensor(table).to(self.device)

    def prepare_images(self):
        """Rebuild every weight image from the current weights (one launch)."""
        if self.image_table is not None:
            call("pn2_linear_prepare", self.image_table.shape[0], ptr(self.image_table, torch.uint8))
            self.images_fresh = True

    def zero_grad(self):
        if self.flat_grads is not None:
            self.flat_grads.zero_()
        else:
            for v in self.trainable():
                if v.grad is not None:
                    v.grad.zero_()

    def state_dict(self):
        return {k: v.data.detach().cpu().numpy().copy() for k, v in self.vars.items()}

    MODEL_SUFFIXES = ("/weights", "/biases", "/bn/gamma", "/bn/beta", "/bn/moving_mean", "/bn/moving_variance")

    def load_state_dict(self, sd):
        """Load ``name -> array`` (a TensorFlow checkpoint of this network, or state_dict()).

        Shapes must match the variable exactly, except that a conv kernel may come as [k,n], [1,k,n]
        or [1,1,k,n] (same element order).  Keys that are not model variables -- the optimizer slots
        and counters of a full TF checkpoint (``*/Adam``, ``*/Adam_1``, ``beta1_power``, the global
        step ...) -- are skipped and returned, never registered as variables."""
        skipped = []
        for k, arr in sd.items():
            if not k.endswith(self.MODEL_SUFFIXES):
                skipped.append(k)
                continue
            t = torch.as_tensor(arr, dtype=F32)
            if k in self.vars:
                dst = self.vars[k].data
                same = tuple(t.shape) == tuple(dst.shape)
                kernel = k.endswith("/weights") and t.dim() >= 2 and tuple(t.shape[-2:]) == tuple(dst.shape[-2:]) \
                    and t.numel() == dst.numel()
                if not (same or kernel):
                    raise ValueError("checkpoint variable %s has shape %s, the model's has %s"
                                     % (k, tuple(t.shape), tuple(dst.shape)))
                dst.copy_(t.reshape(dst.shape).to(self.device))
                self.images_fresh = False
            else:
                trainable = not (k.endswith("moving_mean") or k.endswith("moving_variance"))
                self.vars[k] = Variable(k, t.to(self.device).contiguous(), trainable)
                self.flat_params = self.flat_grads = None
        return skipped


_store = None


def default_store():
    global _store
    if _store is None:
        _store = VariableStore()
    return _store


def set_default_store(store):
    global _store
    _store = store
    return store


@contextlib.contextmanager
def variable_scope(name):
    st = default_store()
    st.scope.append(name)
    try:
        yield name
    finally:
        st.scope.pop()


def _as_bool(x):
    if isinstance(x, torch.Tensor):
        return bool(x.item())
    return bool(x)


# ------------------------------------------------------------------------------------------
# the shared-MLP chain: [1x1 conv + bias (+BN) (+ReLU)] x L (+ max-pool over nsample)
# ------------------------------------------------------------------------------------------
class LayerSpec:
    """Variables and flags of one conv layer (created by ``make_layer`` inside its scope)."""
    __slots__ = ("w", "b", "gamma", "beta", "mm", "mv", "bn", "relu", "rank4", "k", "n", "img")


def make_layer(scope, k, n, bn, act, use_xavier=True, stddev=1e-3, kernel_rank=4):
    st = default_store()
    L = LayerSpec()
    with variable_scope(scope):
        init = st.xavier if use_xavier else st.truncated_normal(stddev)
        kshape = [1, 1, k, n] if kernel_rank == 4 else [1, k, n]
        L.w = st.get_variable("weights", kshape, init)
        L.b = st.get_variable("biases", [n], 0.0)
        L.bn, L.relu, L.rank4, L.k, L.n = bool(bn), act is not None, kernel_rank == 4, k, n
        # persistent weight images (forward, dgrad) when the store prepared them for this pass
        L.img = st.images.get(L.w.name) if st.images_fresh else None
        if bn:
            with variable_scope("bn"):
                L.gamma = st.get_variable("gamma", [n], 1.0)
                L.beta = st.get_variable("beta", [n], 0.0)
                L.mm = st.get_variable("moving_mean", [n], 0.0, trainable=False)
                L.mv = st.get_variable("moving_variance", [n], 1.0, trainable=False)
        else:
            L.gamma = L.beta = L.mm = L.mv = None
    return L


class _ZeroArena:
    """fp64 scratch that must start at zero (BatchNorm statistics, backward reductions): slices of one
    buffer that a training step clears with a single memset (`reset`) instead of one fill kernel per
    layer.  Outside a Trainer step (arena not armed or exhausted) `take` falls back to torch.zeros."""

    def __init__(self):
        self.buf, self.off, self.armed = None, 0, False

    def reset(self, device, capacity=1 << 16):
        if self.buf is None or self.buf.device != torch.device(device) or self.buf.numel() < capacity:
            self.buf = torch.empty(capacity, dtype=F64, device=device)
        self.buf.zero_()
        self.off, self.armed = 0, True

    def disarm(self):
        self.armed = False

    def take(self, n, device):
        n8 = (n + 1) // 2 * 2  # keep 16-byte alignment
        if not self.armed or self.buf is None or self.buf.device != torch.device(device) \
                or self.off + n8 > self.buf.numel():
            return torch.zeros(n, dtype=F64, device=device)
        out = self.buf[self.off:self.off + n]
        self.off += n8
        return out


zero_arena = _ZeroArena()


# ---- weight gradients on a second stream -------------------------------------------------------------------
# In the backward pass of a shared-MLP chain the weight gradient of a layer (dW = A^T dY) and everything else
# (dX = dY W^T, then the BatchNorm backward of the layer below) are independent once dY exists.  With a side
# stream set here, the chains issue every pn2_linear_wgrad on it (ordered behind the kernel that produced dY by an
# event) with the persistent kernel sized for `wgrad_sms` SMs, and the input-gradient GEMMs for the rest of the
# device, so the tensor-core wgrad runs next to the HBM-bound BatchNorm kernels of the main chain instead of
# after them.  The owner (train_step.Trainer) joins the side stream before the gradients are consumed.
_wgrad_side = [None, 0]


def set_wgrad_stream(stream, wgrad_sms=0):
    """stream: torch.cuda.Stream for the weight-gradient launches (None = same stream as everything else)."""
    _wgrad_side[0], _wgrad_side[1] = stream, int(wgrad_sms)


# Event the first GEMM of a pass has to wait for (the weight images of the step, prepared on a side stream while
# the main stream already gathers the first layer's input); consumed by the first shared-MLP chain that runs.
_images_event = [None]


def set_images_event(ev):
    _images_event[0] = ev


def _sm_budget(sms):
    from .. import _ffi
    _ffi.lib().pn2_set_sm_budget(int(sms))

# Test hook: a dict here makes every shared-MLP chain record its ReLU masks ("<scope>/relu_mask", uint8
# (M,N)) and max-pool winners ("<scope>/argmax", int32 (G,N)).  The fp64 oracle then differentiates the
# SAME piecewise-linear function (an element whose pre-activation is within fp32 rounding of zero may
# otherwise sit on the other side of the kink), which is what allows a strict gradient tolerance.
debug_capture = None

# While set, train-mode BatchNorm normalises with the batch statistics but leaves moving_mean /
# moving_variance alone: passes that exist only to create variables or to warm up a CUDA-graph
# capture must not add EMA updates the reference never applies (tf_util.py:572-581 updates once
# per session.run of the train op).
_freeze_moving = [False]


@contextlib.contextmanager
def frozen_moving_stats():
    prev = _freeze_moving[0]
    _freeze_moving[0] = True
    try:
        yield
    finally:
        _freeze_moving[0] = prev

_ws_cache = {}


def _workspace(L, dev):
    """Caller-owned scratch for the tensor-core path (3xTF32 weight image), one per layer shape."""
    key = (L.k, L.n, str(dev))
    ws = _ws_cache.get(key)
    if ws is None:
        from .._ffi import lib
        nbytes = int(lib().pn2_linear_workspace_bytes(L.k, L.n))
        ws = torch.empty(max(nbytes // 4, 4), dtype=F32, device=dev)
        _ws_cache[key] = ws
    return ws


IMAGE_READY = 16  # PN2_GEMM_IMAGE_READY (include/pn2_b200.h)


class BnFinalize(ctypes.Structure):
    """pn2_bn_finalize (include/pn2_b200.h): the train-mode BatchNorm finalize the GEMM's last CTA performs."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("gamma", "beta", "moving_mean", "moving_var", "scale", "shift",
                                               "saved", "counter")] + \
               [("eps", ctypes.c_float), ("decay", ctypes.c_float), ("unbiased_moving", ctypes.c_int)]


def _weight_image(L, dgrad, dev, gemm_mode):
    """(workspace, mode) of one GEMM call: the layer's persistent image when the store prepared the
    images for this pass (mode + IMAGE_READY), else the shared per-shape scratch (image rebuilt per call)."""
    if gemm_mode == 0:
        return None, 0
    if L.img is not None and L.img[dgrad] is not None:
        return L.img[dgrad], gemm_mode + IMAGE_READY
    return _workspace(L, dev), gemm_mode


class _MLPChain(torch.autograd.Function):
    """x (M,K0) -> out (M,N_last) or, with pool_ns>0, (M/pool_ns, N_last) max-pooled.

    Forward keeps only the pre-activation tensors Y_i; BatchNorm+ReLU of layer i is applied
    on the fly when layer i+1 loads its A operand, and by the final affine(+pool) pass."""

    @staticmethod
    def forward(ctx, x, anchor, layers, is_training, bn_decay, pool_ns, gemm_mode, dx_cols=None):
        M, K0 = x.shape
        dev = x.device
        if _images_event[0] is not None:
            torch.cuda.current_stream(dev).wait_event(_images_event[0])
            _images_event[0] = None
        # rows may be padded (a column slice of a wider buffer: the SA/FP concat buffers are
        # allocated with a 16-byte aligned row pitch for the TMA tensor maps)
        if not (x.stride(1) == 1 and x.stride(0) >= K0):
            x = x.contiguous()
        a_ptr, lda = ptr_rows(x, F32)
        decay = 0.9 if bn_decay is None else float(bn_decay)
        a_sc, a_sh, a_relu = None, None, 0
        Ys, scs, shs, saveds = [], [], [], []
        for L in layers:
            N = L.n
            Y = torch.empty((M, N), dtype=F32, device=dev)
            use_stats = L.bn and is_training
            stats = zero_arena.take(2 * N, dev) if use_stats else None
            ws, mode_i = _weight_image(L, 0, dev, gemm_mode)
            sc = sh = saved = None
            if use_stats:
                # GEMM + train-mode BatchNorm finalize in one entry point (the GEMM's last CTA finalises)
                sc = torch.empty(N, dtype=F32, device=dev)
                sh = torch.empty(N, dtype=F32, device=dev)
                saved = torch.empty(2 * N, dtype=F32, device=dev)
                counter = zero_arena.take(1, dev)
                upd = not _freeze_moving[0]
                fin = BnFinalize(L.gamma.data.data_ptr(), L.beta.data.data_ptr(),
                                 L.mm.data.data_ptr() if upd else None, L.mv.data.data_ptr() if upd else None,
                                 sc.data_ptr(), sh.data_ptr(), saved.data_ptr(), counter.data_ptr(), BN_EPS, decay,
                                 1 if L.rank4 else 0)
                call("pn2_linear_fwd_bn", M, L.k, N, a_ptr, lda, ptr(a_sc, F32, True), ptr(a_sh, F32, True),
                     a_relu, ptr(L.w.data, F32), ptr(L.b.data, F32), ptr(Y, F32), ptr(stats, F64),
                     ctypes.byref(fin), ptr(ws, F32, True), 0 if ws is None else ws.numel() * 4, mode_i)
            else:
                call("pn2_linear_fwd", M, L.k, N, a_ptr, lda, ptr(a_sc, F32, True),
                     ptr(a_sh, F32, True), a_relu, ptr(L.w.data, F32), ptr(L.b.data, F32), ptr(Y, F32),
                     ptr(stats, F64, True), ptr(ws, F32, True), 0 if ws is None else ws.numel() * 4,
                     mode_i)
                if L.bn:  # inference: moving statistics
                    sc = torch.empty(N, dtype=F32, device=dev)
                    sh = torch.empty(N, dtype=F32, device=dev)
                    call("pn2_bn_eval_affine", N, ptr(L.gamma.data, F32), ptr(L.beta.data, F32),
                         ptr(L.mm.data, F32), ptr(L.mv.data, F32), BN_EPS, ptr(sc, F32),
                         ptr(sh, F32))
            if not L.bn and L.relu:
                sc = torch.ones(N, dtype=F32, device=dev)
                sh = torch.zeros(N, dtype=F32, device=dev)
            Ys.append(Y)
            scs.append(sc)
            shs.append(sh)
            saveds.append(saved)
            a_ptr, lda, a_sc, a_sh, a_relu = ptr(Y, F32), N, sc, sh, 1 if L.relu else 0
        L = layers[-1]
        N = L.n
        arg = None
        if pool_ns:
            G = M // pool_ns
            out = torch.empty((G, N), dtype=F32, device=dev)
            arg = torch.empty((G, N), dtype=I32, device=dev)
            call("pn2_affine_act_maxpool", G, pool_ns, N, ptr(Ys[-1], F32), ptr(scs[-1], F32, True),
                 ptr(shs[-1], F32, True), 1 if L.relu else 0, ptr(out, F32), ptr(arg, I32))
        elif scs[-1] is None and not L.relu:
            out = Ys[-1]
        else:
            out = torch.empty((M, N), dtype=F32, device=dev)
            call("pn2_affine_act", M, N, ptr(Ys[-1], F32), ptr(scs[-1], F32, True),
                 ptr(shs[-1], F32, True), 1 if L.relu else 0, ptr(out, F32), N)
        if debug_capture is not None:  # parity tests: the ReLU decisions and pooling winners of this chain
            for L_, Y_, sc_, sh_ in zip(layers, Ys, scs, shs):
                if L_.relu:
                    mk = torch.empty((M, L_.n), dtype=torch.uint8, device=dev)
                    call("pn2_relu_mask", M, L_.n, ptr(Y_, F32), ptr(sc_, F32, True), ptr(sh_, F32, True),
                         ptr(mk, torch.uint8))
                    debug_capture[L_.w.name[:-len("/weights")] + "/relu_mask"] = mk
            if arg is not None:
                debug_capture[L.w.name[:-len("/weights")] + "/argmax"] = arg
        ctx.layers, ctx.pool_ns, ctx.is_training, ctx.gemm_mode = layers, pool_ns, is_training, gemm_mode
        ctx.dx_cols = dx_cols
        ctx.x, ctx.Ys, ctx.scs, ctx.shs, ctx.saveds, ctx.arg = x, Ys, scs, shs, saveds, arg
        return out

    @staticmethod
    def backward(ctx, d_out):
        layers, pool_ns = ctx.layers, ctx.pool_ns
        if not ctx.is_training and any(L.bn for L in layers):
            raise NotImplementedError("backward through inference-mode BatchNorm is not supported")
        x, Ys, scs, shs, saveds = ctx.x, ctx.Ys, ctx.scs, ctx.shs, ctx.saveds
        M = x.shape[0]
        dev = x.device
        up = d_out.contiguous()
        for i in range(len(layers) - 1, -1, -1):
            L = layers[i]
            N, Y = L.n, Ys[i]
            relu_i = 1 if L.relu else 0
            pooled = pool_ns and i == len(layers) - 1
            red = zero_arena.take(2 * N, dev) if L.bn else None
            dg = ptr(L.gamma.ensure_grad(), F32) if L.bn else None
            db_ = ptr(L.beta.ensure_grad(), F32) if L.bn else None
            if pooled:
                G = M // pool_ns
                if L.bn:
                    call("pn2_bn_bwd_reduce_pool", G, pool_ns, N, ptr(up, F32), ptr(ctx.arg, I32),
                         ptr(Y, F32), ptr(scs[i], F32), ptr(shs[i], F32), ptr(saveds[i], F32),
                         relu_i, ptr(red, F64))
                dY = torch.empty((M, N), dtype=F32, device=dev)
                call("pn2_bn_bwd_apply_pool", G, pool_ns, N, ptr(up, F32), ptr(ctx.arg, I32),
                     ptr(Y, F32), ptr(scs[i], F32, True), ptr(shs[i], F32, True),
                     ptr(saveds[i], F32, True), ptr(L.gamma.data, F32) if L.bn else None, relu_i,
                     1 if L.bn else 0, ptr(red, F64, True), ptr(dY, F32), dg, db_)
            elif L.bn or L.relu:
                if L.bn:
                    call("pn2_bn_bwd_reduce", M, N, ptr(up, F32), N, ptr(Y, F32), ptr(scs[i], F32),
                         ptr(shs[i], F32), ptr(saveds[i], F32), relu_i, ptr(red, F64))
                dY = torch.empty((M, N), dtype=F32, device=dev)
                call("pn2_bn_bwd_apply", M, N, ptr(up, F32), N, ptr(Y, F32), ptr(scs[i], F32, True),
                     ptr(shs[i], F32, True), ptr(saveds[i], F32, True),
                     ptr(L.gamma.data, F32) if L.bn else None, relu_i, 1 if L.bn else 0,
                     ptr(red, F64, True), ptr(dY, F32), dg, db_)
            else:
                dY = up
            if i == 0:
                (a_ptr, lda), a_sc, a_sh, a_relu = ptr_rows(x, F32), None, None, 0
            else:
                P = layers[i - 1]
                a_ptr, lda, a_sc, a_sh, a_relu = ptr(Ys[i - 1], F32), P.n, scs[i - 1], shs[i - 1], \
                    1 if P.relu else 0
            # A bias that feeds a train-mode BatchNorm has an exactly zero gradient (BN removes the
            # column mean, so sum_rows dY == 0); it is left at zero instead of accumulating the
            # fp32 rounding noise of an M-term sum.
            db = None if L.bn else ptr(L.b.ensure_grad(), F32)
            L.b.ensure_grad()
            dw = ptr(L.w.ensure_grad(), F32)
            side, side_sms = _wgrad_side
            if side is None:
                call("pn2_linear_wgrad", M, L.k, N, a_ptr, lda, ptr(a_sc, F32, True),
                     ptr(a_sh, F32, True), a_relu, ptr(dY, F32), dw, db, ctx.gemm_mode)
            else:
                main = torch.cuda.current_stream(dev)
                side.wait_stream(main)  # dY (and, first time round, the zeroed gradient buffer) is complete
                total = torch.cuda.get_device_properties(dev).multi_processor_count
                # the first layer of a chain whose input needs no gradient ends the backward pass (SA1): nothing runs
                # next to its weight gradient, so it gets the whole device
                last = i == 0 and not ctx.needs_input_grad[0] and _os.environ.get("PN2_WGRAD_TAIL_FULL", "1") != "0"
                with torch.cuda.stream(side):
                    _sm_budget(0 if last else side_sms)
                    try:
                        call("pn2_linear_wgrad", M, L.k, N, a_ptr, lda, ptr(a_sc, F32, True),
                             ptr(a_sh, F32, True), a_relu, ptr(dY, F32), dw, db, ctx.gemm_mode)
                    finally:
                        _sm_budget(total - side_sms if 0 < side_sms < total else 0)
                # buffers of this pass that the side stream reads after this function has dropped them
                for t in (dY, x if i == 0 else Ys[i - 1], a_sc, a_sh):
                    if t is not None:
                        t.record_stream(side)
            if i > 0 or ctx.needs_input_grad[0]:
                dX = torch.empty((M, L.k), dtype=F32, device=dev)
                k0, k1 = (0, L.k) if (i > 0 or ctx.dx_cols is None) else ctx.dx_cols
                if (k0, k1) == (0, L.k):
                    ws, mode_i = _weight_image(L, 1, dev, ctx.gemm_mode)
                    call("pn2_linear_dgrad", M, L.k, N, ptr(dY, F32), ptr(L.w.data, F32), ptr(dX, F32),
                         L.k, ptr(ws, F32, True), 0 if ws is None else ws.numel() * 4, mode_i)
                elif k1 > k0:
                    # only columns [k0, k1) of the input gradient are consumed (the caller's xyz / skip columns need
                    # none): dX[:, k] = dY W[k, :]^T is independent per column, so the same entry point runs on
                    # the row range k0..k1 of W and writes into the column range of dX -- FP4's first layer
                    # (131 -> 128 columns) drops its second 128-wide column block, 86 -> 32 us
                    ws, mode_i = _weight_image(L, 1, dev, ctx.gemm_mode)
                    if not (mode_i >= IMAGE_READY - 1 and k0 == 0 and k1 % 128 == 0 and L.k > 128):
                        # a prefix of whole 128-column blocks reuses the prepared image; anything else gets a
                        # per-call image of just those rows
                        ws, mode_i = (None, 0) if ctx.gemm_mode == 0 else (_workspace(L, dev), ctx.gemm_mode)
                    call("pn2_linear_dgrad", M, k1 - k0, N, ptr(dY, F32),
                         ctypes.c_void_p(L.w.data.data_ptr() + 4 * k0 * N),
                         ctypes.c_void_p(dX.data_ptr() + 4 * k0), L.k, ptr(ws, F32, True),
                         0 if ws is None else ws.numel() * 4, mode_i)
                up = dX
            else:
                up = None
        ctx.Ys = ctx.x = None
        if _wgrad_side[0] is not None:
            _sm_budget(0)
        return up, None, None, None, None, None, None, None


import os as _os

# -1 auto (tcgen05 3xTF32 where the shape allows, else fp32 CUDA cores), 0 exact fp32 CUDA-core
# kernel only, 1 force tcgen05.  PN2_GEMM_MODE overrides the default.
GEMM_MODE = int(_os.environ.get("PN2_GEMM_MODE", "-1"))


def mlp_chain(x2d, layers, is_training, bn_decay, pool_ns=0, dx_cols=None):
    """dx_cols = (k0, k1): only these columns of the gradient w.r.t. x2d will be read by the caller (the rest of
    the returned gradient is left unwritten)."""
    return _MLPChain.apply(x2d, default_store().anchor, layers, _as_bool(is_training), bn_decay,
                           int(pool_ns), GEMM_MODE, dx_cols)


# ------------------------------------------------------------------------------------------
# reference-named layer functions
# ------------------------------------------------------------------------------------------
def _check_1x1(kernel_size, stride):
    ks = list(kernel_size) if isinstance(kernel_size, (list, tuple)) else [kernel_size]
    st = list(stride) if isinstance(stride, (list, tuple)) else [stride]
    if any(k != 1 for k in ks) or any(s != 1 for s in st):
        raise NotImplementedError("only 1x1 kernels with stride 1 are on the SA/FP hot path")


def conv2d(inputs, num_output_channels, kernel_size, scope, stride=[1, 1], padding="SAME",
           data_format="NHWC", use_xavier=True, stddev=1e-3, weight_decay=None,
           activation_fn=relu, bn=False, bn_decay=None, is_training=None):
    """1x1 2-D convolution + bias (+BN) (+ReLU) on a (B,H,W,C) tensor (tf_util.py:128-204)."""
    _check_1x1(kernel_size, stride)
    if data_format != "NHWC":
        raise NotImplementedError("channels-last only (use_nchw is a layout no-op here)")
    if weight_decay is not None:
        raise NotImplementedError("weight_decay is never set on the SA/FP path")
    shp = inputs.shape
    L = make_layer(scope, shp[-1], num_output_channels, bn, activation_fn, use_xavier, stddev, 4)
    y = mlp_chain(inputs.reshape(-1, shp[-1]), [L], True if is_training is None else is_training,
                  bn_decay)
    return y.view(*shp[:-1], num_output_channels)


def conv1d(inputs, num_output_channels, kernel_size, scope, stride=1, padding="SAME",
           data_format="NHWC", use_xavier=True, stddev=1e-3, weight_decay=None,
           activation_fn=relu, bn=False, bn_decay=None, is_training=None):
    """k=1 1-D convolution + bias (+BN) (+ReLU) on a (B,L,C) tensor (tf_util.py:54-125)."""
    _check_1x1(kernel_size, stride)
    if data_format != "NHWC":
        raise NotImplementedError("channels-last only")
    if weight_decay is not None:
        raise NotImplementedError("weight_decay is never set on the SA/FP path")
    shp = inputs.shape
    L = make_layer(scope, shp[-1], num_output_channels, bn, activation_fn, use_xavier, stddev, 3)
    y = mlp_chain(inputs.reshape(-1, shp[-1]), [L], True if is_training is None else is_training,
                  bn_decay)
    return y.view(*shp[:-1], num_output_channels)


class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, keep_prob, seed, seed_dev):
        x = x.contiguous()
        out = torch.empty_like(x)
        call("pn2_dropout", x.numel(), ptr(x, F32), float(keep_prob), int(seed),
             ptr(seed_dev, torch.int64, True), ptr(out, F32))
        ctx.keep_prob, ctx.seed, ctx.seed_dev = keep_prob, seed, seed_dev
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        out = torch.empty_like(g)
        call("pn2_dropout", g.numel(), ptr(g, F32), float(ctx.keep_prob), int(ctx.seed),
             ptr(ctx.seed_dev, torch.int64, True), ptr(out, F32))
        return out, None, None, None


_dropout_seed = [0x5EED]
_dropout_seed_dev = [None]  # optional device-resident increment (CUDA-graph replay, see train_step)


def set_dropout_seed(seed):
    _dropout_seed[0] = int(seed)


def set_dropout_seed_device(t):
    """t: 1-element int64 CUDA tensor added to the host seed inside the kernel (or None)."""
    _dropout_seed_dev[0] = t


def dropout_mask(numel, keep_prob, seed, device="cuda"):
    """The 0/1 keep mask ``dropout`` uses for (seed, element index) -- test hook."""
    m = torch.empty(numel, dtype=torch.uint8, device=device)
    call("pn2_dropout_mask", numel, float(keep_prob), int(seed), ptr(m, torch.uint8))
    return m


def dropout(inputs, is_training, scope, keep_prob=0.5, noise_shape=None):
    """tf_util.py:646-665: kept values are scaled by 1/keep_prob in training, identity else."""
    if noise_shape is not None:
        raise NotImplementedError("noise_shape is never set on the SA/FP path")
    if not _as_bool(is_training):
        return inputs
    seed = _dropout_seed[0]
    if _dropout_seed_dev[0] is None:  # eager mode: advance the host seed every call
        _dropout_seed[0] = (seed * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
    return _Dropout.apply(inputs, keep_prob, seed, _dropout_seed_dev[0])
