"""ctypes bindings for the parity oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Two libraries are wrapped with the same Python surface:

* ``Oracle``   -> oracle/liboracle.so, the plain-C restatement (oracle/fm_oracle.c);
* ``RefOracle``-> oracle/_ref/libdifacto_ref.so, the unmodified reference compiled by
  ``make -C oracle ref`` (present only where /root/reference was available at build time).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (difacto_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ORACLE = os.path.join(HERE, "liboracle.so")
LIB_REF = os.path.join(HERE, "_ref", "libdifacto_ref.so")

K_FEA_COUNT, K_WEIGHT, K_GRADIENT = 1, 2, 3  # include/difacto/store.h:33-35


def build(ref=True, quiet=True):
    """compile liboracle.so and, when /root/reference exists, _ref/libdifacto_ref.so"""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", HERE, "all"], stdout=out)
    if ref and os.path.isdir(os.environ.get("DIFACTO_REF", "/root/reference")):
        subprocess.check_call(["make", "-C", HERE, "ref", "-j8"], stdout=out)
        # the reference's own SGDLearner linked with the GPU plugins of ../integration (the drop-in proof; a test target)
        if os.path.exists(os.path.join(HERE, "..", "difacto_b200", "lib", "libdifacto_b200.so")):
            subprocess.check_call(["make", "-C", HERE, "ref_gpu", "-j8"], stdout=out)


def have_ref():
    return os.path.exists(LIB_REF)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class OrcParam(C.Structure):
    _fields_ = [("l1", C.c_float), ("l2", C.c_float), ("V_l2", C.c_float),
                ("lr", C.c_float), ("lr_beta", C.c_float), ("V_lr", C.c_float),
                ("V_lr_beta", C.c_float), ("V_init_scale", C.c_float),
                ("V_dim", C.c_int), ("V_threshold", C.c_int), ("seed", C.c_uint)]


PARAM_KEYS = ["l1", "l2", "V_l2", "lr", "lr_beta", "V_lr", "V_lr_beta", "V_init_scale",
              "V_dim", "V_threshold", "seed"]


def _load_oracle():
    if not os.path.exists(LIB_ORACLE):
        build(ref=False)
    L = C.CDLL(LIB_ORACLE)
    vp, sz, u64 = C.c_void_p, C.c_size_t, C.c_uint64
    L.orc_reverse_bytes.restype = u64
    L.orc_reverse_bytes.argtypes = [u64]
    L.orc_rand_r.restype = C.c_int
    L.orc_rand_r.argtypes = [C.POINTER(C.c_uint)]
    L.orc_owner.restype = C.c_uint32
    L.orc_owner.argtypes = [u64, C.c_uint32]
    L.orc_localize.restype = sz
    L.orc_localize.argtypes = [sz, vp, vp, u64, vp, vp, vp]
    L.orc_model_create.restype = vp
    L.orc_model_create.argtypes = [C.POINTER(OrcParam)]
    L.orc_model_free.argtypes = [vp]
    L.orc_model_size.restype = sz
    L.orc_model_size.argtypes = [vp]
    L.orc_model_seed.restype = C.c_uint
    L.orc_model_seed.argtypes = [vp]
    L.orc_get.restype = sz
    L.orc_get.argtypes = [vp, vp, sz, vp, vp, C.POINTER(sz)]
    L.orc_update_feacnt.argtypes = [vp, vp, sz, vp]
    L.orc_update_grad.restype = C.c_int
    L.orc_update_grad.argtypes = [vp, vp, sz, vp, sz, vp, sz]
    L.orc_model_lookup.restype = C.c_int
    L.orc_model_lookup.argtypes = [vp, u64, vp, vp]
    L.orc_get_pos.argtypes = [vp, sz, vp, vp]
    L.orc_fm_predict.argtypes = [C.c_int, sz, vp, vp, vp, vp, sz, vp, vp, sz, vp, vp]
    L.orc_evaluate.restype = C.c_float
    L.orc_evaluate.argtypes = [vp, vp, sz]
    L.orc_fm_calc_grad.argtypes = [C.c_int, sz, vp, vp, vp, vp, vp, sz, vp, vp, sz, vp, vp]
    L.orc_auc.restype = C.c_float
    L.orc_auc.argtypes = [vp, vp, sz]
    L.orc_penalty.restype = C.c_float
    L.orc_penalty.argtypes = [C.POINTER(OrcParam), vp, sz, vp, vp, sz]
    L.orc_sgd_step.argtypes = [vp, sz, vp, vp, vp, vp, C.c_int, C.c_int, vp]
    L.orc_param_default.argtypes = [C.POINTER(OrcParam)]
    return L


_ORC = None


def orc():
    global _ORC
    if _ORC is None:
        _ORC = _load_oracle()
    return _ORC


def make_param(**kw):
    p = OrcParam()
    orc().orc_param_default(C.byref(p))
    for k, v in kw.items():
        if k not in PARAM_KEYS:
            raise KeyError(k)
        setattr(p, k, v)
    return p


def reverse_bytes(x):
    x = np.asarray(x, dtype=np.uint64)
    f = orc().orc_reverse_bytes
    if x.ndim == 0:
        return np.uint64(f(int(x)))
    return np.array([f(int(v)) for v in x.ravel()], dtype=np.uint64).reshape(x.shape)


def reverse_bytes_np(x):
    """vectorised numpy restatement of include/difacto/base.h:39-51 (nibble reversal)"""
    x = np.asarray(x, dtype=np.uint64).copy()
    x = (x << np.uint64(32)) | (x >> np.uint64(32))
    x = ((x & np.uint64(0x0000FFFF0000FFFF)) << np.uint64(16)) | ((x & np.uint64(0xFFFF0000FFFF0000)) >> np.uint64(16))
    x = ((x & np.uint64(0x00FF00FF00FF00FF)) << np.uint64(8)) | ((x & np.uint64(0xFF00FF00FF00FF00)) >> np.uint64(8))
    x = ((x & np.uint64(0x0F0F0F0F0F0F0F0F)) << np.uint64(4)) | ((x & np.uint64(0xF0F0F0F0F0F0F0F0)) >> np.uint64(4))
    return x


def owner(keys, S):
    keys = np.asarray(keys, dtype=np.uint64)
    width = np.uint64(0xFFFFFFFFFFFFFFFF // S)
    o = keys // width
    return np.minimum(o, np.uint64(S - 1)).astype(np.uint32)


def localize(offset, index, max_index=0xFFFFFFFFFFFFFFFF, want_cnt=True):
    """Localizer::Compact -> (local_index u32[nnz], keys u64[U] ascending reversed, cnt f32[U])"""
    offset = np.ascontiguousarray(offset, dtype=np.uint64)
    index = np.ascontiguousarray(index, dtype=np.uint64)
    nrows = len(offset) - 1
    nnz = int(offset[-1]) if nrows > 0 else 0
    lidx = np.zeros(max(nnz, 1), np.uint32)
    keys = np.zeros(max(nnz, 1), np.uint64)
    cnt = np.zeros(max(nnz, 1), np.float32)
    U = orc().orc_localize(nrows, _p(offset), _p(index), max_index, _p(lidx), _p(keys),
                           _p(cnt) if want_cnt else None)
    return lidx[:nnz], keys[:U].copy(), (cnt[:U].copy() if want_cnt else None)


def get_pos(lens):
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    w_pos = np.zeros(len(lens), np.int32)
    V_pos = np.zeros(len(lens), np.int32)
    orc().orc_get_pos(_p(lens), len(lens), _p(w_pos), _p(V_pos))
    return w_pos, V_pos


def fm_predict(V_dim, offset, lidx, value, weights, w_pos=None, V_pos=None, want_xv=False):
    offset = np.ascontiguousarray(offset, dtype=np.uint64)
    lidx = np.ascontiguousarray(lidx, dtype=np.uint32)
    value = _f32(value)
    weights = _f32(weights)
    nrows = len(offset) - 1
    pred = np.zeros(nrows, np.float32)
    xv = np.zeros(max(nrows * V_dim, 1), np.float32) if want_xv else None
    npos = 0 if w_pos is None else len(w_pos)
    orc().orc_fm_predict(V_dim, nrows, _p(offset), _p(lidx), _p(value), _p(weights), len(weights),
                         _p(w_pos), _p(V_pos), npos, _p(pred), _p(xv))
    return (pred, xv[:nrows * V_dim].reshape(nrows, V_dim)) if want_xv else pred


def evaluate(label, pred):
    label, pred = _f32(label), _f32(pred)
    return float(orc().orc_evaluate(_p(label), _p(pred), len(pred)))


def fm_calc_grad(V_dim, offset, lidx, value, label, weights, pred, w_pos=None, V_pos=None):
    offset = np.ascontiguousarray(offset, dtype=np.uint64)
    lidx = np.ascontiguousarray(lidx, dtype=np.uint32)
    value, label, weights, pred = _f32(value), _f32(label), _f32(weights), _f32(pred)
    nrows = len(offset) - 1
    grad = np.zeros(len(weights), np.float32)
    npos = 0 if w_pos is None else len(w_pos)
    orc().orc_fm_calc_grad(V_dim, nrows, _p(offset), _p(lidx), _p(value), _p(label), _p(weights),
                           len(weights), _p(w_pos), _p(V_pos), npos, _p(pred), _p(grad))
    return grad


def auc(label, pred):
    label, pred = _f32(label), _f32(pred)
    return float(orc().orc_auc(_p(label), _p(pred), len(pred)))


def penalty(param, weights, w_pos=None, V_pos=None):
    weights = _f32(weights)
    npos = 0 if w_pos is None else len(w_pos)
    return float(orc().orc_penalty(C.byref(param), _p(weights), len(weights), _p(w_pos), _p(V_pos), npos))


class Oracle:
    """SGDUpdater + FMLoss restated (one 'server' worth of model state)."""

    def __init__(self, **kw):
        self.param = make_param(**kw)
        self.V_dim = self.param.V_dim
        self.h = orc().orc_model_create(C.byref(self.param))

    def __del__(self):
        if getattr(self, "h", None):
            orc().orc_model_free(self.h)
            self.h = None

    def size(self):
        return orc().orc_model_size(self.h)

    def seed(self):
        return orc().orc_model_seed(self.h)

    def get(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        n = len(keys)
        vals = np.zeros(max(n * (1 + self.V_dim), 1), np.float32)
        lens = np.zeros(max(n, 1), np.int32)
        nl = C.c_size_t(0)
        nv = orc().orc_get(self.h, _p(keys), n, _p(vals), _p(lens), C.byref(nl))
        return vals[:nv].copy(), lens[:nl.value].copy()

    def update_feacnt(self, keys, cnt):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        cnt = _f32(cnt)
        orc().orc_update_feacnt(self.h, _p(keys), len(keys), _p(cnt))

    def update_grad(self, keys, grads, lens):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        grads = _f32(grads)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        rc = orc().orc_update_grad(self.h, _p(keys), len(keys), _p(grads), len(grads), _p(lens), len(lens))
        if rc != 0:
            raise RuntimeError("reference CHECK would have failed in SGDUpdater::Update")

    def lookup(self, key):
        scal = np.zeros(4, np.float32)
        V2k = np.zeros(max(2 * self.V_dim, 1), np.float32)
        rc = orc().orc_model_lookup(self.h, int(key), _p(scal), _p(V2k))
        if rc < 0:
            return None
        k = self.V_dim
        return dict(fea_cnt=scal[0], w=scal[1], sqrt_g=scal[2], z=scal[3],
                    V=V2k[:k].copy() if rc == 1 else None, cg=V2k[k:2 * k].copy() if rc == 1 else None)

    def predict(self, offset, lidx, value, weights, w_pos=None, V_pos=None):
        return fm_predict(self.V_dim, offset, lidx, value, weights, w_pos, V_pos)

    def calc_grad(self, offset, lidx, value, label, weights, pred, w_pos=None, V_pos=None):
        return fm_calc_grad(self.V_dim, offset, lidx, value, label, weights, pred, w_pos, V_pos)

    def sgd_step(self, offset, index, value, label, is_train=True, push_cnt=False, progress=None):
        offset = np.ascontiguousarray(offset, dtype=np.uint64)
        index = np.ascontiguousarray(index, dtype=np.uint64)
        value, label = _f32(value), _f32(label)
        if progress is None:
            progress = np.zeros(5, np.float32)
        orc().orc_sgd_step(self.h, len(offset) - 1, _p(offset), _p(index), _p(value), _p(label),
                           int(is_train), int(push_cnt), _p(progress))
        return progress


# ---------------------------------------------------------------------------
# the compiled reference
# ---------------------------------------------------------------------------
_REF = None


def ref():
    global _REF
    if _REF is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libdifacto_ref.so not built (needs /root/reference)")
        L = C.CDLL(LIB_REF)
        vp, sz, u64 = C.c_void_p, C.c_size_t, C.c_uint64
        L.ref_reverse_bytes.restype = u64
        L.ref_reverse_bytes.argtypes = [u64]
        L.ref_read_batch.restype = vp
        L.ref_read_batch.argtypes = [C.c_char_p, C.c_char_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint,
                                     C.c_float, C.c_uint]
        L.ref_batch_rows.restype = sz
        L.ref_batch_rows.argtypes = [vp]
        L.ref_batch_nnz.restype = sz
        L.ref_batch_nnz.argtypes = [vp]
        L.ref_batch_has_value.argtypes = [vp]
        L.ref_batch_copy.argtypes = [vp, vp, vp, vp, vp]
        L.ref_batch_free.argtypes = [vp]
        L.ref_localize.restype = sz
        L.ref_localize.argtypes = [sz, vp, vp, vp, vp, u64, C.c_int, vp, vp, vp, vp]
        L.ref_engine_create.restype = vp
        L.ref_engine_create.argtypes = [vp, vp, C.c_int, C.c_int]
        L.ref_engine_destroy.argtypes = [vp]
        L.ref_updater_get.restype = sz
        L.ref_updater_get.argtypes = [vp, vp, sz, vp, vp, C.POINTER(sz)]
        L.ref_updater_update.argtypes = [vp, vp, sz, C.c_int, vp, sz, vp, sz]
        L.ref_fm_predict.argtypes = [vp, sz, vp, vp, vp, vp, vp, sz, vp, vp, sz, vp]
        L.ref_fm_calc_grad.argtypes = [vp, sz, vp, vp, vp, vp, vp, sz, vp, vp, sz, vp, vp]
        L.ref_evaluate.restype = C.c_float
        L.ref_evaluate.argtypes = [vp, vp, vp, sz]
        L.ref_auc.restype = C.c_float
        L.ref_auc.argtypes = [vp, vp, sz]
        L.ref_sgd_step.argtypes = [vp, sz, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]
        L.ref_sgd_learner_run.restype = C.c_int
        L.ref_sgd_learner_run.argtypes = [vp, vp, C.c_int, vp, C.c_int]
        _REF = L
    return _REF


def _kwargs_arrays(kw):
    ks = [str(k).encode() for k in kw.keys()]
    vs = [str(v).encode() for v in kw.values()]
    n = len(ks)
    ka = (C.c_char_p * n)(*ks)
    va = (C.c_char_p * n)(*vs)
    return ka, va, n


def ref_read_batch(path, fmt="libsvm", part=0, nparts=1, batch_size=100, shuffle_buf=0,
                   neg_sampling=1.0, which=0):
    L = ref()
    h = L.ref_read_batch(path.encode(), fmt.encode(), part, nparts, batch_size, shuffle_buf,
                         neg_sampling, which)
    if not h:
        return None
    n, nnz = L.ref_batch_rows(h), L.ref_batch_nnz(h)
    off = np.zeros(n + 1, np.uint64)
    lab = np.zeros(n, np.float32)
    idx = np.zeros(nnz, np.uint64)
    val = np.zeros(nnz, np.float32) if L.ref_batch_has_value(h) else None
    L.ref_batch_copy(h, _p(off), _p(lab), _p(idx), _p(val))
    L.ref_batch_free(h)
    return off, lab, idx, val


def ref_localize(offset, index, value=None, label=None, max_index=0xFFFFFFFFFFFFFFFF, nthreads=2,
                 want_cnt=True):
    offset = np.ascontiguousarray(offset, dtype=np.uint64)
    index = np.ascontiguousarray(index, dtype=np.uint64)
    value = _f32(value)
    nrows = len(offset) - 1
    nnz = int(offset[-1])
    label = np.zeros(nrows, np.float32) if label is None else _f32(label)
    lidx = np.zeros(max(nnz, 1), np.uint32)
    ooff = np.zeros(nrows + 1, np.uint64)
    keys = np.zeros(max(nnz, 1), np.uint64)
    cnt = np.zeros(max(nnz, 1), np.float32)
    U = ref().ref_localize(nrows, _p(offset), _p(index), _p(value), _p(label), max_index, nthreads,
                           _p(lidx), _p(ooff), _p(keys), _p(cnt) if want_cnt else None)
    return lidx[:nnz], keys[:U].copy(), (cnt[:U].copy() if want_cnt else None), ooff


def ref_auc(label, pred):
    label, pred = _f32(label), _f32(pred)
    return float(ref().ref_auc(_p(label), _p(pred), len(pred)))


class RefOracle:
    """The real SGDUpdater + FMLoss, same Python surface as ``Oracle``."""

    def __init__(self, nthreads=2, **kw):
        self.V_dim = int(kw.get("V_dim", 0))
        kw.setdefault("V_dim", 0)
        ka, va, n = _kwargs_arrays(kw)
        self.h = ref().ref_engine_create(ka, va, n, nthreads)

    def __del__(self):
        if getattr(self, "h", None):
            ref().ref_engine_destroy(self.h)
            self.h = None

    def get(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        n = len(keys)
        vals = np.zeros(max(n * (1 + self.V_dim), 1), np.float32)
        lens = np.zeros(max(n, 1), np.int32)
        nl = C.c_size_t(0)
        nv = ref().ref_updater_get(self.h, _p(keys), n, _p(vals), _p(lens), C.byref(nl))
        return vals[:nv].copy(), lens[:nl.value].copy()

    def update_feacnt(self, keys, cnt):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        cnt = _f32(cnt)
        ref().ref_updater_update(self.h, _p(keys), len(keys), K_FEA_COUNT, _p(cnt), len(cnt), None, 0)

    def update_grad(self, keys, grads, lens):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        grads = _f32(grads)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        ref().ref_updater_update(self.h, _p(keys), len(keys), K_GRADIENT, _p(grads), len(grads),
                                 _p(lens) if len(lens) else None, len(lens))

    def predict(self, offset, lidx, value, weights, w_pos=None, V_pos=None, label=None):
        offset = np.ascontiguousarray(offset, dtype=np.uint64)
        lidx = np.ascontiguousarray(lidx, dtype=np.uint32)
        value, weights = _f32(value), _f32(weights)
        nrows = len(offset) - 1
        label = np.zeros(nrows, np.float32) if label is None else _f32(label)
        pred = np.zeros(nrows, np.float32)
        npos = 0 if w_pos is None else len(w_pos)
        ref().ref_fm_predict(self.h, nrows, _p(offset), _p(lidx), _p(value), _p(label), _p(weights),
                             len(weights), _p(w_pos), _p(V_pos), npos, _p(pred))
        return pred

    def calc_grad(self, offset, lidx, value, label, weights, pred, w_pos=None, V_pos=None):
        """must follow predict() on the same batch (FMLoss keeps XV_/XX_ member state)"""
        offset = np.ascontiguousarray(offset, dtype=np.uint64)
        lidx = np.ascontiguousarray(lidx, dtype=np.uint32)
        value, label, weights, pred = _f32(value), _f32(label), _f32(weights), _f32(pred)
        nrows = len(offset) - 1
        grad = np.zeros(len(weights), np.float32)
        npos = 0 if w_pos is None else len(w_pos)
        ref().ref_fm_calc_grad(self.h, nrows, _p(offset), _p(lidx), _p(value), _p(label), _p(weights),
                               len(weights), _p(w_pos), _p(V_pos), npos, _p(pred), _p(grad))
        return grad

    def evaluate(self, label, pred):
        label, pred = _f32(label), _f32(pred)
        return float(ref().ref_evaluate(self.h, _p(label), _p(pred), len(pred)))

    def sgd_step(self, offset, index, value, label, is_train=True, push_cnt=False, progress=None,
                 seconds=None):
        offset = np.ascontiguousarray(offset, dtype=np.uint64)
        index = np.ascontiguousarray(index, dtype=np.uint64)
        value, label = _f32(value), _f32(label)
        if progress is None:
            progress = np.zeros(5, np.float32)
        ref().ref_sgd_step(self.h, len(offset) - 1, _p(offset), _p(index), _p(value), _p(label),
                           int(is_train), int(push_cnt), _p(progress), _p(seconds))
        return progress


def ref_sgd_learner_run(max_epochs=64, **kw):
    """run the reference SGDLearner; returns array [epochs, 10] of train/val Progress"""
    ka, va, n = _kwargs_arrays(kw)
    out = np.zeros((max_epochs, 10), np.float32)
    ne = ref().ref_sgd_learner_run(ka, va, n, _p(out), max_epochs)
    return out[:ne]
