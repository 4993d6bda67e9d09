"""CPU backend for difacto_b200.sharded.ShardedStore built on the oracle -- TEST INFRASTRUCTURE.

It lets the world_size-2 gloo tests exercise the sharding protocol (key-range partition, count
exchange, all_to_all_v splits, per-source application order) without a GPU.  The product only
ever constructs CudaBackend."""
import numpy as np
import torch

from oracle import oracle as O


def _keys_np(t):
    return t.numpy().view(np.uint64) if len(t) else np.zeros(0, np.uint64)


class OracleBackend:
    def __init__(self, **kw):
        self.M = O.Oracle(**kw)
        self.V_dim = self.M.V_dim
        self.ks = (self.V_dim + 3) // 4 * 4
        self.progress = np.zeros(5, np.float64)   # loss, penalty, auc, nnz_w, nrows

    def empty(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def feacnt(self, keys, cnt):
        if len(keys):
            self.M.update_feacnt(_keys_np(keys), cnt.numpy())

    def pull_rows(self, keys, w, hasv, V):
        k = self.V_dim
        vals, lens = self.M.get(_keys_np(keys))
        p = 0
        for i in range(len(keys)):
            w[i] = float(vals[p])
            p += 1
            if k and lens[i] > 1:
                V[i, :k] = torch.from_numpy(vals[p:p + k].copy())
                hasv[i] = 1
                p += k
            else:
                hasv[i] = -1

    def _ragged(self, w, hasv, V):
        k = self.V_dim
        U = len(w)
        if k == 0:
            return w.numpy().copy(), None, None, np.zeros(0, np.int32)
        lens = np.where(hasv.numpy() > 0, k + 1, 1).astype(np.int32)
        w_pos, V_pos = O.get_pos(lens)
        vals = np.zeros(int(lens.sum()), np.float32)
        vals[w_pos] = w.numpy()
        for i in range(U):
            if V_pos[i] >= 0:
                vals[V_pos[i]:V_pos[i] + k] = V[i, :k].numpy()
        return vals, w_pos, V_pos, lens

    def fm_step(self, batch, w, hasv, V, is_train, gw, gV):
        k = self.V_dim
        off = batch["off"].numpy().view(np.uint64)
        lidx = batch["lidx"].numpy().view(np.uint32)
        val = batch["val"].numpy() if batch.get("val") is not None else None
        lab = batch["lab"].numpy()
        vals, w_pos, V_pos, lens = self._ragged(w[:batch["U"]], hasv[:batch["U"]], V[:batch["U"]])
        pred = O.fm_predict(k, off, lidx, val, vals, w_pos, V_pos)
        self.progress[0] += O.evaluate(lab, pred)
        self.progress[1] += O.penalty(self.M.param, vals, w_pos, V_pos)
        self.progress[2] += O.auc(lab, pred)
        self.progress[4] += len(lab)
        if not is_train:
            return
        g = O.fm_calc_grad(k, off, lidx, val, lab, vals, pred, w_pos, V_pos)
        if k == 0:
            gw[:] = torch.from_numpy(g)
            return
        gw[:] = torch.from_numpy(g[w_pos])
        for i in range(batch["U"]):
            if V_pos[i] >= 0:
                gV[i, :k] = torch.from_numpy(g[V_pos[i]:V_pos[i] + k].copy())

    def push_rows(self, keys, gw, hasv, gV):
        if not len(keys):
            return
        k = self.V_dim
        if k == 0:
            self.M.update_grad(_keys_np(keys), gw.numpy(), np.zeros(0, np.int32))
            return
        lens = np.where(hasv.numpy() > 0, k + 1, 1).astype(np.int32)
        w_pos, V_pos = O.get_pos(lens)
        g = np.zeros(int(lens.sum()), np.float32)
        g[w_pos] = gw.numpy()
        for i in range(len(keys)):
            if V_pos[i] >= 0:
                g[V_pos[i]:V_pos[i] + k] = gV[i, :k].numpy()
        self.M.update_grad(_keys_np(keys), g, lens)
