"""CPU restatement of the reference's YOLO output head - TEST INFRASTRUCTURE ONLY (tests/, smoke()).

Follows src/activ_functions.c (CPU) / src/cuda/cuda_activ_functions.cu (CUDA): activation :1480-1600 / :477-597,
overlap measures :1035-1126 / :599-698, association + error signal :1602-2315 / :700-1406, loss monitor
:2318-2985 / :1409-2075.  Plain Python loops over (image, cell) in float32 arithmetic: small cases only.
Pinned against tests/golden/yolo_*.npz (made by the compiled reference, tests/golden/make_golden_yolo.py) in
tests/test_oracle.py.

The association is stated the way the CUDA product organises it (tables addressed by the target's own index, bit
sets for the box states) rather than with upstream's per-cell scratch arrays, so that a disagreement between this
file and the fixtures points at the restructuring and not at the kernel.  Only the deterministic association
branches are stated (rand_startup = 0, rand_prob = rand_prob_best_box_assoc = 0, min_prior_forced_scaling = 0).
"""
import numpy as np

f32 = np.float32
IOU, GIOU, DIOU, DIOU2 = 0, 1, 2, 3
DIST_IOU, DIST_SIZE, DIST_OFFSET = 0, 1, 2
IOU_NAMES = {"IoU": IOU, "GIoU": GIOU, "DIoU": DIOU, "DIoU2": DIOU2}
DIST_NAMES = {"IoU": DIST_IOU, "IOU": DIST_IOU, "SIZE": DIST_SIZE, "OFFSET": DIST_OFFSET}
DEFAULT_LIMITS = {IOU: (0.5, 0.1, 0.0, 0.0, 0.2, 0.2, 0.5, 0.3), GIOU: (0.4, -0.5, -1.0, -1.0, -0.3, -0.3, 0.4, 0.2),
                  DIOU: (0.3, -0.6, -1.0, -1.0, -0.5, -0.5, 0.3, 0.1), DIOU2: (0.3, -0.5, -1.0, -1.0, -0.4, -0.4, 0.3, 0.1)}


class YoloSetup:
    """defaults and overrides of set_yolo_params + set_yolo_activ (src/activ_functions.c:970-1032, :1129-1380)"""

    def __init__(self, y, in_dim, grid):
        self.nb_box, self.nb_class, self.nb_param = y["nb_box"], y.get("nb_class", 0), y.get("nb_param", 0)
        self.max_nb_obj = y["max_nb_obj_per_image"]
        self.diff_flag, self.class_softmax = y.get("diff_flag", 0), y.get("class_softmax", 0)
        ps = np.asarray(y["prior_size"], dtype=f32)
        self.fit_dim = y.get("fit_dim", 0) or ps.shape[0]
        self.IoU_type = IOU_NAMES.get(y.get("IoU_type", "empty"), GIOU)
        self.prior_dist_type = DIST_NAMES.get(y.get("prior_dist_type", "empty"), DIST_SIZE)
        self.complete = y.get("error_type", "empty") == "complete"
        self.strict = y.get("strict_box_size", 0)
        prior = np.zeros((self.nb_box, 3), dtype=f32)
        prior[:, : self.fit_dim] = ps[: self.fit_dim].T
        self.prior = np.maximum(prior, f32(1.0))
        self.noobj = np.asarray(y.get("prior_noobj_prob", [0.2] * self.nb_box), dtype=f32)
        self.pis = np.asarray(y.get("param_ind_scales", [1.0] * max(self.nb_param, 1)), dtype=f32)
        sc = np.array([2, 2, 1, 2, 1, 1], dtype=f32)
        u = np.asarray(y.get("error_scales", [-1.0] * 6), dtype=f32)
        self.scale = np.where(u > 0, u, sc).astype(f32)
        sm = np.array([[1, 6, -6], [1, 1.6, -1.6], [1, 6, -6], [1, 6, -6], [1, 6, -6], [1, 1.2, -0.2]], dtype=f32)
        if "slopes_and_maxes" in y:
            u = np.asarray(y["slopes_and_maxes"], dtype=f32).reshape(6, 3)
            sm[:, 0] = np.where(u[:, 0] > 0, u[:, 0], sm[:, 0])
            sm[:, 1] = np.where(u[:, 1] < 100000.0, u[:, 1], sm[:, 1])
            sm[:, 2] = np.where(u[:, 2] > -100000.0, u[:, 2], sm[:, 2])
        self.sm = sm
        lim = np.array(DEFAULT_LIMITS[self.IoU_type], dtype=f32)
        u = np.asarray(y.get("IoU_limits", [-2.0] * 8), dtype=f32)
        self.lim = np.where(u > -1.99, u, lim).astype(f32)
        fit = np.array([1, 1, 1, 1, 1 if self.nb_class > 0 else -1, 1 if self.nb_param > 0 else -1])
        u = np.asarray(y.get("fit_parts", [-2] * 6))
        self.fit = np.where(u > -2, u, fit)
        self.grid_w, self.grid_h = grid
        self.cell = (in_dim[0] // grid[0], in_dim[1] // grid[1], 1)
        self.per = 8 + self.nb_class + self.nb_param
        self.tlen = 7 + self.nb_param + self.diff_flag


def activation(s, x):
    """x: raw head values [C][B][cells] -> activated copy"""
    a = np.array(x, dtype=f32)
    sm = s.sm
    for c in range(a.shape[0]):
        col = c % s.per
        if col < 3:
            if s.fit_dim > col:
                v = np.clip(-sm[0, 0] * a[c], sm[0, 2], sm[0, 1])
                a[c] = f32(1) / (f32(1) + np.exp(v, dtype=f32))
            else:
                a[c] = 0.5
        elif col < 6:
            a[c] = np.clip(sm[1, 0] * a[c], sm[1, 2], sm[1, 1]) if s.fit_dim > col - 3 else 0.0
        elif col < 8 or (col < 8 + s.nb_class and not s.class_softmax):
            r = 2 if col == 6 else (3 if col == 7 else 4)
            v = np.clip(-sm[r, 0] * a[c], sm[r, 2], sm[r, 1])
            a[c] = f32(1) / (f32(1) + np.exp(v, dtype=f32))
        elif col < 8 + s.nb_class:
            if col == 8:
                blk = a[c: c + s.nb_class]
                e = np.exp(blk - blk.max(axis=0), dtype=f32)
                a[c: c + s.nb_class] = e / e.sum(axis=0, dtype=f32)
        else:
            a[c] = np.clip(sm[5, 0] * a[c], sm[5, 2], sm[5, 1])
    return a


def overlap(kind, o, t):
    iw = max(f32(0), min(o[3], t[3]) - max(o[0], t[0]))
    ih = max(f32(0), min(o[4], t[4]) - max(o[1], t[1]))
    idp = max(f32(0), min(o[5], t[5]) - max(o[2], t[2]))
    inter = f32(f32(iw * ih) * idp)
    uni = f32(f32(f32(abs(o[3] - o[0]) * abs(o[4] - o[1])) * abs(o[5] - o[2]))
              + f32(f32(abs(t[3] - t[0]) * abs(t[4] - t[1])) * abs(t[5] - t[2])) - inter)
    with np.errstate(divide="ignore", invalid="ignore"):
        if kind == IOU:
            return f32(inter / uni)
        ew = max(o[3], t[3]) - min(o[0], t[0])
        eh = max(o[4], t[4]) - min(o[1], t[1])
        ed = max(o[5], t[5]) - min(o[2], t[2])
        if kind == GIOU:
            enc = f32(f32(ew * eh) * ed)
            return f32(f32(inter / uni) - f32(f32(enc - uni) / enc))
        dx = f32((o[3] + o[0]) * f32(0.5)) - f32((t[3] + t[0]) * f32(0.5))
        dy = f32((o[4] + o[1]) * f32(0.5)) - f32((t[4] + t[1]) * f32(0.5))
        dz = f32((o[5] + o[2]) * f32(0.5)) - f32((t[5] + t[2]) * f32(0.5))
        dist = f32(f32(dx * dx + dy * dy) + dz * dz)
        diag = f32(f32(ew * ew + eh * eh) + ed * ed)
        if kind == DIOU:
            dist, diag = np.sqrt(dist), np.sqrt(diag)
        return f32(f32(inter / uni) - f32(dist / diag))


def _sat_log(r, lo, hi):
    return np.log(lo) if r < lo else (np.log(hi) if r > hi else np.log(r))


def prior_distance(s, prior, ts, lo, hi):
    if s.prior_dist_type == DIST_IOU:
        sg = np.array([-0.5, -0.5, -0.5, 0.5, 0.5, 0.5], dtype=f32)
        return f32(f32(1) - overlap(s.IoU_type, sg * np.tile(prior, 2), sg * np.tile(ts, 2)))
    if s.prior_dist_type == DIST_OFFSET:
        return f32(sum(abs(f32(_sat_log(f32(ts[l] / prior[l]), lo, hi))) for l in range(3)))
    d = ts - prior
    return f32(np.sqrt(f32(f32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])))


def _cell(s, a, t_row, b, cx, cy, mode):
    """one grid cell.  a: activated [C][B][cells]; returns (delta[C], state[nb_box], err[C], monitor[nb_box][2])"""
    nb, per, tlen = s.nb_box, s.per, s.tlen
    ci = cy * s.grid_w + cx
    out = a[:, b, ci]
    cell_pos = (cx, cy, 0)
    lo, hi = np.exp(s.sm[1, 2]), np.exp(s.sm[1, 1])
    loss_mode = mode == "loss"
    complete = (not loss_mode) or s.complete
    natural = loss_mode and not s.complete
    delta = np.zeros(out.shape[0], dtype=f32)
    err = np.zeros(out.shape[0], dtype=f32)
    mon = np.full((nb, 2), -1.0, dtype=f32)
    nb_obj = int(t_row[0])
    class_only = f32(-2.0)
    if nb_obj == -1:
        nb_obj, class_only = 1, s.lim[0]
    nb_obj = min(nb_obj, s.max_nb_obj)
    tg = t_row[1:]

    bx = np.zeros((nb, 6), dtype=f32)
    s_p_i, best = 0, f32(1e8)
    sg = np.array([-0.5, -0.5, -0.5, 0.5, 0.5, 0.5], dtype=f32)
    for k in range(nb):
        c = np.zeros(6, dtype=f32)
        for l in range(3):
            c[l] = f32(f32(out[k * per + l] + f32(cell_pos[l])) * f32(s.cell[l]))
            c[l + 3] = f32(s.prior[k, l] * np.exp(out[k * per + l + 3], dtype=f32))
        for l in range(6):
            bx[k, l] = f32(c[l % 3] + f32(sg[l] * c[3 + l % 3]))
        dist = np.sqrt(f32((s.prior[k] * s.prior[k]).sum(dtype=f32)))
        if dist < best:
            best, s_p_i = dist, k

    def in_cell(j):
        t = tg[j * tlen: (j + 1) * tlen]
        for l in range(3):
            if int(f32(f32(f32(t[4 + l] + t[1 + l]) * f32(0.5)) / f32(s.cell[l]))) != cell_pos[l]:
                return False
        return True

    lock1, lock2 = set(), set()
    table, allowed = {}, {}
    mine = []
    for j in range(nb_obj):
        ti = tg[j * tlen + 1: j * tlen + 7]
        own = in_cell(j)
        row = np.zeros(nb, dtype=f32)
        for k in range(nb):
            row[k] = overlap(s.IoU_type, bx[k], ti)
            if row[k] > s.lim[0]:
                lock1.add(k)
        if not own:
            continue
        mine.append(j)
        table[j] = row
        ok = set(range(nb))
        if complete and s.strict > 0:
            ts = ti[3:] - ti[:3]
            dp = np.array([prior_distance(s, s.prior[k], ts, lo, hi) for k in range(nb)], dtype=f32)
            for _ in range(s.strict):
                bd = f32(1e6)
                for k in range(nb):
                    if dp[k] > 0 and dp[k] < bd:
                        bd = dp[k]
                for k in range(nb):
                    if abs(f32(dp[k] - bd)) < f32(0.001):
                        dp[k] = -2.0
            ok = {k for k in range(nb) if dp[k] < -1}
        allowed[j] = ok

    for _ in range(len(mine)):
        max_iou, resp_box, resp_j = f32(-2.0), -1, -1
        for j in mine:
            for k in range(nb):
                if table[j][k] > max_iou and k in allowed[j]:
                    max_iou, resp_j, resp_box = table[j][k], j, k
        if resp_box == -1:
            continue
        t = tg[resp_j * tlen: (resp_j + 1) * tlen]
        ti = t[1:7]
        ts = ti[3:] - ti[:3]
        if complete and max_iou < s.lim[1]:
            dp = np.array([prior_distance(s, s.prior[k], ts, lo, hi) for k in range(nb)], dtype=f32)
            bd = min(f32(100000.0), dp.min())
            bv = f32(-2.0)
            for k in range(nb):
                if abs(f32(dp[k] - bd)) < f32(0.001) and table[resp_j][k] > bv:
                    bv, resp_box = table[resp_j][k], k
        table[resp_j][:] = -2.0
        max_iou = overlap(s.IoU_type, bx[resp_box], ti)
        if max_iou > f32(0.98):
            max_iou = f32(0.98)
        if class_only > -2.0:
            max_iou = class_only
        lo_ = resp_box * per
        diff = int(t[7 + s.nb_param]) if s.diff_flag else 0
        if s.diff_flag and diff > 0 and (natural or max_iou < s.lim[6] or out[lo_ + 7] < s.lim[7]):
            continue
        for j in mine:
            table[j][resp_box] = -2.0
        lock2.add(resp_box)
        want = np.zeros(6, dtype=f32)
        for l in range(3):
            want[l] = f32(f32(f32(f32(ti[l + 3] + ti[l]) * f32(0.5)) - f32(cell_pos[l] * s.cell[l])) / f32(s.cell[l]))
            want[l + 3] = _sat_log(f32(ts[l] / s.prior[resp_box, l]), lo, hi)
        geom_ok = class_only < -1.9 and (s.diff_flag == 0 or diff < 3)
        cls_ok = s.diff_flag == 0 or diff < 2
        cls = int(t[0]) - 1
        sc, sm, fit = s.scale, s.sm, s.fit
        if loss_mode:
            mon[resp_box] = (out[lo_ + 7], max_iou)
        for k in range(3):
            o, z = out[lo_ + k], out[lo_ + k + 3]
            if fit[0] == 1 and s.fit_dim > k and geom_ok:
                delta[lo_ + k] = sm[0, 0] * sc[0] * o * (f32(1) - o) * (o - want[k])
                err[lo_ + k] = f32(0.5) * sc[0] * (o - want[k]) * (o - want[k])
            elif fit[0] == 0 and s.fit_dim > k:
                delta[lo_ + k] = sm[0, 0] * sc[0] * o * (f32(1) - o) * (o - f32(0.5))
                err[lo_ + k] = f32(0.5) * sc[0] * o * o
            if fit[1] == 1 and s.fit_dim > k and geom_ok:
                delta[lo_ + k + 3] = sm[1, 0] * sc[1] * (z - want[k + 3])
                err[lo_ + k + 3] = f32(0.5) * sc[1] * (z - want[k + 3]) * (z - want[k + 3])
            elif fit[1] == 0 and s.fit_dim > k:
                delta[lo_ + k + 3] = sm[1, 0] * sc[1] * z
                err[lo_ + k + 3] = f32(0.5) * sc[1] * z * z
        o = out[lo_ + 6]
        if fit[2] == 1:
            if max_iou > s.lim[2]:
                delta[lo_ + 6] = sm[2, 0] * sc[2] * o * (f32(1) - o) * (o - f32(0.98))
            if max_iou > s.lim[2] or natural:
                err[lo_ + 6] = f32(0.5) * sc[2] * (o - f32(0.98)) * (o - f32(0.98))
        elif fit[2] == 0:
            delta[lo_ + 6] = sm[2, 0] * sc[2] * o * (f32(1) - o) * (o - f32(0.5))
            err[lo_ + 6] = f32(0.5) * sc[2] * (o - f32(0.5)) * (o - f32(0.5))
        o = out[lo_ + 7]
        if fit[3] == 1:
            e = float(o) - (1.0 + float(max_iou)) * 0.5
            if max_iou > s.lim[3]:
                delta[lo_ + 7] = f32(float(sm[3, 0] * sc[3] * o * (f32(1) - o)) * e)
            if max_iou > s.lim[3] or natural:
                err[lo_ + 7] = f32(float(f32(0.5) * sc[3]) * e * e)
        elif fit[3] == 0:
            delta[lo_ + 7] = sm[3, 0] * sc[3] * o * (f32(1) - o) * (o - f32(0.5))
            err[lo_ + 7] = f32(float(f32(0.5) * sc[3]) * (float(o) - 0.5) ** 2)
        for k in range(s.nb_class):
            o = out[lo_ + 8 + k]
            if fit[4] == 1:
                if max_iou > s.lim[4] and cls_ok:
                    if s.class_softmax:
                        delta[lo_ + 8 + k] = sc[4] * (o - f32(1.0 if k == cls else 0.0))
                    else:
                        delta[lo_ + 8 + k] = sm[4, 0] * sc[4] * o * (f32(1) - o) * (o - f32(0.98 if k == cls else 0.02))
                if (max_iou > s.lim[4] and cls_ok) or natural:
                    if s.class_softmax:
                        err[lo_ + 8 + k] = sc[4] * -np.log(max(o, f32(1e-7))) if k == cls else 0.0
                    else:
                        w = f32(0.98 if k == cls else 0.02)
                        err[lo_ + 8 + k] = f32(0.5) * sc[4] * (o - w) * (o - w)
            elif fit[4] == 0 and not s.class_softmax:
                delta[lo_ + 8 + k] = sm[4, 0] * sc[4] * o * (f32(1) - o) * (o - f32(0.5))
                err[lo_ + 8 + k] = f32(0.5) * sc[4] * (o - f32(0.5)) * (o - f32(0.5))
        for k in range(s.nb_param):
            o = out[lo_ + 8 + s.nb_class + k]
            if fit[5] == 1:
                w = t[7 + k]
                if max_iou > s.lim[5] and cls_ok:
                    delta[lo_ + 8 + s.nb_class + k] = s.pis[k] * sm[5, 0] * sc[5] * (o - w)
                if (max_iou > s.lim[5] and cls_ok) or natural:
                    err[lo_ + 8 + s.nb_class + k] = s.pis[k] * f32(0.5) * sc[5] * (o - w) * (o - w)
            elif fit[5] == 0:
                delta[lo_ + 8 + s.nb_class + k] = s.pis[k] * sm[5, 0] * sc[5] * (o - f32(0.5))
                err[lo_ + 8 + s.nb_class + k] = s.pis[k] * f32(0.5) * sc[5] * (o - f32(0.5)) * (o - f32(0.5))

    state = np.zeros(nb, dtype=np.int32)
    for k in range(nb):
        state[k] = 2 if k in lock2 else (1 if k in lock1 else 0)
        if state[k] == 2:
            continue
        lo_ = k * per
        delta[lo_: lo_ + per] = 0.0
        err[lo_: lo_ + per] = 0.0
        if state[k] == 0:
            for part, col in ((2, 6), (3, 7)):
                o = out[lo_ + col]
                w = f32(0.02) if s.fit[part] == 1 else f32(0.5)
                if s.fit[part] >= 0:
                    delta[lo_ + col] = s.sm[part, 0] * s.noobj[k] * s.scale[part] * o * (f32(1) - o) * (o - w)
                    err[lo_ + col] = f32(0.5) * s.noobj[k] * s.scale[part] * (o - w) * (o - w)
    return delta, state, err, mon


def run(s, a, t, mode, tc_scale=1.0):
    """a: activated head output [C][B][cells] (FP32), t: target rows [B][1 + max_nb_obj*tlen].
    mode "delta": returns (delta [C][B][cells], box state [B][cells][nb_box]);
    mode "loss":  returns (per-element loss [C][B][cells], monitor [B][cells][nb_box][2])"""
    C, B, cells = a.shape
    first = np.zeros((C, B, cells), dtype=f32)
    state = np.zeros((B, cells, s.nb_box), dtype=np.int32)
    mon = np.zeros((B, cells, s.nb_box, 2), dtype=f32)
    with np.errstate(over="ignore", under="ignore"):
        for b in range(B):
            for ci in range(cells):
                d, st, e, m = _cell(s, a, np.asarray(t[b], dtype=f32), b, ci % s.grid_w, ci // s.grid_w, mode)
                first[:, b, ci] = d * f32(tc_scale) if mode == "delta" else e
                state[b, ci], mon[b, ci] = st, m
    return (first, state) if mode == "delta" else (first, mon)
