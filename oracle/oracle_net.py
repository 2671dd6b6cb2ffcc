"""Whole-network runner built on the NumPy restatement (oracle/cianna_oracle.py): same spec dicts as
oracle/ref_driver.py, same step structure as the reference training loop (src/auxil.c:1797-1849:
forward over layers, output error, backprop in reverse with the optimizer applied per layer).

TEST INFRASTRUCTURE ONLY - it is the checker that travels to the GPU box; it is itself pinned against the
compiled reference (tests/test_oracle.py) and the fixtures in tests/golden/.
"""
import numpy as np

from . import cianna_oracle as co

_DEFAULT_BIAS = {"RELU": 0.1, "LOGI": -1.0, "SMAX": 0.1, "LIN": 0.5, "YOLO": 0.5}


def _activ_kind(s):
    for k in ("RELU", "LOGI", "SMAX", "YOLO", "LIN"):
        if s.startswith(k):
            return k
    return "LIN"


class OracleNet:
    def __init__(self, spec):
        self.spec = spec
        self.B = spec["batch"]
        W, H = spec["in_dim"]
        c, h, w = spec["in_ch"], H, W
        self.layers = []
        prev_kind = None
        for idx, (kind, a) in enumerate(spec["layers"]):
            L = dict(kind=kind, idx=idx, in_c=c, in_h=h, in_w=w)
            act = _activ_kind(a.get("activation", "RELU" if kind in ("conv", "dense") else "LIN"))
            L["act"] = act
            if kind == "conv":
                k = a["f_size"][0]
                L.update(k=k, stride=a.get("stride", (1, 1))[0], pad=a.get("padding", (0, 0))[0], n=a["nb_filters"])
                L["bias"] = spec.get("bias", 0.1) if idx == 0 else a.get("bias", _DEFAULT_BIAS[act])
                ho = (h + 2 * L["pad"] - k) // L["stride"] + 1
                wo = (w + 2 * L["pad"] - k) // L["stride"] + 1
                L["weights"] = np.zeros((L["n"], k * k * c + 1), dtype=np.float32)
                L["update"] = np.zeros_like(L["weights"])
                c, h, w = L["n"], ho, wo
            elif kind == "pool":
                if a.get("p_global", 0):
                    p, s = h, h
                else:
                    p = a.get("p_size", (2, 2))[0]
                    s = a.get("stride", (p, p))[0]
                L.update(p=p, stride=s, pad=a.get("padding", (0, 0))[0], type=a.get("p_type", "MAX"))
                h = (h + 2 * L["pad"] - p) // s + 1
                w = (w + 2 * L["pad"] - p) // s + 1
            elif kind == "norm":
                gs = a.get("group_size", 8)
                G = c // gs
                L.update(gs=gs, set_off=a.get("set_off", 0), gamma=np.ones(G, np.float32), beta=np.zeros(G, np.float32),
                         gamma_update=np.zeros(G, np.float32), beta_update=np.zeros(G, np.float32))
            elif kind == "dense":
                n = a["nb_neurons"]
                in_size = c * h * w + 1
                L.update(n=n, in_size=in_size, from_dense=(prev_kind == "dense"))
                L["bias"] = spec.get("bias", 0.1) if idx == 0 else a.get("bias", _DEFAULT_BIAS[act])
                L["weights"] = np.zeros((in_size, n + 1), dtype=np.float32)
                L["update"] = np.zeros_like(L["weights"])
                c, h, w = n, 1, 1
            else:
                raise ValueError(kind)
            # dropout (conv / pool / dense): active above 0.01 as upstream; the 0/1 mask of a training pass is GIVEN
            # (L["mask"], layout of the layer's output) because it is a random draw on either side
            L["drop"] = float(a.get("drop_rate", 0.0)) if kind != "norm" else 0.0
            L.update(out_c=c, out_h=h, out_w=w)
            self.layers.append(L)
            prev_kind = kind

    # ------------------------------------------------------------------
    def _dropout(self, L, pre):
        """src/naiv/naiv_pool_layer.c:320-356 and the conv / dense twins: the PRE-activation output times the mask
        (training, MC_MODEL inference) or times (1 - rate) (AVG_MODEL inference; a dense layer's bias node is left alone)"""
        if L["drop"] <= 0.01:
            return pre
        if self.inference:
            out = (pre * np.float32(1.0 - L["drop"])).astype(np.float32)
            if L["kind"] == "dense":
                out[:, -1] = pre[:, -1]
            return out
        return (pre * L["mask"]).astype(np.float32)

    def forward(self, x, length=None, inference=False):
        B = self.B
        length = B if length is None else length
        self.length = length
        self.inference = inference
        cur = x
        for L in self.layers:
            first = L["idx"] == 0
            L["input"] = cur
            if L["kind"] == "conv":
                pre, col = co.conv_forward(cur, L["weights"], first, B, L["in_c"], L["in_h"], L["in_w"], L["k"], L["stride"], L["pad"], L["bias"])
                L["col"] = col
                cur = self._activate(L, self._dropout(L, pre), length)
            elif L["kind"] == "pool":
                out, pmap = co.pool_forward(cur, B, L["in_c"], L["in_h"], L["in_w"], L["p"], L["stride"], L["pad"], L["type"])
                L["map"] = pmap
                cur = self._activate(L, self._dropout(L, out), length)
            elif L["kind"] == "norm":
                cur, L["mean"], L["var"] = co.group_norm_forward(cur, L["gamma"], L["beta"], L["gs"], L["set_off"], length)
            elif L["kind"] == "dense":
                if first:
                    flat = cur
                elif L["from_dense"]:
                    flat = cur
                else:
                    flat = co.flatten_for_dense(cur, L["bias"])
                L["flat"] = flat
                pre = self._dropout(L, co.dense_forward(flat, L["weights"]))
                if L["act"] == "RELU":
                    cur = co.relu_forward_dense(pre, length)
                elif L["act"] == "LOGI":
                    cur = co.logistic_forward_dense(pre, length)
                elif L["act"] == "SMAX":
                    cur = co.softmax_dense(pre, length)
                else:
                    cur = pre
            L["output"] = cur
        return cur

    @staticmethod
    def _activate(L, pre, length):
        if L["act"] == "RELU":
            return co.relu_forward(pre, length)
        if L["act"] == "LOGI":
            return co.logistic_forward(pre, length)
        if L["act"] == "SMAX":
            return co.softmax_conv(pre, length)
        return pre

    def _deriv(self, L, delta):
        """previous->deriv_activation applied to the delta that reaches layer L's output.
        L["deriv_value"] (optional) replaces the activated output the derivative is evaluated on: the mixed-precision
        parity tests put the product's own 16-bit activations there (and the product's argmax in L["map"]) so that both
        sides take the same DISCRETE decisions (leaky-ReLU slope, max-pool winner) and only the arithmetic is compared."""
        if L["act"] == "RELU":
            value = L.get("deriv_value", L["output"])
            if L["kind"] == "dense":
                return co.relu_deriv_dense(delta, value, self.length)
            return co.relu_deriv(delta, value, self.length)
        if L["act"] == "LOGI":
            if L["kind"] == "dense":
                return co.logistic_deriv_dense(delta, L["output"], self.length)
            return co.logistic_deriv(delta, L["output"], self.length)
        return delta

    def backward(self, target, lr, momentum=0.0, weight_decay=0.0):
        B = self.B
        last = self.layers[-1]
        if last["kind"] == "dense":
            delta = co.output_delta_dense(last["output"], target, self.length)
        else:
            delta = co.output_delta_conv(last["output"], target, self.length)
        if last["act"] in ("RELU", "LOGI"):
            # quadratic error of a non-linear output layer goes through the layer's own derivative
            # (ReLU_deriv_output_error / logistic_deriv_output_error, src/activ_functions.c:469-478,697-705)
            delta = self._deriv(last, delta)
        for L in reversed(self.layers):
            if L["drop"] > 0.01:      # the layer's delta is masked in place before anything reads it
                delta = (delta * L["mask"]).astype(np.float32)
            L["delta"] = delta
            first = L["idx"] == 0
            prev = self.layers[L["idx"] - 1] if not first else None
            d_prev = None
            if L["kind"] == "conv":
                if not first:
                    d_prev = co.conv_backward_data(delta, L["weights"], B, L["in_c"], L["in_h"], L["in_w"], L["k"], L["stride"], L["pad"])
                grad = co.conv_weight_grad(L["col"], delta)
                L["weights"], L["update"] = co.sgd_update(L["weights"], L["update"], grad, lr, B, momentum, weight_decay)
            elif L["kind"] == "pool":
                if not first:
                    d_prev = co.pool_backward(delta, L["map"], B, L["in_c"], L["in_h"], L["in_w"], L["p"], L["stride"], L["pad"], L["type"])
            elif L["kind"] == "norm":
                d_prev, L["d_gamma"], L["d_beta"] = co.group_norm_backward(L["input"], delta, L["gamma"], L["mean"], L["var"], L["gs"], L["set_off"], self.length)
                L["gamma"], L["beta"], L["gamma_update"], L["beta_update"] = co.group_norm_update(
                    L["gamma"], L["beta"], L["gamma_update"], L["beta_update"], L["d_gamma"], L["d_beta"], lr, B, momentum, L["set_off"])
            elif L["kind"] == "dense":
                if not first:
                    flat_d = (delta.astype(np.float64) @ L["weights"].astype(np.float64).T).astype(np.float32)  # [B][in_size]
                    if L["from_dense"]:
                        d_prev = flat_d
                    else:
                        C, A = L["in_c"], L["in_h"] * L["in_w"]
                        d_prev = flat_d[:, : C * A].reshape(B, C, A).transpose(1, 0, 2).copy()
                grad = L["flat"].astype(np.float64).T @ delta.astype(np.float64)
                L["weights"], L["update"] = co.sgd_update(L["weights"], L["update"], grad, lr, B, momentum, weight_decay, skip_last=1)
            if d_prev is not None:
                delta = self._deriv(prev, d_prev)

    def loss(self, target):
        last = self.layers[-1]
        if last["kind"] == "dense":
            if last["act"] == "SMAX":
                raise NotImplementedError
            return co.quadratic_dense(last["output"], target, self.length)
        if last["act"] == "SMAX":
            return co.cross_entropy_conv(last["output"], target, self.length)
        return co.quadratic_conv(last["output"], target, self.length)
