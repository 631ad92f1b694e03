"""``mlx.optimizers`` stand-in: Adam as MLX defines it — betas (0.9, 0.999), eps 1e-8, NO bias correction by default:
m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g^2;  p = p - lr * m / (sqrt(v) + eps)."""
from __future__ import annotations

from . import core as mx


class Optimizer:
    def __init__(self):
        self.state = {"step": mx.array(0, dtype=mx.int64)}

    def update(self, model, gradients):
        model.update(self.apply_gradients(gradients, model))

    def apply_gradients(self, gradients, parameters):
        self.state["step"] = self.state["step"] + 1

        def walk(g, p, st):
            out = {}
            for k, gv in g.items():
                if isinstance(gv, dict):
                    out[k] = walk(gv, p[k], st.setdefault(k, {}))
                else:
                    if k not in st:
                        st[k] = self.init_single(p[k])
                    out[k] = self.apply_single(gv, p[k], st[k])
            return out
        return walk(gradients, parameters, self.state)


class Adam(Optimizer):
    def __init__(self, learning_rate, betas=(0.9, 0.999), eps=1e-8, bias_correction=False):
        super().__init__()
        self.learning_rate = learning_rate
        self.betas, self.eps, self.bias_correction = list(betas), eps, bias_correction
        self.state["learning_rate"] = mx.array(float(learning_rate))

    def init_single(self, parameter):
        return {"m": mx.zeros_like(parameter), "v": mx.zeros_like(parameter)}

    def apply_single(self, gradient, parameter, state):
        lr, (b1, b2), eps = float(self.learning_rate), self.betas, self.eps
        m = b1 * state["m"] + (1 - b1) * gradient
        v = b2 * state["v"] + (1 - b2) * mx.square(gradient)
        state["m"], state["v"] = m, v
        if self.bias_correction:
            step = int(self.state["step"])
            num = (lr / (1 - b1 ** step)) * m
            den = mx.sqrt(v) / ((1 - b2 ** step) ** 0.5) + eps
            return parameter - num / den
        return parameter - lr * m / (mx.sqrt(v) + eps)
