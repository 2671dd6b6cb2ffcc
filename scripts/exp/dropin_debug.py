"""debug driver: one training step through the drop-in library, nothing silenced"""
import faulthandler, sys, os
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import ref_cuda_driver as rc
from oracle import ref_driver as rd
from tests import netdefs
from tests.common import rel_err
name = sys.argv[1] if len(sys.argv) > 1 else "mini_darknet"
mode = sys.argv[2] if len(sys.argv) > 2 else "off"
which = sys.argv[3] if len(sys.argv) > 3 else "dropin"
spec = {"mini_darknet": lambda: netdefs.mini_darknet(batch=4, size=16, classes=6), "tc_darknet": lambda: netdefs.tc_darknet(batch=4, size=16),
        "lenet": lambda: netdefs.lenet(batch=4, size=16, d1=24, d2=12), "lrn": netdefs.lrn_net}[name]()
kinds = [k for k, _ in spec["layers"]]
ref = rd.RefNet(spec, "C_BLAS")
gpu = rc.CudaBackendNet(spec, mode, which=which, quiet=False)
print("built", flush=True)
for l, k in enumerate(kinds):
    if k in ("conv", "dense"):
        gpu.set_weights(l, ref.weights_view(l))
if os.environ.get("SET_NORM"):
    for l, k in enumerate(kinds):
        if k == "norm":
            g = ref.norm_view(l, "gamma"); g[...] = 1.0 + 0.1 * np.cos(np.arange(g.size, dtype=np.float32))
            b = ref.norm_view(l, "beta"); b[...] = 0.05 * np.sin(np.arange(b.size, dtype=np.float32))
            gpu.set_norm(l, g, b)
print("weights set", flush=True)
x, t = rd.make_inputs(spec, seed=11)
ref.forward(x); gpu.forward(x)
print("forward done", flush=True)
for l in range(len(kinds)):
    a, b = gpu.output(l), ref.output(l)
    d = np.abs(a - b) / np.abs(b).max()
    print(l, kinds[l], "out err", d.max(), "q99", np.quantile(d, 0.99), "argmax", np.unravel_index(d.argmax(), d.shape), a.ravel()[d.argmax()], b.ravel()[d.argmax()], flush=True)
print("loss", gpu.loss(t).sum(), ref.loss(t).sum(), flush=True)
ref.backward(t, 0.02, 0.9, 0.0005); gpu.backward(t, 0.02, 0.9, 0.0005)
print("backward done", flush=True)
for l in range(len(kinds)):
    print(l, kinds[l], "delta err", rel_err(gpu.delta(l), ref.delta(l)), flush=True)
    if kinds[l] in ("conv", "dense"):
        print("   w err", rel_err(gpu.weights(l), ref.weights_view(l)), flush=True)
