"""Small host-side helpers shared by bench.py, __graft_entry__.py and the tests (product side: nothing from oracle/)."""
import os
import sys

import numpy as np


def i_ar(v):
    return np.array(v, dtype="int32")


def build_network(cnn, spec, comp_meth="C_CUDA", mixed_precision="off", network=None, dynamic_load=1, inference_only=0):
    """Builds `spec` (dict: in_dim, in_ch, out_dim, bias, batch, layers=[(kind, kwargs), ...]) through the CIANNA
    Python API of module `cnn` - the same call sequence a user script makes (examples/ImageNET/imagenet_train.py upstream)."""
    kw = {} if network is None else {"network": network}
    cnn.init(in_dim=i_ar(spec["in_dim"]), in_nb_ch=spec["in_ch"], out_dim=spec["out_dim"], bias=spec.get("bias", 0.1),
             b_size=spec["batch"], comp_meth=comp_meth, dynamic_load=dynamic_load, mixed_precision=mixed_precision,
             inference_only=inference_only, no_logo=1, **kw)
    if "yolo" in spec:    # YOLO head set-up: between init and the layers, as in upstream's detection scripts
        cnn.set_yolo_params(network=0 if network is None else network, **spec["yolo"])
    for kind, a in spec["layers"]:
        a = dict(a)
        for key in ("f_size", "stride", "padding", "int_padding", "p_size"):
            if key in a:
                a[key] = i_ar(a[key])
        a.update(kw)
        getattr(cnn, kind)(**a)


class Quiet:
    """silences C-level stdout (layer creation prints one block per layer, like upstream)"""

    def __enter__(self):
        sys.stdout.flush()
        self._fd = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *a):
        os.dup2(self._fd, 1)
        os.close(self._null)
        os.close(self._fd)
