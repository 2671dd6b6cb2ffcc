"""Network definitions of the BASELINE.json configs, as `spec` dicts for cianna_b200.utils.build_network."""


def darknet19(batch, size=448, classes=1000):
    """The north-star network: Darknet19 ImageNet classifier as defined upstream in
    examples/ImageNET/imagenet_train.py:45-92 (19 convs, 18 group-norms, 5 max-pools, global average pool + softmax)."""
    L = []

    def c(f, n, act="RELU"):
        L.append(("conv", dict(f_size=(f, f), nb_filters=n, padding=(f // 2, f // 2), activation=act)))

    def gn(g):
        L.append(("norm", dict(normalization="GN", group_size=g, set_off=0)))

    def mp():
        L.append(("pool", dict(p_size=(2, 2), p_type="MAX")))

    c(3, 32); gn(4); mp()
    c(3, 64); gn(8); mp()
    c(3, 128); gn(8); c(1, 64); gn(8); c(3, 128); gn(8); mp()
    c(3, 256); gn(16); c(1, 128); gn(16); c(3, 256); gn(16); mp()
    c(3, 512); gn(16); c(1, 256); gn(16); c(3, 512); gn(16); c(1, 256); gn(16); c(3, 512); gn(16); mp()
    c(3, 1024); gn(32); c(1, 512); gn(16); c(3, 1024); gn(32); c(1, 512); gn(16); c(3, 1024); gn(32)
    c(1, classes, "LIN")
    L.append(("pool", dict(p_type="AVG", p_global=1, activation="SMAX")))
    return dict(in_dim=(size, size), in_ch=3, out_dim=classes, bias=0.1, batch=batch, layers=L)


def lenet(batch=16, size=28, d1=256, d2=128):
    """MNIST LeNet-5-like of examples/MNIST/mnist_train.py:67-73 upstream (dropout removed)."""
    return dict(in_dim=(size, size), in_ch=1, out_dim=10, bias=0.1, batch=batch, layers=[
        ("conv", dict(f_size=(5, 5), nb_filters=8, padding=(2, 2), activation="RELU")),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(5, 5), nb_filters=16, padding=(2, 2), activation="RELU")),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("dense", dict(nb_neurons=d1, strict_size=1, activation="RELU")),
        ("dense", dict(nb_neurons=d2, strict_size=1, activation="RELU")),
        ("dense", dict(nb_neurons=10, strict_size=1, activation="SMAX")),
    ])
