"""Network specs shared by the reference side (oracle/ref_driver.py) and the product side."""


def mini_darknet(batch=4, size=16, classes=6, in_ch=3, width=8):
    """Darknet19-shaped slice (examples/ImageNET/imagenet_train.py:45-92 upstream): C3x3-GN-MaxP, C3x3-GN-C1x1-GN-C3x3-GN-MaxP,
    C1x1 LIN head, global average pool + softmax."""
    w = width
    return dict(in_dim=(size, size), in_ch=in_ch, out_dim=classes, bias=0.1, batch=batch, layers=[
        ("conv", dict(f_size=(3, 3), nb_filters=w, padding=(1, 1), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=4)),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(3, 3), nb_filters=2 * w, padding=(1, 1), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=8)),
        ("conv", dict(f_size=(1, 1), nb_filters=w, padding=(0, 0), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=4)),
        ("conv", dict(f_size=(3, 3), nb_filters=2 * w, padding=(1, 1), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=8)),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(1, 1), nb_filters=classes, padding=(0, 0), activation="LIN")),
        ("pool", dict(p_type="AVG", p_global=1, activation="SMAX")),
    ])


def tc_darknet(batch=8, size=16, classes=16):
    """Same shape with channel counts that route every conv but the first through the tcgen05 kernels
    (>= 16 input channels, multiples of 16)."""
    return dict(in_dim=(size, size), in_ch=3, out_dim=classes, bias=0.1, batch=batch, layers=[
        ("conv", dict(f_size=(3, 3), nb_filters=32, padding=(1, 1), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=4)),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(3, 3), nb_filters=64, padding=(1, 1), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=8)),
        ("conv", dict(f_size=(1, 1), nb_filters=32, padding=(0, 0), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=8)),
        ("conv", dict(f_size=(3, 3), nb_filters=128, padding=(1, 1), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=16)),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(1, 1), nb_filters=classes, padding=(0, 0), activation="LIN")),
        ("pool", dict(p_type="AVG", p_global=1, activation="SMAX")),
    ])


def lenet(batch=8, size=28, d1=256, d2=128):
    """MNIST LeNet-5-like of examples/MNIST/mnist_train.py:67-73 upstream (dropout removed for parity runs)."""
    return dict(in_dim=(size, size), in_ch=1, out_dim=10, bias=0.1, batch=batch, layers=[
        ("conv", dict(f_size=(5, 5), nb_filters=8, padding=(2, 2), activation="RELU")),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(5, 5), nb_filters=16, padding=(2, 2), activation="RELU")),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("dense", dict(nb_neurons=d1, strict_size=1, activation="RELU")),
        ("dense", dict(nb_neurons=d2, strict_size=1, activation="RELU")),
        ("dense", dict(nb_neurons=10, strict_size=1, activation="SMAX")),
    ])


def darknet19(batch, size=448, classes=1000):
    """The north-star network, examples/ImageNET/imagenet_train.py:45-92 upstream."""
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
